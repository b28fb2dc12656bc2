#!/usr/bin/env python
"""ncu target: the dominant kernels on layers of KNOWN shape, so that the dram bytes of a capture can be set against the algorithmic bytes.

    ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc_p|k_spade_tc|k_wgrad_tc_mn' -o gpurun_out/prof_layers_r2 -f \
        python scripts/ncu_layers.py gpurun_out/ncu_layers.json

Layers (pose 512x512, batch 2, the shapes that carry the most time in the round-2 timeline):
  conv64   3x3 256 -> 256 at 64x64     forward (k_conv_tc_p, 128 tiles) + data gradient + weight gradient
  conv512  3x3 64 -> 32 at 512x512     forward + data gradient + weight gradient
  spade512 fused SPADE, 64 channels at 512x512 read through the x2 upsample, two 32-channel maps: forward + backward
Writes the algorithmic bytes / FLOPs of each launch class to the JSON given as argv[1] (scripts/summarize_profiles_final.py joins them
with the capture)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'few-shot-vid2vid_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    from fsv import ops
    from fsv.networks.layers import SPADE
    torch.cuda.set_stream(torch.cuda.Stream())
    ops.CONV_USE_TC = -1
    g = torch.Generator(device='cuda').manual_seed(0)
    info = {}

    def conv(name, n, h, w, cin, cout, reps=2):
        x = torch.randn(n, h, w, cin, device='cuda', generator=g).requires_grad_(True)
        wt = (torch.randn(cout, 3, 3, cin, device='cuda', generator=g) * 0.05).requires_grad_(True)
        for _ in range(reps):
            y = ops.conv2d(x, wt, None, pad=1, use_tc=-1)
            y.backward(torch.ones_like(y))
            ops.side_join()
            torch.cuda.synchronize()
        px = n * h * w
        info[name] = dict(shape='%dx%dx%d %d->%d k3' % (n, h, w, cin, cout), flops=2.0 * px * 9 * cin * cout,
                          fwd_bytes=4.0 * (px * cin + px * cout + 9 * cin * cout), dgrad_bytes=4.0 * (px * cout + px * cin + 9 * cin * cout),
                          wgrad_bytes=4.0 * (px * cin + px * cout + 9 * cin * cout))

    conv('conv64', 2, 64, 64, 256, 256)
    conv('conv512', 2, 512, 512, 64, 32)

    n, hs, c, ks = 2, 256, 64, [32, 32]
    mod = SPADE(c, ks, norm='spectralspadesyncbatch', ks=1, params_free=False).cuda()
    mod.train()
    x = torch.randn(n, hs, hs, c, device='cuda', generator=g).requires_grad_(True)
    maps = [torch.randn(n, 2 * hs, 2 * hs, k, device='cuda', generator=g).requires_grad_(True) for k in ks]
    for _ in range(2):
        y = mod(x, maps, None, up=2, act=ops.ACT_LRELU)
        y.backward(torch.ones_like(y))
        ops.side_join()
        torch.cuda.synchronize()
    px = n * 4 * hs * hs
    info['spade512'] = dict(shape='%dx%dx%d C=%d up=2 maps %s' % (n, 2 * hs, 2 * hs, c, ks),
                            fwd_bytes=4.0 * (px * c / 4 + sum(px * k for k in ks) + px * c),
                            bwd_bytes=4.0 * (px * c / 4 + sum(px * k for k in ks) + px * c + px * c + 2 * len(ks) * px * c),
                            note='bwd: reads x/4, maps, dout; writes dxhat and dgamma_i, dbeta_i per map')
    if len(sys.argv) > 1:
        json.dump(info, open(sys.argv[1], 'w'), indent=1)
    print(json.dumps(info))


if __name__ == '__main__':
    main()
