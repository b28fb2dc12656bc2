#!/usr/bin/env python
"""Bring-up experiment for halo-tile operand reuse in the tcgen05 convolution: does a UMMA shared-memory descriptor whose start address is
NOT aligned to the 1024-byte SWIZZLE_128B pattern (start + 128 B = one pixel row later) read the rows it should?  Run with
FSV_TC_PERSIST=0 FSV_TC_SHIFT_EXP={1,2}: the kernel loads every A box one pixel to the left and starts the descriptor one row later
(2: with base_offset = (start >> 7) & 7), so every output pixel except the last column of each 16-wide tile must still be right."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'few-shot-vid2vid_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    from fsv import ops
    g = torch.Generator().manual_seed(3)
    for (n, h, w, cin, cout) in [(2, 32, 32, 64, 64), (1, 16, 48, 32, 128)]:
        x = torch.randn(n, cin, h, w, generator=g)
        wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
        ref = F.conv2d(x.double(), wt.double(), padding=1).float()
        y = ops.conv2d(x.permute(0, 2, 3, 1).contiguous().cuda(), wt.permute(0, 2, 3, 1).contiguous().cuda(), None, pad=1, use_tc=-1)
        y = y.permute(0, 3, 1, 2).cpu()
        err = (y - ref).abs()
        cols = torch.arange(w)
        good = err[..., cols % 16 != 15].max().item()
        last = err[..., cols % 16 == 15].max().item()
        print('FSV_TC_SHIFT_EXP=%s shape %s: max |err| columns tw<15 %.3e, column tw==15 %.3e, ref max %.2f' %
              (os.environ.get('FSV_TC_SHIFT_EXP', '0'), (n, h, w, cin, cout), good, last, ref.abs().max().item()))
    # weight gradient (MN-major operands, K rows = pixels): dy is zero wherever it would meet the garbage K row (w % 32 == 31)
    for (n, h, w, cin, cout) in [(2, 32, 64, 64, 64), (1, 16, 96, 32, 128), (2, 32, 32, 128, 32)]:
        x = torch.randn(n, cin, h, w, generator=g).double()
        wt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.1).double().requires_grad_(True)
        go = torch.randn(n, cout, h, w, generator=g).double()
        go[..., torch.arange(w) % 32 == 31] = 0
        (F.conv2d(x, wt, padding=1) * go).sum().backward()
        xg = x.float().permute(0, 2, 3, 1).contiguous().cuda()
        wg = wt.detach().float().permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)
        yg = ops.conv2d(xg, wg, None, pad=1, use_tc=-1)
        yg.backward(go.float().permute(0, 2, 3, 1).contiguous().cuda())
        ops.side_join()
        torch.cuda.synchronize()
        err = (wg.grad.cpu().permute(0, 3, 1, 2).double() - wt.grad).abs().max().item()
        print('FSV_WG_SHIFT_EXP=%s shape %s: max |err| dW %.3e, ref max %.2f' % (os.environ.get('FSV_WG_SHIFT_EXP', '0'), (n, h, w, cin, cout), err,
                                                                             wt.grad.abs().max().item()))


if __name__ == '__main__':
    main()
