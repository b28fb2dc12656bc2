"""The reference arms of bench.py: the UNMODIFIED reference (baseline/_ref) timed through its own public path --
`create_model(opt)` -> `model(data_list, mode='discriminator'|'generator')` + `loss_backward` exactly as train.py:55-62 --
on the GPU (`--impl reference-gpu`: cuDNN/ATen, cudnn.benchmark=True as train.py:25 sets it, TF32 on or off) or on the host
cores (`--impl reference`: the same code with `.cuda()` turned into a no-op; reported baseline, not a target).

Imports neither fsv nor oracle: nothing of this repo's product runs in these arms.
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import refenv  # noqa: E402
import synth   # noqa: E402


def _extra(wl):
    return list(wl.get('ref_extra', []))


def build(wl, batch, gpu, seed=0):
    opt = refenv.parse_opt(wl['kind'], wl['H'], wl['W'], batch, extra=_extra(wl), gpu=gpu, vgg=bool(wl.get('vgg')))
    if not gpu:
        cpu_mode()
    model, opt_g, opt_d = (refenv.create_model(opt) if gpu else _create_cpu(opt))
    if wl.get('temporal'):
        (model.module if gpu else model).init_temporal_model()
        m = model.module if gpu else model
        opt_g, opt_d = m.optimizer_G, m.optimizer_D
    return opt, model, opt_g, opt_d


def cpu_mode():
    """CPU arm only: the reference hard-codes .cuda() (input_process.py:17-41, base_model.py:263, models.py:86); on the host
    cores those calls become no-ops.  Arithmetic is untouched."""
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


def _create_cpu(opt):
    from models.vid2vid_model import Vid2VidModel
    m = Vid2VidModel()
    m.initialize(opt, 0)
    if opt.no_vgg_loss and not hasattr(m.lossCollector, 'criterionVGG'):
        m.lossCollector.criterionVGG = lambda a, b: 0
    return m, m.optimizer_G, m.optimizer_D


def run_gpu(wl, batch, steps, warmup, tf32=True):
    """-> dict(ms_per_step, frames_per_s, e2e...) for the reference on cuda:0: K steps with device-resident inputs, then K steps
    end to end (pinned host inputs copied in, every loss read back)."""
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    opt, model, opt_g, opt_d = build(wl, batch, gpu=True)
    host = synth.make(wl['kind'], batch, wl['H'], wl['W'], seed=1234, K=wl.get('K', 1), temporal=bool(wl.get('temporal')))
    host = {k: v.pin_memory() for k, v in host.items()}
    dev = {k: v.cuda(non_blocking=True) for k, v in host.items()}
    dl = refenv.data_list(dev)
    d2h = [0]

    def step():
        return refenv.train_iteration(opt, model, opt_g, opt_d, dl)

    def step_e2e():
        d, g = refenv.train_iteration(opt, model, opt_g, opt_d, refenv.data_list(host, device='cuda'))
        out = torch.stack([x.detach().reshape(()) for x in list(d) + list(g)]).cpu()
        d2h[0] = out.numel() * out.element_size()
        return out

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out
    for _ in range(max(warmup, 3)):
        step()
    ms, out = timed(step, steps)
    step_e2e()
    ms_e, _ = timed(step_e2e, steps)
    return dict(ms_per_step=ms, frames_per_s=batch / (ms / 1e3), tf32=bool(tf32), batch=batch,
                losses=[float(x) for x in list(out[0]) + list(out[1])],
                e2e={'value': batch / (ms_e / 1e3), 'unit': 'frames/s', 'h2d_bytes_per_step': sum(v.numel() * v.element_size() for v in host.values()),
                     'd2h_bytes_per_step': d2h[0]})


def run_cpu(wl, batch, steps, warmup, threads):
    torch.set_num_threads(threads)
    opt, model, opt_g, opt_d = build(wl, batch, gpu=False)
    dl = refenv.data_list(synth.make(wl['kind'], batch, wl['H'], wl['W'], seed=1234, K=wl.get('K', 1), temporal=bool(wl.get('temporal'))))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        refenv.train_iteration(opt, model, opt_g, opt_d, dl)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return dict(ms_per_step=t * 1e3, frames_per_s=batch / t, batch=batch, threads=threads)
