"""Data-parallel correctness check, run under torchrun (>= 2 ranks, NCCL): the gradients parallel.GradSync hands to the optimizer
must be the mean over ranks of the rank-local gradients, in every mode (collect after backward / buckets fired on the side stream
as they complete), eager and inside the replayed CUDA graph; and after several replayed iterations every rank must hold the
same parameters.  Prints one JSON line on rank 0; exit code 1 on failure.  (tests/test_gpu_dp.py launches it when >= 2 GPUs exist.)"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import bench   # noqa: E402
import synth   # noqa: E402


def mark(msg):
    print('[check_dp r%s] %s' % (os.environ.get('RANK', '?'), msg), file=sys.stderr, flush=True)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('CHECK_DP_DUMP_S', '120')), exit=True)      # a hang prints every thread's stack and exits
    from fsv import model, parallel, trainer
    from fsv.networks import layers
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    torch.cuda.set_stream(torch.cuda.Stream())
    layers.SPECTRAL_GROUP_CHUNKS, layers.SPECTRAL_CHUNK_MIN_NUMEL = 3, 0
    opt = bench.make_opt('tiny')
    opt.gpu_ids = [local]
    torch.manual_seed(0)
    step = model.Vid2VidStep(opt)
    mods = [step.netG] + step.d_modules()
    for m in mods:
        m.train()
        parallel.broadcast_state(m)
    batch = {k: v.cuda() for k, v in synth.make('pose', 2, 64, 64, seed=100 + rank).items()}
    snap = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in mods]

    def restore():
        with torch.no_grad():
            for m, sd in zip(mods, snap):
                for k, v in m.state_dict().items():
                    v.copy_(sd[k])

    class NoStep:
        def step(self):
            pass

        def zero_grad(self):
            pass

    def g_grads(sync):
        restore()
        for p in step.netG.parameters():
            p.grad = None
        c = step.prepare(batch)
        step.generator_losses(batch, c)                  # first call records the spectral plan; the second runs on the chunked groups
        restore()
        g, _, _ = step.generator_losses(batch, c)
        trainer.loss_backward(g, NoStep(), sync)
        torch.cuda.synchronize()
        return [None if p.grad is None else p.grad.detach().clone() for p in step.netG.parameters()]

    mark('model built')
    local_g = g_grads(None)
    mark('local gradients done')
    want = []
    for g, p in zip(local_g, step.netG.parameters()):
        t = torch.zeros_like(p) if g is None else g.clone()
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        want.append(t)
    report, ok = {}, True
    for name, kw in (('collect', dict(stream_fire=False)), ('stream', dict(stream_fire=True))):
        sync = parallel.GradSync(step.netG.parameters(), bucket_mb=0.05, **kw)
        mark('mode %s: start' % name)
        got = g_grads(sync)
        mark('mode %s: done' % name)
        num = sum(float((a - b).double().square().sum()) for a, b in zip(got, want))
        den = sum(float(b.double().square().sum()) for b in want)
        worst = max(float((a - b).abs().max() / (b.abs().max() + 1e-20)) for a, b in zip(got, want) if float(b.abs().max()) > 1e-8)
        report[name] = dict(rel_l2=(num / den) ** 0.5, worst_rel_max=worst, buckets=len(sync.buckets),
                            fired_in_backward=None)
        ok &= report[name]['rel_l2'] < 1e-3
        for h in sync._hooks:
            h.remove()
        for p in step.netG.parameters():
            p.grad = None
    # replayed graph, both syncs inside: ranks must stay in lock-step (identical parameters) and the losses finite
    restore()
    optG, optD = trainer.make_step_optimizers(opt, step, capturable=True)
    syncG = parallel.GradSync(step.netG.parameters(), bucket_mb=0.05)
    syncD = parallel.GradSync(step.d_parameters(), bucket_mb=0.05)
    mark('graph: capture')
    graphed = trainer.GraphedStep(step, optG, optD, batch, sync_G=syncG, sync_D=syncD)
    mark('graph: replay')
    for _ in range(4):
        d, g, fake, _ = graphed(batch)
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().reshape(-1) for m in mods for p in m.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    report['graph'] = dict(param_spread=float((hi - lo).abs().max()), finite=bool(torch.isfinite(flat).all()),
                           stream_fire=syncG.stream_fire, losses=[float(v.mean()) for v in list(d.values()) + list(g.values())])
    ok &= report['graph']['param_spread'] == 0.0 and report['graph']['finite']
    report['ok'] = bool(ok)
    report['world'] = world
    if rank == 0:
        print(json.dumps(report))
    dist.barrier()
    torch.cuda.synchronize()
    del graphed, syncG, syncD, optG, optD          # the captured graph holds NCCL work: it must be gone before the communicator is torn down
    import gc
    gc.collect()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0 if ok else 1)          # (no destroy_process_group: a communicator that was captured into a CUDA graph may block in teardown)


if __name__ == '__main__':
    main()
