#!/usr/bin/env python
"""bench.py -- frames/sec of one full G+D forward-backward training step of the few-shot vid2vid hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl fsv|reference] [--workload face256|face512|tiny]

A "step" is one reference training iteration for one frame (train.py:58-62): D-step (G forward under
no_grad, D on [fake;real], hinge, backward, Adam) then G-step (G forward, D forward, GAN + feature
matching + warp + mask losses, backward through D and G, Adam), flags --no_flow_gt --no_vgg_loss, single-frame
phase, synthetic face tensors.  Prints ONE JSON line (rank 0).

  value   : whole-job frames/s with the step's inputs already resident in HBM
  e2e     : same metric through the public API with HOST inputs: pinned-host -> device copies of every
            input and a device -> host read of the losses inside the timed region
  roofline: dominant kernel family of the step (conv, tensor-bound) measured with CUDA events in an
            instrumented pass; roofline_spade: the fused SPADE kernel (HBM-bound) -- BASELINE.json names both
  cpu_baseline / --impl reference: the CPU oracle (port of the reference's algorithm) timed on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'few-shot-vid2vid_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (H, W, per-GPU batch, option overrides)
    'face256': dict(H=256, W=256, batch=8, opt=dict()),
    'face512': dict(H=512, W=512, batch=2, opt=dict(fineSize=512)),
    # temporal phase (warp_prev: second flow pass, previous-frame embedding, 3-map SPADEs).  Parity of this
    # path is golden-tested on the GPU; this bench workload itself was added after the round-1 GPU budget was spent (not yet timed).
    'face256t': dict(H=256, W=256, batch=8, opt=dict(), temporal=True),
    'tiny': dict(H=64, W=64, batch=2, opt=dict(ngf=8, nff=8, ndf=8, n_downsample_G=4, n_adaptive_layers=3, n_blocks_F=2, fineSize=64)),
}

# options/base_options.py defaults + the BASELINE config flags (--adaptive_spade --warp_ref --spade_combine)
BASE_OPT = dict(
    n_downsample_G=5, n_downsample_A=2, ngf=32, norm_G='spectralspadesyncbatch', conv_ks=3, embed_ks=1, spade_ks=1,
    spade_combine=True, n_sc_layers=2, add_raw_output_loss=False, adaptive_spade=True, no_adaptive_embed=False,
    adaptive_conv=False, n_adaptive_layers=4, use_label_ref='mul', fineSize=256, aspect_ratio=1, n_fc_layers=2,
    label_nc=0, input_nc=1, output_nc=3, res_for_ref=False, netS='encoderdecoder', n_shot=1, lambda_kld=0.0,
    warp_ref=True, for_face=False, sc_arch='unet', norm_F='spectralsyncbatch', nff=32, n_blocks_F=6, n_downsample_F=3,
    flow_multiplier=20, isTrain=True, gpu_ids=[0], print_G=False, print_D=False, init_type='xavier', init_variance=0.02,
    netG='fewshot', n_frames_G=2, sep_flow_prev=False, no_sep_warp_embed=False, which_model_netD='multiscale',
    adaptive_D_layers=1, ndf=32, n_layers_D=4, num_D=1, norm_D='spectralinstance', netD_subarch='n_layers',
    gan_mode='hinge', lambda_feat=10.0, lambda_flow=10.0, lambda_mask=10.0, lambda_vgg=10.0, lambda_temp=0.0,
    no_ganFeat_loss=False, no_vgg_loss=True, no_flow_gt=True, dataset_mode='fewshot_face', add_face_D=False,
    lr=0.0004, beta1=0.5, beta2=0.999, no_TTUR=False)

# forward MACs per frame at face 256x256 from BASELINE.md section 2 (G 49.21 GMAC, D pair 2.37 GMAC); step = 4 G + 5 D
FLOP_PER_FRAME = {'face256': 417e9, 'face512': 1645e9, 'face256t': 4 * 149.9e9 + 5 * 4.73e9}


def make_opt(workload):
    from argparse import Namespace
    o = dict(BASE_OPT)
    o.update(WORKLOADS[workload]['opt'])
    return Namespace(**o)


def synth_inputs(workload, batch, seed, device='cpu', pin=False):
    """SURVEY.md section 8(d) face inputs: 1-channel edge maps in {0,1} (Bernoulli 0.03, 3x3 dilated), images U(-1,1)."""
    wl = WORKLOADS[workload]
    H, W = wl['H'], wl['W']
    g = torch.Generator().manual_seed(seed)

    def edges(*shape):
        e = (torch.rand(*shape, generator=g) < 0.03).float()
        return torch.nn.functional.max_pool2d(e.view(-1, 1, H, W), 3, 1, 1).view(*shape)
    t = dict(tgt_label=edges(batch, 1, H, W), tgt_image=torch.rand(batch, 3, H, W, generator=g) * 2 - 1,
             ref_labels=edges(batch, 1, 1, H, W), ref_images=torch.rand(batch, 1, 3, H, W, generator=g) * 2 - 1)
    if wl.get('temporal'):
        t.update(prev_label=edges(batch, 1, H, W), prev_image=torch.rand(batch, 3, H, W, generator=g) * 2 - 1)
    if pin and torch.cuda.is_available():
        t = {k: v.pin_memory() for k, v in t.items()}
    if device != 'cpu':
        t = {k: v.to(device) for k, v in t.items()}
    return t


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            f = [x.strip() for x in s.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tflops=p.get('bf16_tflops_sustained', p['bf16_tflops']), src='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tflops=1400.0, src='fallback (B200_PROFILING.md)')


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_step_time(workload, batch, threads, steps=1, warmup=0):
    """One reference-schedule step (D-step + G-step forward/backward, no optimiser) of the CPU oracle."""
    from oracle import nets as ON
    torch.set_num_threads(threads)
    opt = make_opt(workload)
    opt.gpu_ids = []
    from fsv import networks
    torch.manual_seed(0)
    g_cpu = networks.define_G(opt)
    if WORKLOADS[workload].get('temporal'):
        g_cpu.init_temporal_network()
    sdG = {k: v.detach().clone() for k, v in g_cpu.state_dict().items()}
    sdD = {k: v.detach().clone() for k, v in
           networks.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, opt.num_D, True, gpu_ids=[]).state_dict().items()}
    for sd in (sdG, sdD):
        for k, v in sd.items():
            if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
                v.requires_grad_(True)
    inp = synth_inputs(workload, batch, seed=1234)
    temporal = bool(WORKLOADS[workload].get('temporal'))
    prev = (inp['prev_label'], inp['prev_image']) if temporal else (None, None)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            fake = ON.generator_forward(sdG, opt, inp['tgt_label'], inp['ref_labels'], inp['ref_images'], prev=prev, training=True,
                                        temporal=temporal)[0]
        dl = ON.discriminator_losses(sdD, inp['tgt_label'], fake, inp['tgt_image'], inp['ref_labels'][:, 0], inp['ref_images'][:, 0],
                                     opt.n_layers_D, opt.num_D)
        sum(v.sum() for v in dl.values()).backward()
        gl, _ = ON.generator_losses(sdG, sdD, opt, inp['tgt_label'], inp['tgt_image'], inp['ref_labels'], inp['ref_images'],
                                    opt.n_layers_D, opt.num_D, prev=prev if temporal else None)
        sum(v.sum() for v in gl.values()).backward()
        for sd in (sdG, sdD):
            for v in sd.values():
                v.grad = None
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 32)   # more threads than this slows the small per-layer CPU convs down
    sample_batch = 1
    t = cpu_step_time(args.workload, sample_batch, threads, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    v = sample_batch / t
    wl = WORKLOADS[args.workload]
    line = {'impl': 'reference', 'metric': 'frames_per_sec_full_G+D_fwd_bwd', 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': '%s %dx%d --adaptive_spade --warp_ref --spade_combine --no_flow_gt --no_vgg_loss' % (args.workload, wl['H'], wl['W']),
                       'per_gpu_batch': wl['batch'], 'sample_batch': sample_batch},
            'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                             'sample': 'CPU oracle (torch-CPU port of the reference), one D-step+G-step fwd/bwd at batch %d, no optimiser' % sample_batch},
            'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(json.dumps(line))


# ----------------------------------------------------------------------------------------------------- GPU arm
def run_fsv(args):
    import torch.distributed as dist
    from fsv import networks, ops, parallel, trainer
    rank, world, local_rank = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_stream(torch.cuda.Stream())      # everything (eager, capture, replay, timing events) on one non-default stream
    wl = WORKLOADS[args.workload]
    batch = wl['batch'] if args.batch is None else args.batch
    opt = make_opt(args.workload)
    opt.gpu_ids = [local_rank]
    torch.manual_seed(0)
    netG = networks.define_G(opt)
    if wl.get('temporal'):
        netG.init_temporal_network()
    netD = networks.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, opt.num_D, True, gpu_ids=[local_rank])
    netG.train(), netD.train()
    parallel.broadcast_state(netG), parallel.broadcast_state(netD)
    use_graph = args.graph          # NCCL collectives are capturable too (one graph per rank, replayed in lock-step)
    optG, optD = trainer.make_optimizers(opt, netG, netD, capturable=use_graph)
    syncG = parallel.GradSync(netG.parameters()) if world > 1 else None
    syncD = parallel.GradSync(netD.parameters()) if world > 1 else None
    host = synth_inputs(args.workload, batch, seed=1234 + rank, pin=True)
    devin = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def eager_step(inp):
        prev = [inp['prev_label'], inp['prev_image']] if 'prev_label' in inp else None
        return trainer.train_step(opt, netG, netD, optG, optD, inp['tgt_label'], inp['tgt_image'], inp['ref_labels'], inp['ref_images'],
                                  sync_G=syncG, sync_D=syncD, prev=prev)
    step = eager_step
    n_eager0 = ops.LAUNCHES[0]
    eager_step(devin)
    launches_per_step = ops.LAUNCHES[0] - n_eager0
    graph_note = None
    if use_graph:
        try:
            step = trainer.GraphedStep(opt, netG, netD, optG, optD, devin, sync_G=syncG, sync_D=syncD)
        except Exception as e:      # capture is an optimisation of the launch path only; the eager path runs the same kernels
            graph_note = 'capture failed, eager launches used: %s' % str(e)[:200]
            use_graph = False
            torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup if args.quick else max(args.warmup, 3)):
        step(devin)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    n0 = ops.LAUNCHES[0]
    ms = timed(lambda: step(devin), args.steps)
    launches = launches_per_step      # C-ABI kernel-launching calls per step (counted on an eager step; a graph replays the same launches)
    clk = clocks.stop() if rank == 0 else None

    if args.quick:
        if rank == 0:
            emit(json.dumps({'quick': True, 'ms_per_step': ms / args.steps, 'note': 'profiling aid, not a bench value'}))
        return
    # e2e: host inputs (pinned) -> device every step, losses read back to the host every step
    d2h = [0]

    def e2e_step():
        inp = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        ld, lg, _ = step(inp)
        out = torch.stack([ld.detach(), lg.detach()]).cpu()
        d2h[0] = out.numel() * out.element_size()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    # instrumented pass: CUDA-event time per C-ABI kernel family (every rank runs the steps -- they contain the gradient
    # all-reduce -- but only rank 0 records events)
    prof = {}
    detail = {}
    spade_bytes = 0.0
    psteps = 2
    if rank == 0:
        real_call = ops._call
        pending = []
        from fsv._lib import lib

        def prof_call(fn, *a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            real_call(fn, *a)
            e.record()
            key = fn.__name__
            if key.startswith('fsv_conv2d') and key != 'fsv_conv2d_tc_eligible':
                cd = a[0]._obj
                key += ' %dx%d %d->%d k%d s%d up%d%s%s' % (cd.H, cd.W, cd.Cin, cd.Cout, cd.kh, cd.stride, cd.up,
                                                          ' persample' if cd.w_nstride else '', ' TC' if (cd.use_tc != 0 and lib.fsv_conv2d_tc_eligible(a[0])) else '')
            detail.setdefault(key, [0.0, 0])
            pending.append((fn.__name__, a, s, e, key))
        ops._call = prof_call
    for _ in range(psteps):
        eager_step(devin)
    torch.cuda.synchronize()
    if rank == 0:
        ops._call = real_call
        for name, a, s, e, key in pending:
            d = prof.setdefault(name, [0.0, 0])
            d[0] += s.elapsed_time(e) / psteps
            d[1] += 1.0 / psteps
            detail[key][0] += s.elapsed_time(e) / psteps
            detail[key][1] += 1.0 / psteps
            if name in ('fsv_spade_fwd', 'fsv_spade_fwd_tc'):
                sd = a[0]._obj
                px = sd.N * sd.H * sd.W
                spade_bytes += 4.0 * (px * sd.C / (sd.up * sd.up) + sum(px * sd.K[i] for i in range(sd.nmaps)) + px * sd.C) / psteps
    if world > 1:
        dist.barrier()

    if rank != 0:
        return
    if args.breakdown:
        with open(args.breakdown, 'w') as f:
            for k, v in sorted(detail.items(), key=lambda kv: -kv[1][0]):
                f.write('%9.3f ms  x%-5.1f %s\n' % (v[0], v[1], k))
    peaks = load_peaks()
    t_step = ms / args.steps / 1e3
    gbatch = batch * world
    value = gbatch / t_step
    flop = FLOP_PER_FRAME.get(args.workload)
    conv_ms = sum(v[0] for k, v in prof.items() if k.startswith('fsv_conv2d'))
    total_ms = sum(v[0] for v in prof.values())
    roof = None
    if flop:
        conv_flops = flop * batch      # conv/linear FLOPs of one step on this rank
        ach = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        roof = {'bound': 'tensor', 'kernel': 'fsv_conv2d_{fwd,dgrad,wgrad} (all launches of one step)', 'achieved': ach,
                'peak': peaks['tflops'], 'unit': 'TFLOP/s', 'frac': ach / peaks['tflops'], 'traffic': None,
                'traffic_note': 'aggregate over launches of many shapes, so no single per-launch figure; one captured launch '
                                '(profiles/ncu_full_r1_summary.txt, k_conv_tc grid 256): 33.9 MB DRAM read+write vs 34.2 MB algorithmic (x + y + w)',
                'peak_source': peaks['src'] + ', dense bf16 sustained', 'share_of_step_kernel_time': conv_ms / total_ms if total_ms else None}
    sp_ms = sum(prof[k][0] for k in ('fsv_spade_fwd', 'fsv_spade_fwd_tc') if k in prof)
    sp = [sp_ms]
    roof_spade = None
    if sp_ms > 0:
        ach = spade_bytes / (sp[0] / 1e3) / 1e9
        roof_spade = {'bound': 'hbm', 'kernel': 'fsv_spade_fwd[_tc] (all launches of one step)', 'achieved': ach, 'peak': peaks['hbm'],
                      'unit': 'GB/s', 'frac': ach / peaks['hbm'], 'traffic': None, 'peak_source': peaks['src']}
    line = {'metric': 'frames_per_sec_full_G+D_fwd_bwd', 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': t_step * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'fp32' if ops.CONV_USE_TC == 0 else 'tf32', 'data': 'synthetic',
            'config': {'workload': '%s %dx%d --adaptive_spade --warp_ref --spade_combine --no_flow_gt --no_vgg_loss' % (args.workload, wl['H'], wl['W']),
                       'per_gpu_batch': batch, 'global_batch': gbatch, 'parallelism': 'dp%d' % world, 'cuda_graph': bool(use_graph), 'cuda_graph_note': graph_note,
                       'l2': 'per-step working set (activations of a %d-frame batch) is far larger than the 126 MB L2' % batch},
            'clocks': clk, 'gpu_launches': int(launches),
            'e2e': {'value': gbatch / (ms_e2e / args.steps / 1e3), 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h[0]},
            'roofline': roof, 'roofline_spade': roof_spade,
            'kernel_ms_per_step': {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
    if world == 1 and not args.no_cpu_baseline:
        threads = min(os.cpu_count() or 1, 32)
        t_cpu = cpu_step_time(args.workload, 1, threads)
        line['cpu_baseline'] = {'value': 1.0 / t_cpu, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                                'sample': 'CPU oracle (torch-CPU port of the reference), one D-step+G-step fwd/bwd at batch 1, no optimiser'}
    emit(json.dumps(line))


_REAL_STDOUT = None


def quiet_stdout():
    """Route everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr: stdout carries the ONE JSON line only."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(text):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + '\n').encode())


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='fsv', choices=['fsv', 'reference'])
    ap.add_argument('--workload', default='face256', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=None, help='per-GPU batch (default: the workload\'s)')
    ap.add_argument('--simt', action='store_true', help='force the exact-fp32 SIMT conv path')
    ap.add_argument('--no-cpu-baseline', dest='no_cpu_baseline', action='store_true')
    ap.add_argument('--graph', dest='graph', action='store_true', default=True, help='replay the step from a CUDA graph (default, single GPU)')
    ap.add_argument('--no-graph', dest='graph', action='store_false')
    ap.add_argument('--quick', action='store_true', help='profiling aid: W warm-up + K timed steps only (no e2e / instrumented / CPU passes); not a bench result')
    ap.add_argument('--breakdown', default=None, help='write a per-kernel/per-shape time breakdown of one step to this file')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: the fsv arm needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    if args.simt:
        from fsv import ops
        ops.CONV_USE_TC = 0
    run_fsv(args)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
