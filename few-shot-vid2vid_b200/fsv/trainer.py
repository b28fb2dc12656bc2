"""Optimizers, the training iteration and its CUDA-graph replay (train.py:55-62, base_model.py:39-48,201-211,259-279,
loss_collector.py:217-228).  The loss graph itself lives in fsv.model.Vid2VidStep.

    D-step: G forward under no_grad -> D (+ face D, temporal D) on [fake ; real] -> hinge -> backward -> Adam
    G-step: G forward -> D forward -> GAN + feature matching + warp + mask (+ pose / face) losses -> backward -> Adam
"""
import torch

from .model import Vid2VidStep, LOSS_NAMES_D, LOSS_NAMES_G  # noqa: F401

import os

OWN_ADAM = os.environ.get('FSV_OWN_ADAM', '1') != '0'      # fsv.optim.Adam (one launch on the C ABI); 0 = torch's fused Adam


def _adam(params, lr, betas, capturable):
    params = list(params)
    on_gpu = all(p.is_cuda for p in params)
    if OWN_ADAM and on_gpu:
        from .optim import Adam
        return Adam(params, lr=lr, betas=betas)
    return torch.optim.Adam(params, lr=lr, betas=betas, capturable=capturable and on_gpu, fused=True if on_gpu else None)


def make_step_optimizers(opt, step, capturable=False):
    """base_model.py:39-48 (TTUR) over netG and over netD [+ netDT] [+ netDf] (base_model.py:204-211,267-277).  Call again after
    ``step.init_temporal_model()`` -- the reference re-creates both optimizers there."""
    if opt.no_TTUR:
        beta1, beta2, g_lr, d_lr = opt.beta1, 0.999, opt.lr, opt.lr
    else:
        beta1, beta2, g_lr, d_lr = 0.0, opt.beta2, opt.lr / 2, opt.lr * 2
    return (_adam(step.netG.parameters(), g_lr, (float(beta1), float(beta2)), capturable),
            _adam(step.d_parameters(), d_lr, (float(beta1), float(beta2)), capturable))


def make_optimizers(opt, netG, netD, capturable=False):
    """Two-network form kept for callers that hold bare modules."""
    class _S:
        pass
    s = _S()
    s.netG, s.d_parameters = netG, (lambda: list(netD.parameters()))
    return make_step_optimizers(opt, s, capturable)


def loss_backward(losses, optimizer, grad_sync=None):
    """loss_collector.py:217-228: mean -> sum -> zero_grad -> backward -> [all-reduce] -> step."""
    loss = sum(torch.mean(v) for v in losses.values())
    if grad_sync is not None:
        grad_sync.zero()             # gradients live in the sync's flat buckets (views): zero them in place
        grad_sync.arm()
    else:
        optimizer.zero_grad()
    loss.backward()
    from . import ops
    ops.side_join()                  # (also done by the engine callback; explicit here so the optimizer below is ordered after it for certain)
    if grad_sync is not None:
        grad_sync()
    optimizer.step()
    return loss


# The discriminator update (D / face D / temporal D forward, backward, Adam) only needs the no-grad frame; the generator update's own
# generator forward needs neither it nor the new discriminator weights until its discriminator forward.  By default (FSV_DSTEP_STREAM=0 disables) the
# discriminator update runs on a second stream concurrently with that generator forward (a parallel branch of the CUDA graph).
DSTEP_STREAM = os.environ.get('FSV_DSTEP_STREAM', '1') != '0'


def train_iteration(step, optG, optD, batch, sync_G=None, sync_D=None):
    """train.py:58-62 for one frame: discriminator update, then generator update.  -> (d_losses, g_losses, fake, prevs_new)"""
    from . import ops
    c = step.prepare(batch)
    if not (DSTEP_STREAM and batch['tgt_image'].is_cuda):
        d_losses = step.discriminator_losses(batch, c)
        loss_backward(d_losses, optD, sync_D)
        g_losses, fake, prevs = step.generator_losses(batch, c)
        loss_backward(g_losses, optG, sync_G)
        return d_losses, g_losses, fake, prevs
    with torch.no_grad():
        gen_d = step.generate(c)                      # generator forward #1 (spectral u/v, BN statistics advance): main stream
    s3 = ops.branch_fork(gen_d[0], index=2)
    with torch.cuda.stream(s3):
        d_losses = step.discriminator_losses(batch, c, gen=gen_d)
        loss_backward(d_losses, optD, sync_D)
    gen_g = step.generate(c)                          # generator forward #2, concurrent with the discriminator update
    ops.branch_join(s3, *d_losses.values())           # the new discriminator weights are needed from here on
    g_losses, fake, prevs = step.generator_losses(batch, c, gen=gen_g)
    loss_backward(g_losses, optG, sync_G)
    return d_losses, g_losses, fake, prevs


class GraphedStep:
    """The whole training iteration (both optimizer updates and, on N > 1 ranks, the gradient all-reduces included) recorded
    once into a CUDA graph and replayed: thousands of kernel launches become one graph launch.  Inputs are copied into static
    buffers before each replay; the returned losses / frame are static tensors overwritten by each replay.

    Construction runs ``warmup`` real iterations on the example batch before capture (allocator / cuBLAS-free warm-up that
    capture needs); by default the training state they touch -- parameters, optimizer moments and step counters, BatchNorm
    running statistics, spectral-norm u / v -- is snapshotted before and restored after, so constructing the graph does not
    advance training.  It must be constructed collectively (every rank, same order) when gradient syncs are passed."""

    def __init__(self, step, optG, optD, example, sync_G=None, sync_D=None, warmup=3, preserve_state=True):
        self.static = {k: v.clone() for k, v in example.items()}
        st = self.static

        def run():
            return train_iteration(step, optG, optD, st, sync_G=sync_G, sync_D=sync_D)
        # Capture on the stream the eager iterations already ran on (it must not be the legacy default stream): autograd's
        # AccumulateGrad nodes remember the stream they were created on, and a cross-stream wait would invalidate capture.
        cur = torch.cuda.current_stream()
        if cur == torch.cuda.default_stream():
            raise RuntimeError('GraphedStep: run the training loop under a non-default stream (torch.cuda.set_stream)')
        mods = [step.netG] + step.d_modules()
        snap = None
        if preserve_state:
            snap = ([{k: v.detach().clone() for k, v in m.state_dict().items()} for m in mods],
                    [_clone_opt_state(o) for o in (optG, optD)])
        for _ in range(warmup):
            run()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=cur):
            self.d_losses, self.g_losses, self.fake, self.prevs = run()
        if snap is not None:
            # restore IN PLACE (the graph holds the parameter / moment / buffer addresses)
            with torch.no_grad():
                for m, sd in zip(mods, snap[0]):
                    for k, v in m.state_dict().items():
                        v.copy_(sd[k])
                for o, s in zip((optG, optD), snap[1]):
                    _restore_opt_state(o, s)
        torch.cuda.synchronize()

    def __call__(self, inp):
        for k, v in inp.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.d_losses, self.g_losses, self.fake, self.prevs


def _clone_opt_state(o):
    return [{k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in o.state.get(p, {}).items()} for g in o.param_groups for p in g['params']]


def _restore_opt_state(o, snap):
    i = 0
    for g in o.param_groups:
        for p in g['params']:
            cur = o.state.get(p, {})
            for k, v in snap[i].items():
                if torch.is_tensor(v) and k in cur:
                    cur[k].copy_(v)
            for k in list(cur):
                if k not in snap[i] and torch.is_tensor(cur[k]):
                    cur[k].zero_()          # state created during warm-up (first step): back to its initial zeros
            i += 1


# ------------------------------------------------------------------------------------------------ bare-module helpers
def _batch(tgt_label, tgt_image, ref_labels, ref_images, prev=None):
    b = dict(tgt_label=tgt_label.unsqueeze(1), tgt_image=tgt_image.unsqueeze(1), ref_label=ref_labels, ref_image=ref_images)
    if prev is not None:
        b.update(prev_label=prev[0].unsqueeze(1), prev_real=prev[1].unsqueeze(1), prev_fake=prev[1].unsqueeze(1))
    return b


def discriminator_losses(opt, netG, netD, tgt_label, tgt_image, ref_labels, ref_images, prev=None, netDf=None, netDT=None):
    """4-D tensor form (tgt_label (B,C,H,W), tgt_image (B,3,H,W)) on caller-held modules; ``prev`` = [prev_label, prev_fake_image]."""
    step = Vid2VidStep(opt, netG, netD, netDf, netDT)
    return step.discriminator_losses(_batch(tgt_label, tgt_image, ref_labels, ref_images, prev))


def generator_losses(opt, netG, netD, tgt_label, tgt_image, ref_labels, ref_images, prev=None, netDf=None, netDT=None):
    step = Vid2VidStep(opt, netG, netD, netDf, netDT)
    g, fake, _ = step.generator_losses(_batch(tgt_label, tgt_image, ref_labels, ref_images, prev))
    return g, fake


def train_step(opt, netG, netD, optG, optD, tgt_label, tgt_image, ref_labels, ref_images, sync_G=None, sync_D=None, prev=None):
    step = Vid2VidStep(opt, netG, netD)
    d, g, fake, _ = train_iteration(step, optG, optD, _batch(tgt_label, tgt_image, ref_labels, ref_images, prev), sync_G, sync_D)
    return sum(v.mean() for v in d.values()).detach(), sum(v.mean() for v in g.values()).detach(), fake
