"""One training iteration of the hot path: the mirror of the reference's inner loop
(train.py:55-62 -> vid2vid_model.py:62-128 -> loss_collector.py:47-228) for the single-frame
phase with ``--no_flow_gt --no_vgg_loss`` on datasets without a foreground mask (face / street):

    D-step: G forward under no_grad -> D on [fake ; real] -> hinge real/fake -> backward -> Adam
    G-step: G forward -> D on [fake ; real] -> GAN + feature matching + warp L1 + mask losses -> backward -> Adam

The loss arithmetic is the reference's own torch code path (it sits ABOVE the define_G/define_D
boundary and is kept); the networks are the fsv drop-in modules.
"""
import torch

from . import ops


def _d_input(tgt_label, fake, real, ref_label, ref_image):
    """loss_collector.py:47-58,104-110 with concat_ref_for_D: batch [fake ; real], channels
    [ref_label, ref_image, tgt_label, image]."""
    tgt = torch.cat([fake, real], dim=0)
    tgt = torch.cat([tgt_label.repeat(2, 1, 1, 1), tgt], dim=1)
    ref = torch.cat([ref_label, ref_image], dim=1).repeat(2, 1, 1, 1)
    return torch.cat([ref, tgt], dim=1)


def _split(pred):
    """base_model.py:141-147 divide_pred."""
    fake = [[t[:t.size(0) // 2] for t in p] for p in pred]
    real = [[t[t.size(0) // 2:] for t in p] for p in pred]
    return fake, real


def _hinge(pred, target_is_real):
    """loss.py:69-78 (for_discriminator=True branch -- also what the reference uses for the generator's GAN
    term, because loss_collector.py:66 omits for_discriminator=False)."""
    z = pred * 0
    return -torch.mean(torch.min(pred - 1, z)) if target_is_real else -torch.mean(torch.min(-pred - 1, z))


def _gan_loss(preds, target_is_real):
    """loss.py:92-104."""
    loss = 0
    for p in preds:
        loss = loss + _hinge(p[-1], target_is_real).view(1)
    return loss / len(preds)


def _feat_match(pred_real, pred_fake, lambda_feat):
    """loss_collector.py:206-215."""
    num_d = len(pred_fake)
    loss = 0
    for i in range(num_d):
        for j in range(len(pred_fake[i]) - 1):
            loss = loss + torch.nn.functional.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) / num_d
    return loss * lambda_feat


def _masked_l1(a, b, m):
    m = m.expand_as(a)
    return torch.nn.functional.l1_loss(a * m, b * m)


def _mask_loss(flow_mask, warped, tgt, lambda_mask):
    """loss_collector.py:191-204."""
    conf = torch.clamp(1 - torch.sum(abs(warped - tgt), dim=1, keepdim=True), 0, 1)
    zero, one = torch.zeros_like(flow_mask), torch.ones_like(flow_mask)
    return (_masked_l1(flow_mask, zero, conf) + _masked_l1(flow_mask, one, 1 - conf)) * lambda_mask


def discriminator_losses(opt, netG, netD, tgt_label, tgt_image, ref_labels, ref_images, prev=None):
    """vid2vid_model.py:106-128.  ``prev`` = [prev_label, prev_image] in the temporal phase (warp_prev), else None."""
    with torch.no_grad():
        fake = (netG(tgt_label, ref_labels, ref_images) if prev is None else netG(tgt_label, ref_labels, ref_images, prev=prev))[0]
    pred = netD(_d_input(tgt_label, fake.detach(), tgt_image, ref_labels[:, 0], ref_images[:, 0]))
    pf, pr = _split(pred)
    return {'D_real': _gan_loss(pr, True), 'D_fake': _gan_loss(pf, False)}


def generator_losses(opt, netG, netD, tgt_label, tgt_image, ref_labels, ref_images, prev=None):
    """vid2vid_model.py:62-104 (non-zero terms under --no_flow_gt --no_vgg_loss, no foreground mask).  With ``prev`` =
    [prev_label, prev_image] (temporal phase) the warp / mask terms of the previous-frame branch are added
    (loss_collector.py:132-136,165-168: both entries of flow / flow_mask / warped image contribute)."""
    out = netG(tgt_label, ref_labels, ref_images) if prev is None else netG(tgt_label, ref_labels, ref_images, prev=prev)
    fake, flow, fmask, warp = out[0], out[1], out[2], out[4]
    pred = netD(_d_input(tgt_label, fake, tgt_image, ref_labels[:, 0], ref_images[:, 0]))
    pf, pr = _split(pred)
    losses = {'G_GAN': _gan_loss(pf, True), 'G_GAN_Feat': _feat_match(pr, pf, opt.lambda_feat)}
    if flow[0] is not None:
        losses['F_Warp'] = torch.nn.functional.l1_loss(warp[0], tgt_image) * opt.lambda_flow
        losses['F_Mask'] = _mask_loss(fmask[0], warp[0], tgt_image, opt.lambda_mask)
    if flow[1] is not None:
        zero = tgt_image.new_zeros(())
        losses['F_Warp'] = losses.get('F_Warp', zero) + torch.nn.functional.l1_loss(warp[1], tgt_image) * opt.lambda_flow
        losses['F_Mask'] = losses.get('F_Mask', zero) + _mask_loss(fmask[1], warp[1], tgt_image, opt.lambda_mask)
    return losses, fake


FUSED_ADAM = True


def make_optimizers(opt, netG, netD, capturable=False):
    """base_model.py:39-48 Adam with TTUR.  ``capturable`` keeps the step counters on the device so the whole
    iteration can be recorded into a CUDA graph (same arithmetic)."""
    if opt.no_TTUR:
        beta1, beta2, g_lr, d_lr = opt.beta1, 0.999, opt.lr, opt.lr
    else:
        beta1, beta2, g_lr, d_lr = 0.0, opt.beta2, opt.lr / 2, opt.lr * 2
    # fused=True: torch's single-kernel multi-tensor Adam (same update rule; ~7 instead of ~30 passes over the 98 M generator
    # parameters and their moments) -- the optimizer is parameter-side plumbing, as in the reference (base_model.py:39-48)
    def adam(net, lr):
        params = list(net.parameters())
        fused = FUSED_ADAM and all(p.is_cuda for p in params)
        return torch.optim.Adam(params, lr=lr, betas=(beta1, beta2), capturable=capturable, fused=True if fused else None)
    return adam(netG, g_lr), adam(netD, d_lr)


class GraphedStep:
    """The whole training iteration (train.py:58-62: D-step + G-step incl. both Adam updates) recorded once into a
    CUDA graph and replayed: ~5000 kernel launches (ours + the parameter-side torch ops) become one graph launch,
    which removes the host-side launch latency between the many small kernels.  Inputs are copied into static
    buffers before each replay; the returned losses are static tensors overwritten by each replay."""

    def __init__(self, opt, netG, netD, optG, optD, example, sync_G=None, sync_D=None, warmup=3):
        self.static = {k: v.clone() for k, v in example.items()}
        st = self.static

        def run():
            prev = [st['prev_label'], st['prev_image']] if 'prev_label' in st else None      # temporal phase inputs, if given
            return train_step(opt, netG, netD, optG, optD, st['tgt_label'], st['tgt_image'], st['ref_labels'], st['ref_images'],
                              sync_G=sync_G, sync_D=sync_D, prev=prev)
        # Capture on the stream the eager iterations already ran on (it must not be the legacy default stream): autograd's
        # AccumulateGrad nodes remember the stream they were created on, and a cross-stream wait would invalidate capture.
        cur = torch.cuda.current_stream()
        if cur == torch.cuda.default_stream():
            raise RuntimeError('GraphedStep: run the training loop under a non-default stream (torch.cuda.set_stream)')
        for _ in range(warmup):
            run()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=cur):
            self.ld, self.lg, self.fake = run()

    def __call__(self, inp):
        for k, v in inp.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.ld, self.lg, self.fake


def loss_backward(losses, optimizer, grad_sync=None):
    """loss_collector.py:217-228: mean -> sum -> zero_grad -> backward -> [allreduce] -> step."""
    loss = sum(torch.mean(v) for v in losses.values())
    optimizer.zero_grad()
    loss.backward()
    if grad_sync is not None:
        grad_sync()
    optimizer.step()
    return loss


def train_step(opt, netG, netD, optG, optD, tgt_label, tgt_image, ref_labels, ref_images, sync_G=None, sync_D=None, prev=None):
    """train.py:58-62: discriminator update, then generator update, for one frame (``prev`` = [prev_label, prev_image] in
    the temporal phase)."""
    d_losses = discriminator_losses(opt, netG, netD, tgt_label, tgt_image, ref_labels, ref_images, prev=prev)
    ld = loss_backward(d_losses, optD, sync_D)
    g_losses, fake = generator_losses(opt, netG, netD, tgt_label, tgt_image, ref_labels, ref_images, prev=prev)
    lg = loss_backward(g_losses, optG, sync_G)
    return ld, lg, fake
