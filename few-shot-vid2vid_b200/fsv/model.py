"""One training iteration of the hot path for all three dataset geometries of BASELINE.json: the mirror of
models/vid2vid_model.py:62-176 (forward_generator / forward_discriminator / generate_images), models/loss_collector.py:47-215,
models/face_refiner.py:24-83 and models/input_process.py:25-113 on the fsv drop-in networks, WITHOUT host synchronisation:

  * face (1-ch edge labels), pose (6-ch DensePose+OpenPose labels: foreground masks in the D input (20 ch), body-part / fg
    warp terms, face-weighted mask terms, ``--add_face_D`` face discriminator on device-cropped regions,
    ``--remove_face_labels``), street (20-class one-hot labels, D input 46 ch);
  * single-frame and temporal phase (``prevs`` ring buffer, concat_prev; netDT terms when ``lambda_temp > 0``);
  * K reference images (the picked reference conditions D, vid2vid_model.py:145).

The reference computes the face box with nonzero() + four .item() per sample (face_refiner.py:57-80) and builds masks with
python loops; here boxes and masks come from fsv kernels and stay on the device, so the whole iteration (both optimizer
steps included) can be recorded into ONE CUDA graph (trainer.GraphedStep).  Flags: --no_flow_gt (F_Flow = 0) and
--no_vgg_loss (G_VGG = 0; the face term of loss_collector.py:82 is then 0 as well, see baseline/refenv.py).

Loss values are returned in the reference's order and naming (loss_collector.py:41-44).  The arithmetic on top of the
network outputs is elementwise torch on device tensors; the reductions that matter for time are listed in DESIGN.md.
"""
import torch
import torch.nn.functional as F

from . import ops
from .networks import define_G, define_D
from .networks.vgg import VGGLoss

LOSS_NAMES_G = ['G_GAN', 'G_GAN_Feat', 'G_VGG', 'Gf_GAN', 'Gf_GAN_feat', 'GT_GAN', 'GT_GAN_Feat', 'F_Flow', 'F_Warp', 'F_Mask']
LOSS_NAMES_D = ['D_real', 'D_fake', 'Df_real', 'Df_fake', 'DT_real', 'DT_fake']


# ------------------------------------------------------------------------------------------------ input processing
def encode_label(opt, label_map):
    """input_process.py:25-45: identity for label_nc == 0, else one-hot over label_nc classes (scatter_ on a zero tensor)."""
    if opt.label_nc == 0:
        return label_map
    size = label_map.shape
    lm = label_map.reshape(-1, *size[-3:]) if len(size) == 5 else label_map
    one_hot = torch.zeros((lm.shape[0], opt.label_nc) + tuple(lm.shape[2:]), device=lm.device, dtype=torch.float32)
    one_hot.scatter_(1, lm.long(), 1.0)
    return one_hot.view(size[0], size[1], -1, *size[-2:]) if len(size) == 5 else one_hot


def _face_mask(part):
    """input_process.py:81-93 get_face_mask on a (..., H, W) DensePose part plane."""
    p = (part / 2 + 0.5) * 24
    return (((p > 23 - 0.1) & (p < 23 + 0.1)) | ((p > 24 - 0.1) & (p < 24 + 0.1))).float()


def use_valid_labels(opt, pose):
    """input_process.py:95-113."""
    if pose is None or 'pose' not in opt.dataset_mode:
        return pose
    if getattr(opt, 'pose_type', 'both') == 'open':
        raise NotImplementedError("pose_type='open' (the reference README: 'only both is supported now')")
    if getattr(opt, 'remove_face_labels', False):
        cd = pose.dim() - 3
        dp, op = pose.narrow(cd, 0, 3), pose.narrow(cd, 3, pose.shape[cd] - 3)
        fm = _face_mask(pose.select(cd, 2)).unsqueeze(cd)
        return torch.cat([dp * (1 - fm) - fm, op], dim=cd)
    return pose


def pick_ref(refs, ref_idx):
    """base_network.py:40-47."""
    if isinstance(refs, (list, tuple)):
        return [pick_ref(r, ref_idx) for r in refs]
    if ref_idx is None:
        return refs[:, 0]
    idx = ref_idx.long().view(-1, 1, 1, 1, 1)
    return refs.gather(1, idx.expand(-1, 1, *refs.shape[2:]))[:, 0]


# ------------------------------------------------------------------------------------------------ loss arithmetic
def _hinge(pred, target_is_real):
    """loss.py:69-78, for_discriminator=True -- also what the reference's generator term uses (loss_collector.py:66 omits
    for_discriminator=False)."""
    z = pred * 0
    return -torch.mean(torch.min(pred - 1, z)) if target_is_real else -torch.mean(torch.min(-pred - 1, z))


def _gan_loss(preds, target_is_real):
    """loss.py:92-104."""
    loss = 0
    for p in preds:
        loss = loss + _hinge(p[-1], target_is_real).view(1)
    return loss / len(preds)


def _split(pred):
    """base_model.py:141-147 divide_pred."""
    fake = [[t[:t.size(0) // 2] for t in p] for p in pred]
    real = [[t[t.size(0) // 2:] for t in p] for p in pred]
    return fake, real


class _frozen:
    """The generator update needs gradients THROUGH the discriminators, not FOR them: with their parameters frozen during that
    forward the backward pass skips every discriminator weight gradient (the reference computes and then discards them:
    optimizer_D.zero_grad() of the next iteration, loss_collector.py:217-228)."""

    def __init__(self, modules):
        self.params = [p for m in modules for p in m.parameters() if p.requires_grad]

    def __enter__(self):
        for p in self.params:
            p.requires_grad_(False)

    def __exit__(self, *exc):
        for p in self.params:
            p.requires_grad_(True)


class Vid2VidStep:
    """Networks + per-iteration loss graph of one rank.  ``batch`` (reference layout, data/fewshot_*_dataset.py):
    tgt_label (B,1,C,H,W), tgt_image (B,1,3,H,W), ref_label (B,K,C,H,W), ref_image (B,K,3,H,W) and, in the temporal phase,
    prev_label / prev_real / prev_fake (B, n_frames_G-1, C|3, H, W)."""

    def __init__(self, opt, netG=None, netD=None, netDf=None, netDT=None, vgg_loss=None):
        self.opt = opt
        self.pose = 'pose' in opt.dataset_mode
        self.has_fg = self.pose
        self.add_face_D = bool(getattr(opt, 'add_face_D', False))
        if getattr(opt, 'refine_face', False):
            raise NotImplementedError('--refine_face (netGf) is outside the hot-path scope (SURVEY.md section 8f rank 3, second half)')
        if getattr(opt, 'n_frames_per_gpu', 1) != 1:
            raise NotImplementedError('n_frames_per_gpu != 1 (the reference: "only 1 is supported now")')
        self.tD = 1
        gpu_ids = list(getattr(opt, 'gpu_ids', []))
        # base_model.py:167-197 define_networks
        input_nc = opt.label_nc if (opt.label_nc != 0 and not self.pose) else opt.input_nc
        opt.for_face = False
        self.netG = netG if netG is not None else define_G(opt)
        nc_d = (input_nc + opt.output_nc + (1 if self.has_fg else 0)) * 2           # concat_ref_for_D (netD_subarch 'n_layers')
        self.netD = netD if netD is not None else define_D(opt, nc_d, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, opt.num_D,
                                                           not opt.no_ganFeat_loss, gpu_ids=gpu_ids)
        self.netDf = netDf
        if self.add_face_D and netDf is None:
            self.netDf = define_D(opt, opt.output_nc * 2, opt.ndf, opt.n_layers_D, opt.norm_D, 'n_layers', 1, not opt.no_ganFeat_loss, gpu_ids=gpu_ids)
        self.netDT = netDT
        # loss.py:107-128 perceptual loss (absent under --no_vgg_loss): a frozen VGG19 feature stack on the same conv kernels
        self.vgg_loss = vgg_loss
        if self.vgg_loss is None and not opt.no_vgg_loss:
            self.vgg_loss = VGGLoss()
            if gpu_ids:
                self.vgg_loss.cuda()
        self.temporal = netDT is not None or getattr(self.netG, 'warp_prev', False)
        self.face_size = int(opt.fineSize / opt.aspect_ratio) // 4                  # face_refiner.py:22

    # -------------------------------------------------------------------------------------------- temporal phase
    def init_temporal_model(self):
        """base_model.py:259-279: temporal generator branches + the temporal discriminator.  Optimizers must be re-created by
        the caller afterwards (new parameters), as the reference does."""
        opt = self.opt
        self.temporal = True
        self.netG.init_temporal_network()
        self.tD = min(opt.n_frames_D, opt.n_frames_G)
        self.netDT = define_D(opt, opt.output_nc * self.tD, opt.ndf, opt.n_layers_D, opt.norm_D, 'n_layers', 1, not opt.no_ganFeat_loss,
                              gpu_ids=list(getattr(opt, 'gpu_ids', [])))
        return self.netDT

    def d_modules(self):
        return [m for m in (self.netD, self.netDT, self.netDf) if m is not None]

    def d_parameters(self):
        """base_model.py:208-211,275-277: netD [+ netDT] [+ netDf]."""
        return [p for m in self.d_modules() for p in m.parameters()]

    def concat_prev(self, prev, now):
        """vid2vid_model.py:169-176."""
        if isinstance(prev, (list, tuple)):
            return [self.concat_prev(p, n) for p, n in zip(prev, now)]
        if prev is None:
            prev = now.unsqueeze(1).repeat(1, self.opt.n_frames_G - 1, 1, 1, 1)
        else:
            prev = torch.cat([prev[:, 1:], now.unsqueeze(1)], dim=1)
        return prev.detach()

    # -------------------------------------------------------------------------------------------- per-iteration constants
    def prepare(self, batch):
        """Everything that does not depend on the generator's output, computed once per iteration and shared by the D-step
        and the G-step: encoded / valid labels, foreground / body-part / face masks, face boxes, the constant part of the
        discriminator inputs."""
        opt = self.opt
        c = {}
        tgt_labels = encode_label(opt, batch['tgt_label'])
        ref_labels = encode_label(opt, batch['ref_label'])
        c['ref_labels'], c['ref_images'] = ref_labels, batch['ref_image']
        c['ref_labels_valid'] = use_valid_labels(opt, ref_labels)
        c['tgt_label'] = tgt_labels[:, 0]
        c['tgt_image'] = batch['tgt_image'][:, 0]
        c['tgt_label_valid'] = use_valid_labels(opt, c['tgt_label'])
        prevs = [batch.get('prev_label'), batch.get('prev_real'), batch.get('prev_fake')]
        if prevs[0] is not None:
            prevs[0] = encode_label(opt, prevs[0])
        c['prevs'] = prevs
        b, _, h, w = c['tgt_label'].shape
        c['prev_t'] = [p.contiguous().view(b, -1, h, w) if p is not None else None for p in (prevs[0], prevs[2])]
        if self.has_fg:
            c['fg_mask'] = ops.fg_mask(c['tgt_label'])                       # generate_images: from tgt_label_t (vid2vid_model.py:150)
            if getattr(opt, 'remove_face_labels', False):
                c['part_t'] = None
        return c

    def _ref_constants(self, c, ref_idx):
        """Constants that depend on which reference was picked (K > 1: known only after the generator ran)."""
        opt = self.opt
        if ref_idx is None and 'r' in c:
            return c['r']          # one reference image: the D-step and the G-step of an iteration share masks, boxes, packed inputs, crops
        ref_label_valid, ref_label_t, ref_image_t = pick_ref([c['ref_labels_valid'], c['ref_labels'], c['ref_images']], ref_idx)
        r = dict(ref_label_valid=ref_label_valid, ref_image=ref_image_t)
        if ref_idx is None:
            c['r'] = r
        if self.has_fg:
            r['ref_fg_mask'] = ops.fg_mask(ref_label_t)                        # generate_images (vid2vid_model.py:150)
            r['fg_union'] = ((c['fg_mask'] > 0) | (r['ref_fg_mask'] > 0)).float()
            # compute_GAN_losses / compute_flow_losses recompute both masks from tgt_label and the VALID reference label
            # (loss_collector.py:107, 148): identical unless --remove_face_labels changed channel 2 of the reference label
            r['ref_fg_mask_v'] = ops.fg_mask(ref_label_valid) if getattr(opt, 'remove_face_labels', False) else r['ref_fg_mask']
        return r

    def _d_base(self, c, r):
        """Constant rows/channels of netD's input: [ref_label(+fg) | ref_image | tgt_label(+fg) | image], image = real for rows [B, 2B)."""
        b, _, h, w = c['tgt_label'].shape
        label = use_valid_labels(self.opt, c['tgt_label'])
        refs = [r['ref_label_valid']] + ([r['ref_fg_mask_v']] if self.has_fg else []) + [r['ref_image']]
        tgts = [label] + ([c['fg_mask']] if self.has_fg else [])
        nc = sum(t.shape[1] for t in refs + tgts) + 3
        cp = ops.pad_channels(nc)
        base = (torch.zeros if cp != nc else torch.empty)((2 * b, h, w, cp), device=label.device, dtype=torch.float32)
        for row0 in (0, b):
            coff = ops.pack_rows(base, row0, refs + tgts, 0)
        ops.pack_rows(base, b, [c['tgt_image']], coff)
        return base, coff

    # -------------------------------------------------------------------------------------------- generator call
    def generate(self, c):
        """vid2vid_model.py:130-158 generate_images for one frame."""
        out = self.netG(c['tgt_label_valid'], c['ref_labels_valid'], c['ref_images'], c['prev_t'])
        fake, flow, fmask, fake_raw, warp = out[0], out[1], out[2], out[3], out[4]
        r = self._ref_constants(c, out[8])
        if fake_raw is not None:
            raise NotImplementedError('fake_raw_image (warp_ref without spade_combine / add_raw_output_loss) is outside the scope of this step')
        prevs_new = self.concat_prev(c['prevs'], [c['tgt_label_valid'], c['tgt_image'], fake])
        return fake, flow, fmask, warp, r, prevs_new, out[7]

    # -------------------------------------------------------------------------------------------- GAN terms
    def _discriminate(self, c, r, fake, for_discriminator):
        """loss_collector.py:47-68 on the packed input."""
        base, coff = self._d_base(c, r) if 'd_base' not in r else r['d_base']
        r['d_base'] = (base, coff)
        pred = self.netD.forward_nhwc(ops.d_input(base, fake, coff))
        pf, pr = _split(pred)
        if for_discriminator:
            return [_gan_loss(pr, True), _gan_loss(pf, False)]
        return [_gan_loss(pf, True), self._gan_feat(pred)]

    def _gan_feat(self, pred):
        """loss_collector.py:206-215 on the UNSPLIT predictions (batch [fake ; real]): one fused L1 per feature map."""
        z = pred[0][0].new_zeros(1)
        if self.opt.no_ganFeat_loss:
            return z
        num_d = len(pred)
        loss = z
        for p in pred:
            for f in p[:-1]:
                loss = loss + ops.halves_l1(f.permute(0, 2, 3, 1)) / num_d
        return loss * self.opt.lambda_feat

    def _face_boxes(self, c, r):
        """face_refiner.py:52-83 via fsv_face_bbox.  Target box from tgt_label; reference box from the reference label AS
        loss_collector.py:108-119 passes it: the valid label with the foreground mask appended, so that in OpenPose mode the
        'last three channels' are [label[-2], label[-1], fg mask]."""
        if 'boxes' in r:
            return r['boxes']
        opt = self.opt
        openpose = not getattr(opt, 'basic_point_only', False) and not getattr(opt, 'remove_face_labels', False)
        tl, rl = c['tgt_label'], r['ref_label_valid']
        if openpose:
            tb = ops.face_bbox([(tl, -3), (tl, -2), (tl, -1)], 0.0, True)
            rb = ops.face_bbox([(rl, -2), (rl, -1), (r['ref_fg_mask_v'], 0)], 0.0, True)
        else:
            tb = ops.face_bbox([(tl, 2)], 0.9, False)
            rb = ops.face_bbox([(rl, 2)], 0.9, False)
        r['boxes'] = (tb, rb)
        return r['boxes']

    def _discriminate_face(self, c, r, fake, for_discriminator):
        """loss_collector.py:70-85."""
        z = fake.new_zeros(1)
        if not self.add_face_D:
            return [z, z]
        tb, rb = self._face_boxes(c, r)
        S = self.face_size
        if 'face_static' not in r:
            r['face_static'] = (ops.crop_resize(c['tgt_image'], tb, S), ops.crop_resize(r['ref_image'], rb, S))
        real_region, ref_region = r['face_static']
        fake_region = ops.crop_resize(fake, tb, S)                                   # NHWC (B, S, S, 3)
        x = ops.cat_channels(torch.cat([ref_region, ref_region], 0), torch.cat([fake_region, real_region], 0))
        pred_f = self.netDf.forward_nhwc(x)
        pf, pr = _split(pred_f)
        lam = self.opt.lambda_face
        if for_discriminator:
            return [_gan_loss(pr, True) * lam, _gan_loss(pf, False) * lam]
        g_gan, g_feat = _gan_loss(pf, True) * lam, self._gan_feat(pred_f) * lam
        g_feat = g_feat + F.l1_loss(fake_region, real_region) * self.opt.lambda_feat
        if self.vgg_loss is not None:                                                     # loss_collector.py:82
            g_feat = g_feat + self.vgg_loss(fake_region.permute(0, 3, 1, 2), real_region.permute(0, 3, 1, 2)) * self.opt.lambda_vgg
        return [g_gan, g_feat]

    def _temporal_gan(self, c, fake, for_discriminator):
        """vid2vid_model.py:69-75,112-118 + loss_collector.py:86-116 (for_temporal): netDT on tD consecutive frames."""
        z = fake.new_zeros(1)
        prevs = c['prevs']
        if not (self.opt.lambda_temp > 0 and prevs[0] is not None) or self.tD < 2:
            return None if for_discriminator else [z, z]
        real_all = torch.cat([prevs[1], c['tgt_image'].unsqueeze(1)], dim=1)
        fake_all = torch.cat([prevs[2], fake.unsqueeze(1)], dim=1)
        bs, t, ch, h, w = real_all.shape
        if t != self.tD:
            raise NotImplementedError('temporal discriminator with n_frames_G != n_frames_D')
        x = torch.cat([fake_all.reshape(bs, ch * t, h, w), real_all.reshape(bs, ch * t, h, w)], dim=0)
        pred_t = self.netDT(x)
        pf, pr = _split(pred_t)
        if for_discriminator:
            return [_gan_loss(pr, True), _gan_loss(pf, False)]
        return [_gan_loss(pf, True) * self.opt.lambda_temp, self._gan_feat(pred_t) * self.opt.lambda_temp]

    # -------------------------------------------------------------------------------------------- the two steps
    def discriminator_losses(self, batch, c=None, gen=None):
        """vid2vid_model.py:106-128.  -> dict in LOSS_NAMES_D order (absent temporal terms omitted, like the reference's list).
        ``gen``: the result of ``generate(c)`` under no_grad when the caller already ran it (trainer: on another stream)."""
        c = self.prepare(batch) if c is None else c
        if gen is None:
            with torch.no_grad():
                gen = self.generate(c)
        fake, r = gen[0], gen[4]
        fake = fake.detach()
        d = self._discriminate(c, r, fake, True) + self._discriminate_face(c, r, fake, True)
        t = self._temporal_gan(c, fake, True)
        names = LOSS_NAMES_D[:4] + (LOSS_NAMES_D[4:] if t else [])
        return dict(zip(names, d + (t or [])))

    def generator_losses(self, batch, c=None, gen=None):
        """vid2vid_model.py:62-104.  -> (dict in LOSS_NAMES_G order, fake image, prevs_new).  ``gen``: a ``generate(c)`` result (with grad)."""
        opt = self.opt
        c = self.prepare(batch) if c is None else c
        fake, flow, fmask, warp, r, prevs_new, _ = self.generate(c) if gen is None else gen
        z = fake.new_zeros(1)
        with _frozen(self.d_modules()):
            gt = self._temporal_gan(c, fake, False)
            g = self._discriminate(c, r, fake, False) + self._discriminate_face(c, r, fake, False)
        tgt = c['tgt_image']

        # ---- flow + mask losses (loss_collector.py:131-204; F_Flow = 0 under --no_flow_gt): ONE fused pass forward, one backward
        # (ops.flow_mask_losses).  Frame tensors are handed over as NHWC (the generator's native layout; its NCHW outputs are views).
        nhwc = lambda t: None if t is None else t.permute(0, 2, 3, 1)   # noqa: E731
        pose_terms = self.pose and flow[0] is not None
        rbw = body = rfw = fg = None
        if pose_terms:
            flow_ref = nhwc(flow[0])
            body = ops.part_masks(c['tgt_label'])                                   # (B, H, W, 9)
            rbw = ops.warp_concat(ops.part_masks(r['ref_label_valid']), flow_ref, None)   # resample(ref_body_mask, flow) (:144-145)
            if self.has_fg:
                fg = nhwc(c['fg_mask'])
                rfw = ops.warp_concat(nhwc(r['ref_fg_mask_v']), flow_ref, None)            # (:148-150)
        face_avg = fg_diff = fake_t = None
        if self.pose and getattr(self.netG, 'warp_ref', False) and fmask[0] is not None:
            face_avg = nhwc(ops.face_mask_avg15(c['tgt_label']))                    # (:176-179)
            fg_diff = nhwc(((r['ref_fg_mask'] - c['fg_mask']) > 0).float())        # the masks of generate_images (vid2vid_model.py:95-96)
            fake_t = nhwc(fake) if opt.spade_combine else None
        if flow[0] is None and flow[1] is None:
            f_warp = f_mask = z
        else:
            fm = ops.flow_mask_losses(nhwc(warp[0]), nhwc(fmask[0]), nhwc(warp[1]), nhwc(fmask[1]), tgt, fake_t, rbw, body, rfw, fg, face_avg, fg_diff)
            f_warp = fm[0:1] * opt.lambda_flow
            f_mask = fm[1:2] * opt.lambda_mask

        g_vgg = z if self.vgg_loss is None else z + self.vgg_loss(fake, tgt) * opt.lambda_vgg        # loss_collector.py:122-129
        vals = [g[0], g[1], g_vgg, g[2], g[3], gt[0], gt[1], z, f_warp, f_mask]
        return dict(zip(LOSS_NAMES_G, vals)), fake, prevs_new
