"""FewShotGenerator / FlowGenerator / LabelEmbedder on the fsv_b200 kernels.

Drop-in for models/networks/generator.py of the reference: same constructor (``opt``), same
``forward(label, label_refs, img_refs, prev=[None, None], t=0, img_coarse=None)`` and 9-tuple
result (generator.py:181,229), same attribute / state_dict names, same ``init_temporal_network``
and eval-mode weight cache.  Internally activations are NHWC fp32 and every op is a call into
libfsv_b200.so; boundary tensors are NCHW like the reference (outputs are NCHW-shaped views).

In scope: ``--adaptive_spade`` (+ ``--warp_ref``, ``--spade_combine``),
``use_label_ref='mul'``, ``netS='encoderdecoder'``, ``sc_arch='unet'``, K >= 1 reference images (``n_shot``: attention
merge + picked reference, generator.py:298-316), the temporal (``warp_prev``) inputs, face / pose / street label widths.
Out-of-scope options raise NotImplementedError.

Execution structure of a forward (all capturable into one CUDA graph): one grouped spectral-norm launch triple up front
(layers.SpectralPlanner); the reference-encoder / hyper-network / label-embedding branch on a second stream concurrent with
the flow / warp / image-embedding branch (ops.branch_fork / branch_join); then the SPADE main branch.
"""
import os

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_NONE, ACT_LRELU, ACT_TANH, ACT_SIGMOID
from .layers import Conv2d, Linear, BatchNorm, ConvNormAct, SPADEResnetBlock, SpectralPlanner, spectral, init_weights

# image / label reference encoders on separate streams (FSV_ENC_SPLIT=0: one stream; -0.35 .. -0.75 ms per pose512 step, session 20)
ENC_SPLIT = os.environ.get('FSV_ENC_SPLIT', '1') != '0'


class BaseNetwork(nn.Module):
    """base_network.py:76-124 surface used from outside the networks."""

    def print_network(self):
        num_params = sum(p.numel() for p in self.parameters())
        print(self)
        print('Total number of parameters: %d' % num_params)

    def init_weights(self, init_type='normal', gain=0.02, reach_spectral=True):
        init_weights(self, init_type, gain, reach_spectral)

    def load_pretrained_net(self, net_src, net_dst):
        source, target = net_src.state_dict(), net_dst.state_dict()
        for k, v in source.items():
            if k in target and target[k].size() == v.size():
                target[k] = v
        net_dst.load_state_dict(target)


class _Act(nn.Module):
    """Placeholder that keeps nn.Sequential indices identical to the reference (activations / upsamples
    are fused into the neighbouring kernels)."""

    def forward(self, x):
        return x


def _norm_pair(conv):
    """normalization.py:54-88 get_nonspade_norm_layer('spectral(sync)batch'): spectral conv without bias + BN."""
    return nn.Sequential(spectral(conv), BatchNorm(conv.out_channels, affine=True))


class FlowGenerator(BaseNetwork):
    """generator.py:456-504."""

    def __init__(self, opt, n_frames_G):
        super().__init__()
        self.opt = opt
        input_nc = (opt.label_nc if opt.label_nc != 0 else opt.input_nc) * n_frames_G
        input_nc += opt.output_nc * (n_frames_G - 1)
        nf, nd = opt.nff, opt.n_downsample_F
        self.n_downsample_F, self.n_blocks = nd, opt.n_blocks_F
        self.flow_multiplier = opt.flow_multiplier
        ch = [min(1024, nf * (2 ** i)) for i in range(nd + 1)]
        if 'batch' not in opt.norm_F or not opt.norm_F.startswith('spectral'):
            raise NotImplementedError("only norm_F='spectral(sync)batch' is in scope")
        down = [_norm_pair(Conv2d(input_nc, nf, 3, padding=1, bias=False)), _Act()]
        for i in range(nd):
            down += [_norm_pair(Conv2d(ch[i], ch[i + 1], 3, stride=2, padding=1, bias=False)), _Act()]
        res = [SPADEResnetBlock(ch[nd], ch[nd], norm=opt.norm_F) for _ in range(opt.n_blocks_F)]
        up = []
        for i in reversed(range(nd)):
            up += [_Act(), _norm_pair(Conv2d(ch[i + 1], ch[i], 3, padding=1, bias=False)), _Act()]
        self.down_flow = nn.Sequential(*down)
        self.res_flow = nn.Sequential(*res)
        self.up_flow = nn.Sequential(*up)
        self.conv_flow = nn.Sequential(Conv2d(nf, 2, 3, padding=1))
        self.conv_mask = nn.Sequential(Conv2d(nf, 1, 3, padding=1), _Act())

    def forward(self, x):
        """x: NHWC concat [label, label_prev, img_prev] (generator.py:499).  Returns NHWC flow (2ch, pixels) and mask (1ch)."""
        for k in range(self.n_downsample_F + 1):
            conv, bn = self.down_flow[2 * k]
            x = bn(conv(x), act=ACT_LRELU)
        for blk in self.res_flow:
            x = blk(x)
        for k in range(self.n_downsample_F):
            conv, bn = self.up_flow[3 * k + 1]
            x = bn(conv(x, up=2), act=ACT_LRELU)
        flow = self.conv_flow[0](x, out_scale=float(self.flow_multiplier))
        mask = self.conv_mask[0](x, act=ACT_SIGMOID)
        return flow, mask


class LabelEmbedder(BaseNetwork):
    """generator.py:506-572 for netS in {'encoderdecoder', 'unet'}."""

    def __init__(self, opt, input_nc, netS=None, params_free_layers=0):
        super().__init__()
        self.opt = opt
        nf = opt.ngf
        self.netS = netS if netS is not None else opt.netS
        self.unet = 'unet' in self.netS
        self.decode = 'decoder' in self.netS or self.unet
        if not self.decode:
            raise NotImplementedError("LabelEmbedder without a decoder (netS='encoder') is outside the hot-path scope")
        self.n_downsample_S = nd = opt.n_downsample_G
        self.params_free_layers = params_free_layers if params_free_layers != -1 else nd
        ch = [min(1024, nf * (2 ** i)) for i in range(nd + 1)]
        self.ch = ch
        self.conv_first = nn.Sequential(Conv2d(input_nc, nf, 3, padding=1), _Act())
        for i in range(nd):
            if i >= params_free_layers or 'decoder' in self.netS:
                setattr(self, 'down_%d' % i, nn.Sequential(Conv2d(ch[i], ch[i + 1], 3, stride=2, padding=1), _Act()))
        for i in reversed(range(nd)):
            ch_i = ch[i + 1] * (2 if self.unet and i != nd - 1 else 1)
            if i >= params_free_layers:
                setattr(self, 'up_%d' % i, nn.Sequential(_Act(), Conv2d(ch_i, ch[i], 3, padding=1), _Act()))

    def forward(self, x, weights=None):
        """x NHWC; weights[i] = (flat, w_off, b_off) for the adaptive 1x1 up-convs (i < params_free_layers)."""
        if x is None:
            return None
        nd = self.n_downsample_S
        out = [self.conv_first[0](x, act=ACT_LRELU)]
        for i in range(nd):
            if i >= self.params_free_layers or self.decode:
                out.append(getattr(self, 'down_%d' % i)[0](out[-1], act=ACT_LRELU))
            else:
                raise NotImplementedError
        if not self.unet:
            out = [out[-1]]
        for i in reversed(range(nd)):
            xi = out[-1]
            if self.unet and i != nd - 1:
                xi = ops.cat_channels(xi, out[i + 1])
            if i >= self.params_free_layers:
                y = getattr(self, 'up_%d' % i)[1](xi, up=2, act=ACT_LRELU)
            else:
                # reference: Upsample -> per-sample 1x1 conv -> LReLU (generator.py:566-568).  A 1x1 conv and a
                # pointwise activation commute with nearest upsampling, so run them at low resolution (4x less
                # work, identical values) and replicate afterwards.
                flat, w_off, b_off = weights[i]
                y = ops.upsample2x(ops.batch_conv1x1(xi, flat, self.ch[i], self.ch[i + 1], w_off, b_off, act=ACT_LRELU))
            out.append(y)
        if self.unet:
            out = out[nd:]
        return out[::-1]


class FewShotGenerator(BaseNetwork):
    """generator.py:20-454."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        if getattr(opt, 'adaptive_conv', False) or getattr(opt, 'res_for_ref', False):
            raise NotImplementedError('adaptive_conv / res_for_ref are outside the hot-path scope (SURVEY.md section 8)')
        if getattr(opt, 'lambda_kld', 0) > 0:
            raise NotImplementedError('lambda_kld > 0 is outside the hot-path scope')
        self.n_shot = getattr(opt, 'n_shot', 1)          # K > 1: attention module, see attention_module
        if 'mul' not in opt.use_label_ref:
            raise NotImplementedError("only use_label_ref='mul' is in scope")
        self.n_downsample_G = nd = opt.n_downsample_G
        self.nf = nf = opt.ngf
        self.nf_max = nf_max = min(1024, nf * (2 ** nd))
        self.ch = ch = [min(nf_max, nf * (2 ** i)) for i in range(nd + 2)]
        self.norm = norm = opt.norm_G
        if opt.conv_ks != 3 or opt.embed_ks != 1 or opt.spade_ks != 1:
            raise NotImplementedError('only conv_ks=3, embed_ks=1, spade_ks=1 (the defaults) are in scope')
        self.spade_combine = opt.spade_combine
        self.n_sc_layers = opt.n_sc_layers
        self.add_raw_output_loss = opt.add_raw_output_loss and opt.spade_combine
        if self.add_raw_output_loss:
            raise NotImplementedError('add_raw_output_loss is outside the hot-path scope')
        self.ch_hidden = [[ch[i]] if not self.spade_combine or i >= self.n_sc_layers else [ch[i]] * 3 for i in range(nd + 1)]
        self.adap_spade = opt.adaptive_spade
        self.adap_embed = opt.adaptive_spade and not opt.no_adaptive_embed
        self.n_adaptive_layers = opt.n_adaptive_layers if opt.n_adaptive_layers != -1 else nd
        self.n_fc_layers = opt.n_fc_layers
        self.mul_label_ref = True
        self.use_kld = False

        norm_ref = norm.replace('spade', '')
        if 'batch' not in norm_ref:
            raise NotImplementedError("reference encoder norm must be a batch norm (norm_G='spectralspade(sync)batch')")
        input_nc = opt.label_nc if opt.label_nc != 0 else opt.input_nc
        self.ref_img_first = ConvNormAct(opt.output_nc, nf)
        self.ref_label_first = ConvNormAct(input_nc, nf)
        for i in range(nd):
            setattr(self, 'ref_img_down_%d' % i, ConvNormAct(ch[i], ch[i + 1], stride=2))
            setattr(self, 'ref_img_up_%d' % i, ConvNormAct(ch[i + 1], ch[i]))
            setattr(self, 'ref_label_down_%d' % i, ConvNormAct(ch[i], ch[i + 1], stride=2))
            setattr(self, 'ref_label_up_%d' % i, ConvNormAct(ch[i + 1], ch[i]))

        if self.adap_spade:
            for i in range(self.n_adaptive_layers):
                ch_in, ch_out = ch[i], ch[i + 1]
                ch_h = self.ch_hidden[i][0]
                names = ['fc_spade_0', 'fc_spade_1', 'fc_spade_s']
                outs = [(ch_h + 1) * 2, (ch_h + 1) * (1 if ch_in != ch_out else 2), (ch_h + 1) * 2]
                if self.adap_embed:
                    names.append('fc_spade_e')
                    outs.append(ch_in + 1)
                for name, fo in zip(names, outs):
                    layers = [spectral(Linear(ch_out, ch_out)), _Act()]
                    for _ in range(1, self.n_fc_layers):
                        layers += [spectral(Linear(ch_out, ch_out)), _Act()]
                    layers += [spectral(Linear(ch_out, fo))]
                    setattr(self, '%s_%d' % (name, i), nn.Sequential(*layers))

        self.label_embedding = LabelEmbedder(opt, input_nc, opt.netS,
                                             params_free_layers=(self.n_adaptive_layers if self.adap_embed else 0))
        for i in reversed(range(nd + 1)):
            setattr(self, 'up_%d' % i, SPADEResnetBlock(ch[i + 1], ch[i], norm=norm, hidden_nc=self.ch_hidden[i],
                                                        norm_params_free=(self.adap_spade and i < self.n_adaptive_layers)))
        self.conv_img = Conv2d(nf, 3, 3, padding=1)

        if self.n_shot > 1:                              # generator.py:127-134
            self.n_downsample_A = opt.n_downsample_A
            self.atn_query_first = ConvNormAct(input_nc, nf)
            self.atn_key_first = ConvNormAct(input_nc, nf)
            for i in range(self.n_downsample_A):
                setattr(self, 'atn_key_%d' % i, ConvNormAct(ch[i], ch[i + 1], stride=2))
                setattr(self, 'atn_query_%d' % i, ConvNormAct(ch[i], ch[i + 1], stride=2))

        self.warp_prev = False
        self.warp_ref = opt.warp_ref and not opt.for_face
        if self.warp_ref:
            self.flow_network_ref = FlowGenerator(opt, 2)
            if self.spade_combine:
                self.img_ref_embedding = LabelEmbedder(opt, opt.output_nc + 1, opt.sc_arch)

    # ------------------------------------------------------------------ temporal phase (generator.py:155-179)
    def init_temporal_network(self):
        opt = self.opt
        self.warp_prev = True
        self.sep_prev_flownet = opt.sep_flow_prev or (opt.n_frames_G != 2) or not opt.warp_ref
        self.sep_prev_embedding = self.spade_combine and (not opt.no_sep_warp_embed or not opt.warp_ref)
        dev = self.conv_img.weight.device
        # generator.py:157,179: the new sub-networks are drawn under seed 0 on every rank (one process per GPU here, so without this
        # the replicas would start from different weights); the caller's RNG streams are restored afterwards.  New parameters: the
        # optimizer must be re-created by the caller (base_model.py:267-269 does), see trainer.make_step_optimizers.
        cpu_state = torch.get_rng_state()
        cuda_state = torch.cuda.get_rng_state(dev) if dev.type == 'cuda' else None
        torch.manual_seed(0)
        try:
            self._build_temporal(opt, dev)
        finally:
            torch.set_rng_state(cpu_state)
            if cuda_state is not None:
                torch.cuda.set_rng_state(cuda_state, dev)

    def _build_temporal(self, opt, dev):
        if self.sep_prev_flownet:
            self.flow_network_temp = FlowGenerator(opt, opt.n_frames_G).to(dev)
            self.flow_network_temp.init_weights(opt.init_type, opt.init_variance)
        else:
            self.flow_network_temp = self.flow_network_ref
        if self.spade_combine:
            if self.sep_prev_embedding:
                self.img_prev_embedding = LabelEmbedder(opt, opt.output_nc + 1, opt.sc_arch).to(dev)
                self.img_prev_embedding.init_weights(opt.init_type, opt.init_variance)
            else:
                self.img_prev_embedding = self.img_ref_embedding
        self.__dict__.pop('_planner', None)          # new spectral modules: the grouped plans are re-recorded
        if self.warp_ref:
            if self.sep_prev_flownet:
                self.load_pretrained_net(self.flow_network_ref, self.flow_network_temp)
            if self.sep_prev_embedding:
                self.load_pretrained_net(self.img_ref_embedding, self.img_prev_embedding)
            self.flow_temp_is_initalized = True

    # ------------------------------------------------------------------ hyper-network (generator.py:245-273)
    def _mlp(self, name, i, x):
        seq = getattr(self, '%s_%d' % (name, i))
        for k in range(self.n_fc_layers):
            x = seq[2 * k](x, act=ACT_LRELU)
        return seq[2 * self.n_fc_layers](x)

    def get_SPADE_weights(self, feat, i):
        """feat (b, c, c) from the reference encoder.  Returns the hyper-weights as *locations* inside the MLPs'
        flat outputs, following base_network.py:132-167 exactly (SURVEY.md section 7 'hard parts'):
        embedding = (flat_e, w_off, b_off); norm = [(flat, wg_off, bg_off, wb_off, bb_off)] for conv_0, conv_1, shortcut."""
        ch_in, ch_out = self.ch[i], self.ch[i + 1]
        ch_h = self.ch_hidden[i][0]
        b, c = feat.shape[0], feat.shape[1]
        x = feat.reshape(b * c, c)
        emb = None
        if self.adap_embed:
            fc_e = self._mlp('fc_spade_e', i, x).reshape(b, -1)
            emb = (fc_e, 0, ch_in * ch_out)          # fc_e[:, :-ch_in] -> [ch_in*ch_out weights | ch_in biases]

        def gb(flat, co):
            n = co * ch_h + co
            assert flat.shape[1] == 2 * n, (flat.shape, n)
            return (flat, 0, co * ch_h, n, n + co * ch_h)
        fc_0 = self._mlp('fc_spade_0', i, x).reshape(b, -1)
        fc_1 = self._mlp('fc_spade_1', i, x).reshape(b, -1)
        fc_s = self._mlp('fc_spade_s', i, x).reshape(b, -1)
        return emb, [gb(fc_0, ch_out), gb(fc_1, ch_in), gb(fc_s, ch_out)]

    # ------------------------------------------------------------------ reference encoder (generator.py:341-393)
    def attention_encode(self, x, name):
        """generator.py:292-296."""
        x = getattr(self, name + '_first')(x)
        for i in range(self.n_downsample_A):
            x = getattr(self, '%s_%d' % (name, i))(x)
        return x

    def attention_module(self, x, label, label_ref, attention=None):
        """generator.py:298-316 on NHWC: x (b*n, h, w, c) -> (b, h, w, c).  attention is kept as (b, h, w, n*h*w) =
        the reference's (b, n*h*w, h*w) transposed, so that the softmax over the reference positions is a channel softmax
        and both GEMMs are per-sample 1x1 convs of the C ABI (no torch.bmm on the path)."""
        bn, h, w, c = x.shape
        n = self.n_shot
        b = bn // n
        if attention is None:
            key = self.attention_encode(label_ref, 'atn_key')                       # (b*n, h, w, c)
            query = self.attention_encode(label, 'atn_query')                       # (b, h, w, c)
            energy = ops.per_sample_matmul(query, key.reshape(b, n * h * w * c), n * h * w, c)
            attention = ops.softmax_channels(energy)                                # softmax over the n*h*w reference positions
        xt = x.reshape(b, n * h * w, c).transpose(1, 2).contiguous().reshape(b, c * n * h * w)
        out = ops.per_sample_matmul(attention, xt, c, n * h * w)
        atn_vis = attention.reshape(b, h * w, n, h * w).sum(3).permute(0, 2, 1).reshape(b, n, h, w)
        return out, attention, atn_vis[-1:, 0:1]

    def attention_rows_per_chunk(self, b, h, w, n):
        """Query rows per chunk of the memory-bounded attention, or None for the one-piece form.  The (b, h*w, n*h*w) attention matrix is
        what runs the reference out of memory in the inference sweep (1024x1024 with K = 5: 86 GB); without autograd nothing needs it
        whole, so it is formed a few query rows at a time.  Budget: ``attention_chunk_bytes`` (FSV_ATTN_CHUNK_MB, default 1024 MB)."""
        if torch.is_grad_enabled():
            return None
        budget = getattr(self, 'attention_chunk_bytes', None)
        if budget is None:
            budget = int(os.environ.get('FSV_ATTN_CHUNK_MB', '1024')) << 20
        if 4 * b * h * w * n * h * w <= budget:
            return None
        return max(1, min(h, budget // (4 * b * w * n * h * w)))

    def attention_chunked(self, x, xl, label, label_ref, rows):
        """generator.py:298-316,359-366 without the whole attention matrix (no-grad paths): per chunk of query rows the energies, their
        softmax over the n*h*w reference positions, both merges (image and label features) and the running per-reference attention mass
        the visualisation and the choice of the warped reference need.  Same kernels, same values as attention_module."""
        bn, h, w, c = x.shape
        n = self.n_shot
        b = bn // n
        key = self.attention_encode(label_ref, 'atn_key').reshape(b, n * h * w * c)
        query = self.attention_encode(label, 'atn_query')
        xt = x.reshape(b, n * h * w, c).transpose(1, 2).contiguous().reshape(b, c * n * h * w)
        cl = xl.shape[3]
        xlt = xl.reshape(b, n * h * w, cl).transpose(1, 2).contiguous().reshape(b, cl * n * h * w)
        out_x = torch.empty((b, h, w, c), device=x.device, dtype=x.dtype)
        out_l = torch.empty((b, h, w, cl), device=x.device, dtype=x.dtype)
        vis = torch.empty((b, n, h, w), device=x.device, dtype=x.dtype)
        for r0 in range(0, h, rows):
            r1 = min(h, r0 + rows)
            atn = ops.softmax_channels(ops.per_sample_matmul(query[:, r0:r1].contiguous(), key, n * h * w, c))   # (b, rows, w, n*h*w)
            out_x[:, r0:r1] = ops.per_sample_matmul(atn, xt, c, n * h * w)
            out_l[:, r0:r1] = ops.per_sample_matmul(atn, xlt, cl, n * h * w)
            vis[:, :, r0:r1] = atn.reshape(b, (r1 - r0) * w, n, h * w).sum(3).permute(0, 2, 1).reshape(b, n, r1 - r0, w)
        ref_idx = torch.argmax(vis.sum((2, 3)), dim=1)
        return out_x, out_l, vis[-1:, 0:1], ref_idx

    def reference_encoding(self, img_ref, label_ref, need_weights, label=None, n=1):
        nd = self.n_downsample_G
        if n == 1 and ENC_SPLIT and getattr(ops, 'BRANCH_STREAMS', False) and img_ref.is_cuda:
            # the image and the label encoder (generator.py:341-372) are independent chains of small convolutions until their feature
            # pyramids meet in the outer products: the label chain runs on a stream of its own (reduction lane 3)
            s_lab = ops.branch_fork(label_ref, index=3)
            with torch.cuda.stream(s_lab):
                xl = self.ref_label_first(label_ref)
                for i in range(nd):
                    xl = getattr(self, 'ref_label_down_%d' % i)(xl)
                enc_lab = [xl]
                if need_weights:
                    for i in reversed(range(nd)):
                        enc_lab.append(getattr(self, 'ref_label_up_%d' % i)(enc_lab[-1]))
            x = self.ref_img_first(img_ref)
            for i in range(nd):
                x = getattr(self, 'ref_img_down_%d' % i)(x)
            encoded = None
            if need_weights:
                enc_img = [x]
                for i in reversed(range(nd)):
                    enc_img.append(getattr(self, 'ref_img_up_%d' % i)(enc_img[-1]))
                ops.branch_join(s_lab, *enc_lab)
                enc_img, enc_lab = enc_img[::-1], enc_lab[::-1]
                used = set(min(nd, i + 1) for i in range(self.n_adaptive_layers)) if self.adap_spade else set()
                encoded = [ops.softmax_outer(enc_img[j], enc_lab[j]) if j in used else None for j in range(nd + 1)]
            else:
                ops.branch_join(s_lab, xl)
            return x, encoded
        x = self.ref_img_first(img_ref)
        xl = self.ref_label_first(label_ref)
        atn_vis = ref_idx = None
        for i in range(nd):
            x = getattr(self, 'ref_img_down_%d' % i)(x)
            xl = getattr(self, 'ref_label_down_%d' % i)(xl)
            rows = self.attention_rows_per_chunk(x.shape[0] // n, x.shape[1], x.shape[2], n) if (n > 1 and i == self.n_downsample_A - 1) else None
            if rows is not None:
                x, xl, atn_vis, ref_idx = self.attention_chunked(x, xl, label, label_ref, rows)
            elif n > 1 and i == self.n_downsample_A - 1:                             # generator.py:359-366
                x, atn, atn_vis = self.attention_module(x, label, label_ref)
                xl, _, _ = self.attention_module(xl, None, None, atn)
                b, h, w = atn.shape[0], atn.shape[1], atn.shape[2]
                ref_idx = torch.argmax(atn.detach().reshape(b, h * w, n, -1).sum((1, 3)), dim=1)
        encoded = None
        if need_weights:
            enc_img, enc_lab = [x], [xl]
            for i in reversed(range(nd)):
                enc_img.append(getattr(self, 'ref_img_up_%d' % i)(enc_img[-1]))
                enc_lab.append(getattr(self, 'ref_label_up_%d' % i)(enc_lab[-1]))
            enc_img, enc_lab = enc_img[::-1], enc_lab[::-1]
            # only levels 1..n_adaptive_layers are consumed (generator.py:407); the reference also forms (and
            # discards) levels 0 and nd -- a pure function of the features, skipped here with identical results.
            used = set(min(nd, i + 1) for i in range(self.n_adaptive_layers)) if self.adap_spade else set()
            encoded = [ops.softmax_outer(enc_img[j], enc_lab[j]) if j in used else None for j in range(nd + 1)]
        if n > 1:
            return x, encoded, atn_vis, ref_idx
        return x, encoded

    def weight_generation(self, img_ref, label_ref, label, t=0, n=1):
        need = self.opt.isTrain or t == 0 or n > 1                                  # generator.py:403
        atn_vis = ref_idx = None
        if n > 1:
            x, encoded, atn_vis, ref_idx = self.reference_encoding(img_ref, label_ref, need, label=label, n=n)
        else:
            x, encoded = self.reference_encoding(img_ref, label_ref, need)
        if need:
            emb_w, norm_w = [], []
            for i in range(self.n_adaptive_layers):
                if self.adap_spade:
                    e, nw = self.get_SPADE_weights(encoded[min(len(encoded) - 1, i + 1)], i)
                    emb_w.append(e)
                    norm_w.append(nw)
            if not self.opt.isTrain:
                self.embedding_weights, self.norm_weights = emb_w, norm_w    # generator.py:415-416
        else:
            emb_w, norm_w = self.embedding_weights, self.norm_weights        # generator.py:418
        enc_label = self.label_embedding(label, weights=(emb_w if self.adap_embed else None))
        if n > 1:
            return x, enc_label, norm_w, atn_vis, ref_idx
        return x, enc_label, norm_w

    # ------------------------------------------------------------------ forward (generator.py:181-229)
    def forward(self, label, label_refs, img_refs, prev=[None, None], t=0, img_coarse=None):
        if img_coarse is not None:
            raise NotImplementedError('forward_face (--refine_face) is a "next" row of SURVEY.md section 8(f)')
        planner = self.__dict__.get('_planner')
        if planner is None:
            planner = self.__dict__['_planner'] = SpectralPlanner(self)
        sig = (self.training, torch.is_grad_enabled(), prev[0] is not None, self.warp_prev, bool(self.opt.isTrain or t == 0), img_refs.shape[1])
        started = planner.begin(sig)
        try:
            return self._forward(label, label_refs, img_refs, prev, t)
        finally:
            if started:
                planner.end()

    def _forward(self, label, label_refs, img_refs, prev, t):
        b, n = img_refs.shape[0], img_refs.shape[1]
        if n != 1 and n != self.n_shot:
            raise ValueError('got %d reference images but the network was built with n_shot=%d' % (n, self.n_shot))
        nd = self.n_downsample_G
        label_n = ops.to_nhwc(label)
        atn_vis = ref_idx = None
        branch = None
        if n == 1:
            lref_pick, iref_pick = label_refs[:, 0], img_refs[:, 0]
            lref_n = ops.to_nhwc(lref_pick)
            iref_n = ops.to_nhwc(iref_pick)
            if getattr(ops, 'BRANCH_STREAMS', False) and self.warp_ref and label_n.is_cuda:
                # reference encoders -> hyper-network -> label embedding on a second stream, concurrent with the flow / warp / image
                # embedding below (independent until the main branch needs both); joined before the first SPADE block
                branch = ops.branch_fork(label_n, lref_n, iref_n)
                with torch.cuda.stream(branch):
                    x, enc_label, norm_w = self.weight_generation(iref_n, lref_n, label_n, t=t)
            else:
                x, enc_label, norm_w = self.weight_generation(iref_n, lref_n, label_n, t=t)
        else:
            # K reference images: encode all b*n, merge by attention, warp the most-attended one (generator.py:396-400,425)
            hh, ww = img_refs.shape[3], img_refs.shape[4]
            lref_all = ops.to_nhwc(label_refs.reshape(b * n, -1, hh, ww))
            iref_all = ops.to_nhwc(img_refs.reshape(b * n, -1, hh, ww))
            x, enc_label, norm_w, atn_vis, ref_idx = self.weight_generation(iref_all, lref_all, label_n, t=t, n=n)
            idx = ref_idx.view(-1, 1, 1, 1, 1)
            lref_pick = label_refs.gather(1, idx.expand(-1, 1, *label_refs.shape[2:]))[:, 0]      # base_network.py:40-47
            iref_pick = img_refs.gather(1, idx.expand(-1, 1, *img_refs.shape[2:]))[:, 0]
            iref_n = ops.to_nhwc(iref_pick)

        # ---- flow estimation + warp (generator.py:424-445)
        flow, fmask, warp, ds = [None, None], [None, None], [None, None], [None, None]
        label_prev, img_prev = prev
        has_prev = label_prev is not None
        if self.warp_ref:
            f, m = self.flow_network_ref(ops.pack_nhwc(label, lref_pick, iref_pick))
            flow[0], fmask[0] = f, m
            if self.spade_combine:
                ds[0] = ops.warp_concat(iref_n, f, m)           # [warp(3), mask(1)] in one kernel
                warp[0] = ds[0][..., :3]
            else:
                warp[0] = ops.warp_concat(iref_n, f, None)
        if self.warp_prev and has_prev:
            f, m = self.flow_network_temp(ops.pack_nhwc(label, label_prev, img_prev))
            flow[1], fmask[1] = f, m
            iprev_n = ops.to_nhwc(img_prev[:, -3:])
            if self.spade_combine:
                ds[1] = ops.warp_concat(iprev_n, f, m)
                warp[1] = ds[1][..., :3]
            else:
                warp[1] = ops.warp_concat(iprev_n, f, None)

        # ---- SPADE combine (generator.py:448-454)
        if self.spade_combine:
            emb_ref = self.img_ref_embedding(ds[0]) if ds[0] is not None else None
            emb_prev = self.img_prev_embedding(ds[1]) if ds[1] is not None else None
            for i in range(self.n_sc_layers):
                enc_label[i] = [enc_label[i], emb_ref[i] if emb_ref is not None else None,
                                emb_prev[i] if emb_prev is not None else None]

        if branch is not None:
            flats = [f[0] for nw in (norm_w or []) for f in nw] if norm_w else []
            ops.branch_join(branch, x, *[m for e in enc_label for m in (e if isinstance(e, list) else [e]) if m is not None], *flats)
        # ---- main branch (generator.py:199-207); the x2 upsample between blocks is folded into the next block
        for i in range(nd, -1, -1):
            nw = norm_w[i] if (self.adap_spade and i < self.n_adaptive_layers) else None
            x = getattr(self, 'up_%d' % i)(x, enc_label[i], norm_weights=nw, up=(1 if i == nd else 2))

        img_raw = self.conv_img(x, act=ACT_TANH, in_act=ACT_LRELU)      # tanh(conv_img(actvn(x))), generator.py:210-211

        # ---- composite (generator.py:213-227)
        if not self.spade_combine:
            if self.warp_ref:
                img_final = ops.warp_blend(iref_n, flow[0], fmask[0], img_raw)
            else:
                img_final = img_raw
                if not self.warp_prev:
                    img_raw = None
            if self.warp_prev and has_prev:
                img_final = ops.warp_blend(iprev_n, flow[1], fmask[1], img_final)
        else:
            img_final, img_raw = img_raw, None

        V = ops.nchw_view
        out_flow = [V(f) if f is not None else None for f in flow]
        out_mask = [V(m) if m is not None else None for m in fmask]
        out_warp = [V(w) if w is not None else None for w in warp]
        return (V(img_final), out_flow, out_mask, V(img_raw) if img_raw is not None else None, out_warp,
                None, None, atn_vis, ref_idx)
