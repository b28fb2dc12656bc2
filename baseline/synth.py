"""Synthetic inputs of SURVEY.md section 8(d) for the three dataset geometries (face / pose / street).

Pure torch, no fsv and no reference imports: shared by both arms of bench.py, the tests and the golden generator so
that every party sees identical tensors for a given (kind, batch, H, W, seed).

Returned dict (CPU fp32 tensors, reference layout -- data/fewshot_*_dataset.py):
    tgt_label  (B, 1, C, H, W)      tgt_image  (B, 1, 3, H, W)
    ref_label  (B, K, C, H, W)      ref_image  (B, K, 3, H, W)
    [prev_label (B, 1, C, H, W), prev_real (B, 1, 3, H, W), prev_fake (B, 1, 3, H, W)]   with temporal=True
C = 1 (face edge map in {0,1}), 6 (pose: DensePose IUV + OpenPose RGB in [-1,1], background exactly -1) or
1 (street: integer class ids 0..19 as float, one-hot encoded later by the model's encode_label).
"""
import torch
import torch.nn.functional as F


def _edges(g, H, W, *shape, p=0.03):
    e = (torch.rand(*shape, H, W, generator=g) < p).float()
    return F.max_pool2d(e.view(-1, 1, H, W), 3, 1, 1).view(*shape, H, W)


def _pose_label(g, n, H, W):
    """fewshot_pose_dataset.py:143-190 look-alike: ch0-1 DensePose UV inside a person box else -1; ch2 part id k/24*2-1 in 24
    horizontal bands of the box (so get_part_mask / get_face_mask's +-0.1 integer tests hit, input_process.py:73-93) else -1;
    ch3-5 OpenPose strokes on a -1 background, plus one 12x12 all-ones head blob per sample (face_refiner.py:58-60: all three
    OpenPose channels > 0 marks the face).  Stroke pixels never have ch3 > 0, so the blob alone defines the target's face box."""
    lab = -torch.ones(n, 6, H, W)
    bh, bw = int(0.6 * H), int(0.4 * W)
    for i in range(n):
        jy = int(torch.randint(-H // 16, H // 16 + 1, (1,), generator=g))
        jx = int(torch.randint(-W // 8, W // 8 + 1, (1,), generator=g))
        y0, x0 = (H - bh) // 2 + jy, (W - bw) // 2 + jx
        lab[i, 0:2, y0:y0 + bh, x0:x0 + bw] = torch.rand(2, bh, bw, generator=g) * 2 - 1
        band = (torch.arange(bh) * 24 // bh).flip(0) + 1          # k = 24 (head) at the top ... 1 at the bottom
        lab[i, 2, y0:y0 + bh, x0:x0 + bw] = (band.float() / 24 * 2 - 1).view(bh, 1).expand(bh, bw)
        stroke = torch.rand(H, W, generator=g) < 0.01
        vals = torch.rand(3, H, W, generator=g) * 2 - 1
        vals[0] = -vals[0].abs()
        lab[i, 3:6] = torch.where(stroke, vals, lab[i, 3:6])
        hy, hx = y0 + bh // 32, x0 + bw // 2 - 6
        lab[i, 3:6, hy:hy + 12, hx:hx + 12] = 1.0
    return lab


def _street_label(g, n, H, W, classes=20, block=16):
    ids = torch.randint(0, classes, (n, 1, (H + block - 1) // block, (W + block - 1) // block), generator=g).float()
    return F.interpolate(ids, scale_factor=block, mode='nearest')[:, :, :H, :W].contiguous()


def make(kind, batch, H, W, seed, K=1, temporal=False):
    g = torch.Generator().manual_seed(seed)
    img = lambda *s: torch.rand(*s, 3, H, W, generator=g) * 2 - 1   # noqa: E731
    if kind == 'face':
        lab = lambda n: _edges(g, H, W, n, 1)                        # noqa: E731
    elif kind == 'pose':
        lab = lambda n: _pose_label(g, n, H, W)                      # noqa: E731
    elif kind == 'street':
        lab = lambda n: _street_label(g, n, H, W)                    # noqa: E731
    else:
        raise ValueError(kind)
    C = 6 if kind == 'pose' else 1
    d = dict(tgt_label=lab(batch).view(batch, 1, C, H, W), tgt_image=img(batch, 1),
             ref_label=lab(batch * K).view(batch, K, C, H, W), ref_image=img(batch, K))
    if temporal:
        d.update(prev_label=lab(batch).view(batch, 1, C, H, W), prev_real=img(batch, 1), prev_fake=img(batch, 1))
    return d
