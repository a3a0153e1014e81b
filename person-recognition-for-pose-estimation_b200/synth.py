"""Seeded synthetic inputs for the selective-pose glue path.

Everything is generated on the CPU with an explicit ``torch.Generator`` so the same seed gives the
same tensors here, in the tests, in the oracle runs and on the GPU box.  Nothing here touches the
CUDA kernels or the oracle; the generators only build *inputs* shaped like the three backbones'
outputs:

* detection-head raw maps ``[B, 64+nc, H_l, W_l]`` for strides 8/16/32
  (layout of ``training/yolopt/nets/nn.py:255-263`` in the reference: 4 sides x 16 DFL bins,
  side-major, followed by ``nc`` class logits),
* AdaFace embeddings ``[M, 512]`` and a gallery ``[N, 512]``
  (``libs/net_adaface.py:333-337``, ``libs/head_adaface.py:79-81``),
* frames ``[B, 3, H, W]`` and COCO-format person boxes (x, y, w, h),
* ViTPose heatmaps ``[P, K, 64, 48]`` and the heatmaps of the mirrored crop
  (``training/lightning/pose_estimation/module.py:470-476``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

STRIDES = (8, 16, 32)
DFL_BINS = 16

# COCO left/right joint pairs — training/lightning/pose_estimation/datamodule.py:25-34.
COCO_FLIP_PAIRS = ((1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16))


def flip_perm(num_joints: int, pairs: Optional[Sequence[Tuple[int, int]]] = COCO_FLIP_PAIRS) -> torch.Tensor:
    """Channel permutation equivalent to swapping every (left, right) pair; identity if ``pairs`` is None."""
    perm = list(range(num_joints))
    if pairs:
        for a, b in pairs:
            if a < num_joints and b < num_joints:
                perm[a], perm[b] = b, a
    return torch.tensor(perm, dtype=torch.int32)


def wholebody_flip_pairs(num_joints: int = 133) -> Tuple[Tuple[int, int], ...]:
    """A left/right pairing for K-joint skeletons larger than COCO-17 (synthetic: the 17 body joints
    keep the COCO pairs, the remaining joints are paired consecutively)."""
    pairs = list(COCO_FLIP_PAIRS)
    k = 17
    while k + 1 < num_joints:
        pairs.append((k, k + 1))
        k += 2
    return tuple(pairs)


def level_shapes(height: int, width: int) -> List[Tuple[int, int]]:
    """Per-level (H_l, W_l) for a letterboxed input; ``height``/``width`` must be multiples of 32."""
    assert height % 32 == 0 and width % 32 == 0, "letterbox to a multiple of 32 first (nn.py:203-209)"
    return [(height // s, width // s) for s in STRIDES]


def num_anchors(height: int, width: int) -> int:
    return sum(h * w for h, w in level_shapes(height, width))


@dataclass
class HeadMaps:
    levels: List[torch.Tensor]          # 3 x [B, 64+nc, H_l, W_l] fp32
    planted: torch.Tensor               # [B, n_obj, 5] (x1, y1, x2, y2, cls) of the planted objects
    height: int
    width: int
    nc: int


def make_head_maps(batch: int, height: int, width: int, n_obj: int = 10, nc: int = 1, seed: int = 0,
                   min_size: float = 24.0, max_size: float = 220.0, dense: bool = False) -> HeadMaps:
    """Raw detection-head maps with ``n_obj`` planted objects per frame.

    Background class logits are N(-9, 0.3) (p ~ 1e-4, below the reference's 0.001 threshold,
    ``training/yolopt/util.py:123``).  Each planted object lights the anchors whose centre falls in
    the middle of its box, on the pyramid level that can express it (every side distance < 15 grid
    units), with class logits U(-0.5, 4) and DFL logits peaked at the true side distance with a
    +-1 bin jitter, so that the decoded boxes overlap heavily and NMS has real work to do.
    ``dense=True`` makes every anchor a candidate (stress case for the suppression kernel).
    """
    g = torch.Generator().manual_seed(seed)
    shapes = level_shapes(height, width)
    no = 4 * DFL_BINS + nc
    levels = []
    for (h, w) in shapes:
        t = torch.randn(batch, no, h, w, generator=g)
        if dense:
            t[:, 4 * DFL_BINS:] = torch.rand(batch, nc, h, w, generator=g) * 6.0 - 4.0
        else:
            t[:, 4 * DFL_BINS:] = t[:, 4 * DFL_BINS:] * 0.3 - 9.0
        levels.append(t)
    planted = torch.zeros(batch, n_obj, 5)
    bins = torch.arange(DFL_BINS, dtype=torch.float32)
    for b in range(batch):
        for o in range(n_obj):
            bw = float(torch.empty(1).uniform_(min_size, max_size, generator=g))
            bh = float(torch.empty(1).uniform_(min_size, max_size, generator=g))
            cx = float(torch.empty(1).uniform_(0.1 * width, 0.9 * width, generator=g))
            cy = float(torch.empty(1).uniform_(0.1 * height, 0.9 * height, generator=g))
            cls = int(torch.randint(0, nc, (1,), generator=g))
            x1, y1, x2, y2 = cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2
            planted[b, o] = torch.tensor([x1, y1, x2, y2, float(cls)])
            # smallest stride whose 15-bin DFL range covers the box from a central anchor
            lvl = 0
            while lvl < 2 and max(bw, bh) * 0.75 / STRIDES[lvl] > 14.0:
                lvl += 1
            s = STRIDES[lvl]
            h, w = shapes[lvl]
            # anchors in the central third of the box
            ax0, ax1 = int(max(0, math.floor((cx - bw / 6) / s))), int(min(w - 1, math.ceil((cx + bw / 6) / s)))
            ay0, ay1 = int(max(0, math.floor((cy - bh / 6) / s))), int(min(h - 1, math.ceil((cy + bh / 6) / s)))
            for ay in range(ay0, ay1 + 1):
                for ax in range(ax0, ax1 + 1):
                    acx, acy = (ax + 0.5) * s, (ay + 0.5) * s
                    d = torch.tensor([(acx - x1) / s, (acy - y1) / s, (x2 - acx) / s, (y2 - acy) / s])
                    if (d < 0.2).any() or (d > 14.5).any():
                        continue
                    d = d + torch.empty(4).uniform_(-1.0, 1.0, generator=g)   # +-1 bin jitter
                    d = d.clamp(0.0, 15.0)
                    logits = -1.5 * (bins[None, :] - d[:, None]) ** 2 + 0.2 * torch.randn(4, DFL_BINS, generator=g)
                    levels[lvl][b, :4 * DFL_BINS, ay, ax] = logits.reshape(-1)
                    levels[lvl][b, 4 * DFL_BINS + cls, ay, ax] = float(torch.empty(1).uniform_(-0.5, 4.0, generator=g))
    return HeadMaps(levels=levels, planted=planted, height=height, width=width, nc=nc)


def make_head_maps_fast(batch: int, height: int, width: int, n_obj: int = 10, nc: int = 1, seed: int = 0) -> HeadMaps:
    """Same recipe as :func:`make_head_maps` but the (deterministic) planting of frame 0..7 is tiled
    across the batch with per-frame background noise — used for the full-size bench configuration,
    where a Python loop over 64 x 10 objects x anchors is needlessly slow."""
    base = make_head_maps(min(batch, 8), height, width, n_obj=n_obj, nc=nc, seed=seed)
    if batch <= 8:
        return base
    g = torch.Generator().manual_seed(seed + 7919)
    reps = (batch + 7) // 8
    levels = []
    for t in base.levels:
        big = t.repeat(reps, 1, 1, 1)[:batch].clone()
        # fresh, tiny dither on the class logits keeps scores tie-free across the tiled frames
        big[:, 4 * DFL_BINS:] += 1e-3 * torch.randn(big[:, 4 * DFL_BINS:].shape, generator=g)
        levels.append(big)
    planted = base.planted.repeat(reps, 1, 1)[:batch].clone()
    return HeadMaps(levels=levels, planted=planted, height=height, width=width, nc=nc)


@dataclass
class MatchSet:
    embeddings: torch.Tensor    # [M, 512] fp32, NOT normalised (backbone output before net_adaface.py:334)
    gallery: torch.Tensor       # [N, 512] fp32, rows unit-norm (enrolment-time normalisation)
    true_ids: torch.Tensor      # [M] int64, -1 for unknown probes


def make_match_set(m: int, n: int, dim: int = 512, unknown_frac: float = 0.1, noise: float = 0.3,
                   seed: int = 0) -> MatchSet:
    """Gallery rows = normalised N(0,1)^dim; probes = ``G[id] + noise*eps`` scaled by a random norm
    (AdaFace feature norms are O(10)), or a fresh random direction for unknown probes."""
    g = torch.Generator().manual_seed(seed)
    gal = torch.randn(n, dim, generator=g)
    gal = gal / gal.norm(dim=1, keepdim=True)
    ids = torch.randint(0, n, (m,), generator=g)
    unknown = torch.rand(m, generator=g) < unknown_frac
    if m >= 5 and unknown_frac > 0:
        unknown[m - 1] = True
    eps = torch.randn(m, dim, generator=g) / math.sqrt(dim)
    probe = gal[ids] + noise * eps
    rnd = torch.randn(m, dim, generator=g)
    probe = torch.where(unknown[:, None], rnd, probe)
    probe = probe / probe.norm(dim=1, keepdim=True)
    scale = torch.empty(m, 1).uniform_(5.0, 30.0, generator=g)
    true_ids = torch.where(unknown, torch.full_like(ids, -1), ids)
    return MatchSet(embeddings=(probe * scale).contiguous(), gallery=gal.contiguous(), true_ids=true_ids)


@dataclass
class CropSet:
    frames: torch.Tensor        # [B, 3, H, W] fp32 in [0, 1]
    boxes: torch.Tensor         # [P, 4] fp32 COCO (x, y, w, h)
    frame_idx: torch.Tensor     # [P] int32


def make_crop_set(batch: int, height: int, width: int, per_frame: int = 10, seed: int = 0,
                  smooth: bool = True) -> CropSet:
    """Frames are smooth-ish random images (low-res noise upsampled + fine noise) so that bilinear
    sampling error is representative; boxes have w in [40,200], h in [80,400] and about one in five
    crosses the frame edge (exercises the zero-outside rule of the HF warp)."""
    g = torch.Generator().manual_seed(seed)
    if smooth:
        low = torch.rand(batch, 3, max(2, height // 16), max(2, width // 16), generator=g)
        frames = torch.nn.functional.interpolate(low, size=(height, width), mode="bilinear", align_corners=False)
        frames = (0.8 * frames + 0.2 * torch.rand(batch, 3, height, width, generator=g)).contiguous()
    else:
        frames = torch.rand(batch, 3, height, width, generator=g)
    p = batch * per_frame
    w = torch.empty(p).uniform_(40.0, 200.0, generator=g)
    h = torch.empty(p).uniform_(80.0, 400.0, generator=g)
    h = torch.minimum(h, torch.tensor(float(height)) * 0.9)
    w = torch.minimum(w, torch.tensor(float(width)) * 0.9)
    x = torch.rand(p, generator=g) * (width - w)
    y = torch.rand(p, generator=g) * (height - h)
    edge = torch.rand(p, generator=g) < 0.2
    x = torch.where(edge, x - 0.6 * w * torch.sign(torch.rand(p, generator=g) - 0.5).clamp(min=0) + torch.where(
        torch.rand(p, generator=g) < 0.5, -0.3 * w, torch.zeros(p)), x)
    y = torch.where(edge & (torch.rand(p, generator=g) < 0.5), y + 0.4 * h, y)
    boxes = torch.stack([x, y, w, h], 1).contiguous()
    frame_idx = torch.arange(batch, dtype=torch.int32).repeat_interleave(per_frame).contiguous()
    return CropSet(frames=frames, boxes=boxes, frame_idx=frame_idx)


@dataclass
class HeatmapSet:
    heatmaps: torch.Tensor      # [P, K, H, W] fp32 (model output on the crop)
    flipped: torch.Tensor       # [P, K, H, W] fp32 (model output on the mirrored crop, NOT flipped back)
    centres: torch.Tensor       # [P, K, 2] planted sub-pixel (x, y)
    negative: torch.Tensor      # [P, K] bool — maps that are entirely <= 0
    perm: torch.Tensor          # [K] int32 left/right channel permutation
    boxes: torch.Tensor = field(default_factory=lambda: torch.zeros(0, 4))


def make_heatmaps(p: int, k: int = 17, height: int = 64, width: int = 48, seed: int = 0,
                  negative_frac: float = 0.05, pairs: Optional[Sequence[Tuple[int, int]]] = None,
                  noise: float = 0.02, chunk: int = 2048) -> HeatmapSet:
    """Per joint a Gaussian bump (sigma in [1,3], amplitude U(0.3,1), random sub-pixel centre, some
    centres within 2 px of the border) plus N(0, noise) noise.  The "flipped" twin is what the
    model would emit for the mirrored crop: the bump mirrored in x on the pair-swapped channel with
    independent noise.  ``negative_frac`` of the maps are made entirely negative (score <= 0 branch of
    HF ``get_keypoint_predictions``, image_processing_vitpose.py:203-204)."""
    g = torch.Generator().manual_seed(seed)
    if pairs is None:
        pairs = COCO_FLIP_PAIRS if k == 17 else wholebody_flip_pairs(k)
    perm = flip_perm(k, pairs)
    cx = torch.rand(p, k, generator=g) * (width - 1)
    cy = torch.rand(p, k, generator=g) * (height - 1)
    sig = torch.empty(p, k).uniform_(1.0, 3.0, generator=g)
    amp = torch.empty(p, k).uniform_(0.3, 1.0, generator=g)
    neg = torch.rand(p, k, generator=g) < negative_frac
    ys = torch.arange(height, dtype=torch.float32)[:, None]
    xs = torch.arange(width, dtype=torch.float32)[None, :]
    hm = torch.empty(p, k, height, width)
    fl = torch.empty(p, k, height, width)
    permL = perm.long()
    for s in range(0, p, chunk):
        e = min(p, s + chunk)
        c_x, c_y = cx[s:e, :, None, None], cy[s:e, :, None, None]
        sg, am = sig[s:e, :, None, None], amp[s:e, :, None, None]
        bump = am * torch.exp(-((xs - c_x) ** 2 + (ys - c_y) ** 2) / (2 * sg * sg))
        a = bump + noise * torch.randn(bump.shape, generator=g)
        # model on mirrored crop: channel j shows joint perm[j] mirrored in x
        b = bump[:, permL].flip(-1) + noise * torch.randn(bump.shape, generator=g)
        ng = neg[s:e, :, None, None]
        a = torch.where(ng, -a.abs() - 1e-3, a)
        b = torch.where(ng[:, permL], -b.abs() - 1e-3, b)
        hm[s:e], fl[s:e] = a, b
    return HeatmapSet(heatmaps=hm, flipped=fl, centres=torch.stack([cx, cy], -1), negative=neg, perm=perm)
