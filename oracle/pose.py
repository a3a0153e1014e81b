"""Oracle (TEST INFRASTRUCTURE): ViTPose heatmap decode, SURVEY.md §8a rows a11-a15.

Three decode variants exist around the reference:
* soft-argmax — the live Lightning module, training/lightning/pose_estimation/module.py:237-296
  (+ back-projection loop :534-546) — runnable reference, pinned by tests/golden/pose_live.npz;
* argmax + DARK/UDP — HF ``VitPoseImageProcessor.post_process_pose_estimation`` (third-party,
  un-vendored, ``transformers>=4.48.1``; ``HF:`` = transformers/models/vitpose/image_processing_vitpose.py,
  5.5.0 installed) — compared live against the installed HF code in the CPU tests;
* argmax + quarter-offset — gluoncv ``get_max_pred``/``get_final_preds`` called at
  training/lightning/pose_estimation/module_v2.py:214-222 — gluoncv is not installed/pinned:
  PARITY UNPINNED, restated from the published Simple-Baselines/HRNet formulas.

Reference quirk Q6: HF's DARK addresses its taps through a float32 flat index (HF:248-250), exact only for the
first 2^24 / ((W+2)(H+2)) maps of a call.  ``dark_refine_full`` / ``hf_dark_decode`` restate HF literally (same numpy
in-place float32 add, so they carry the quirk); ``dark_decode_local`` is the per-joint restatement with the taps at
their true positions.  The two are bit-identical below the limit (tests/test_oracle_golden.py).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .crop import box_to_center_and_scale, center_scale_v2


# ----------------------------------------------------------------------------------------------
# a11 flip test
# ----------------------------------------------------------------------------------------------

def flip_back(flipped: torch.Tensor, perm: Optional[torch.Tensor]) -> torch.Tensor:
    """Correct flip-back: swap left/right channels, then mirror the width axis —
    ``module copy.py:465-472`` and HF modeling_vitpose.py:80-117 (``flip_back``).  ``perm`` is the
    channel permutation of the pairs; ``None`` reproduces module_v2.py:201 (quirk Q2: no swap)."""
    x = flipped if perm is None else flipped[:, perm.long()]
    return x.flip(-1)


def flip_average(hm: torch.Tensor, flipped: torch.Tensor, perm: Optional[torch.Tensor]) -> torch.Tensor:
    """module.py:484 / module_v2.py:204 — ``(hm + flipped_back) * 0.5`` in fp32."""
    return (hm + flip_back(flipped, perm)) * 0.5


def flip_back_quirk_q1(flipped: torch.Tensor, pairs) -> torch.Tensor:
    """The live module's version, module.py:479-481: after the width flip it executes
    ``hm[:, pair] = hm[:, pair].flip(0)`` which reverses the BATCH axis of those channels instead of
    swapping them.  Kept only to document the bug; identity on the pairs at B=1."""
    x = flipped.flip(-1).clone()
    for pair in pairs:
        x[:, list(pair)] = x[:, list(pair)].flip(0)
    return x


# ----------------------------------------------------------------------------------------------
# a12 + a15 soft-argmax (live module)
# ----------------------------------------------------------------------------------------------

def soft_argmax_decode(heatmaps: torch.Tensor, boxes: Optional[torch.Tensor] = None):
    """module.py:237-296 — softmax over the flattened map, expected column/row + 0.5, normalised by
    W/H; score = max probability, optionally scaled by clamp(sqrt(box area)/96, 0.5, 2)."""
    b, k, h, w = heatmaps.shape
    yg, xg = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                            indexing="ij")
    prob = F.softmax(heatmaps.reshape(b, k, -1), dim=2).reshape(b, k, h, w)
    x = (prob * xg[None, None]).sum(dim=(2, 3)) + 0.5
    y = (prob * yg[None, None]).sum(dim=(2, 3)) + 0.5
    scores = prob.reshape(b, k, -1).max(dim=2)[0]
    coords = torch.stack([x / w, y / h], dim=-1)
    if boxes is not None:
        area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
        wgt = torch.clamp(torch.sqrt(area).view(-1, 1, 1) / 96.0, min=0.5, max=2.0)
        scores = scores * wgt.squeeze(-1)
    return coords, scores


def backproject_live(coords: torch.Tensor, scores: torch.Tensor, boxes_xyxy: torch.Tensor,
                     keypoint_thresh: float = 0.3) -> torch.Tensor:
    """module.py:534-546 — ``x = kx*(x2-x1) + x1``, ``y = ky*(y2-y1) + y1``, ``v = 2 if score > thr else 1``.
    Returns ``[P, K, 3]`` fp32 (the reference builds Python floats; fp32 arithmetic on the tensors)."""
    bw = (boxes_xyxy[:, 2] - boxes_xyxy[:, 0])[:, None]
    bh = (boxes_xyxy[:, 3] - boxes_xyxy[:, 1])[:, None]
    x = coords[..., 0] * bw + boxes_xyxy[:, 0:1]
    y = coords[..., 1] * bh + boxes_xyxy[:, 1:2]
    v = torch.where(scores > keypoint_thresh, 2.0, 1.0)
    return torch.stack([x, y, v], -1)


# ----------------------------------------------------------------------------------------------
# a14 argmax + DARK + UDP back-projection (HF)
# ----------------------------------------------------------------------------------------------

def argmax_predictions(heatmaps: np.ndarray):
    """HF:175-205 — flat argmax (first maximum), score = max; coordinates -1 where score <= 0."""
    n, k, _, w = heatmaps.shape
    flat = heatmaps.reshape(n, k, -1)
    idx = np.argmax(flat, 2)
    scores = np.amax(flat, 2)[..., None]
    preds = np.stack([idx % w, idx // w], -1).astype(np.float32)
    preds = np.where(np.tile(scores, (1, 1, 2)) > 0.0, preds, -1).astype(np.float32)
    return preds, scores, idx


def dark_refine_full(coords: np.ndarray, heatmaps: np.ndarray, kernel: int = 11) -> np.ndarray:
    """HF:208-265 on whole maps: Gaussian blur (sigma 0.8, radius (kernel-1)//2, scipy 'reflect'),
    clip to [1e-3, 50], log, edge-pad by 1, 7-tap gradient/Hessian at the arg-max, Newton step with
    the (H + eps*I) inverse.  Including HF's indexing of the flattened padded batch — which for
    joints with coordinate -1 (score <= 0) reads the tail of the PREVIOUS map (numpy negative
    indices wrap to the last map for the first one)."""
    from scipy.ndimage import gaussian_filter

    n, k, h, w = heatmaps.shape
    r = int((kernel - 1) // 2)
    blur = np.array([[gaussian_filter(m, sigma=0.8, radius=(r, r), axes=(0, 1)) for m in maps] for maps in heatmaps])
    blur = np.log(np.clip(blur, 0.001, 50))
    pad = np.pad(blur, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge").flatten()
    coords = coords.copy()
    index = coords[..., 0] + 1 + (coords[..., 1] + 1) * (w + 2)
    index += (w + 2) * (h + 2) * np.arange(0, n * k).reshape(-1, k)
    index = index.astype(int).reshape(-1, 1)
    i_ = pad[index]
    ix1, iy1 = pad[index + 1], pad[index + w + 2]
    ix1y1, ix1_y1_ = pad[index + w + 3], pad[index - w - 3]
    ix1_, iy1_ = pad[index - 1], pad[index - 2 - w]
    dx, dy = 0.5 * (ix1 - ix1_), 0.5 * (iy1 - iy1_)
    deriv = np.concatenate([dx, dy], axis=1).reshape(n, k, 2, 1)
    dxx = ix1 - 2 * i_ + ix1_
    dyy = iy1 - 2 * i_ + iy1_
    dxy = 0.5 * (ix1y1 - ix1 - iy1 + i_ + i_ - ix1_ - iy1_ + ix1_y1_)
    hess = np.concatenate([dxx, dxy, dxy, dyy], axis=1).reshape(n, k, 2, 2)
    hess = np.linalg.inv(hess + np.finfo(np.float32).eps * np.eye(2))
    coords -= np.einsum("ijmn,ijnk->ijmk", hess, deriv).squeeze(-1)
    return coords


def transform_preds(coords: np.ndarray, center: np.ndarray, scale: np.ndarray, out_hw) -> np.ndarray:
    """HF:268-313 — UDP back-projection, the exact inverse of the crop warp (fp32 arithmetic)."""
    scale = scale * 200.0
    sy = scale[1] / (out_hw[0] - 1.0)
    sx = scale[0] / (out_hw[1] - 1.0)
    out = np.ones_like(coords)
    out[:, 0] = coords[:, 0] * sx + center[0] - scale[0] * 0.5
    out[:, 1] = coords[:, 1] * sy + center[1] - scale[1] * 0.5
    return out


def hf_dark_decode(heatmaps: np.ndarray, boxes: Sequence[Sequence[float]], kernel: int = 11,
                   crop_hw: Tuple[int, int] = (256, 192)):
    """HF:450-463 + 465-520 — ``post_process_pose_estimation``: boxes (COCO x,y,w,h) -> centre/scale
    (aspect of the 192x256 model input), arg-max, DARK, back-projection.  Returns
    ``(keypoints[P,K,2] fp32 image px, scores[P,K] fp32, argmax_idx[P,K] int64)``."""
    n, k, h, w = heatmaps.shape
    heatmaps = np.ascontiguousarray(heatmaps, dtype=np.float32)
    coords, scores, idx = argmax_predictions(heatmaps)
    preds = dark_refine_full(coords, heatmaps, kernel=kernel)
    for i in range(n):
        c, s = box_to_center_and_scale(boxes[i], image_width=crop_hw[1], image_height=crop_hw[0])
        preds[i] = transform_preds(preds[i], c, s, [h, w])
    return preds.astype(np.float32), scores[..., 0].astype(np.float32), idx.astype(np.int64)


# ---- windowed restatement of the same thing: what a fused per-(crop, joint) kernel computes ------

def _gauss_weights(sigma: float = 0.8, radius: int = 5) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) — fp64, normalised."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def _reflect(i: int, n: int) -> int:
    """scipy 'reflect' (d c b a | a b c d | d c b a)."""
    if n == 1:
        return 0
    period = 2 * n
    i %= period
    return i if i < n else period - 1 - i


def _blurred_log_at(m: np.ndarray, y: int, x: int, wts: np.ndarray, r: int) -> np.float32:
    """log(clip(blur(m)[y, x])) where blur = separable Gaussian, axis 0 first with an fp32
    intermediate (scipy filters each axis in fp64 and stores the result in the fp32 output array)."""
    h, w = m.shape
    col = np.empty(2 * r + 1, np.float32)
    for j, dx in enumerate(range(-r, r + 1)):
        xx = _reflect(x + dx, w)
        acc = 0.0
        for i, dy in enumerate(range(-r, r + 1)):
            acc += wts[i] * float(m[_reflect(y + dy, h), xx])
        col[j] = np.float32(acc)
    acc = 0.0
    for j in range(2 * r + 1):
        acc += wts[j] * float(col[j])
    v = np.float32(acc)
    v = np.float32(min(max(v, np.float32(0.001)), np.float32(50)))
    return np.log(v)


def dark_decode_local(heatmaps: np.ndarray, boxes, kernel: int = 11, crop_hw=(256, 192)):
    """Per-(crop, joint) restatement of :func:`hf_dark_decode` that touches only the 13x13 raw window
    around the arg-max (7 taps x 11x11 blur), with the taps edge-clamped to the map (``np.pad(mode=
    'edge')``), and the score<=0 case reading the bottom corners of the previous map exactly as HF's
    flat indexing does.  Python loops: small cases only."""
    n, k, h, w = heatmaps.shape
    r = int((kernel - 1) // 2)
    wts = _gauss_weights(0.8, r)
    coords, scores, idx = argmax_predictions(np.ascontiguousarray(heatmaps, np.float32))
    out = coords.copy()
    eps = np.finfo(np.float32).eps
    flat = heatmaps.reshape(n * k, h, w)
    for q in range(n * k):
        m = flat[q]
        cx, cy = coords.reshape(-1, 2)[q]

        def L(mm, yy, xx):
            return _blurred_log_at(mm, min(max(yy, 0), h - 1), min(max(xx, 0), w - 1), wts, r)

        if cx >= 0:
            x, y = int(cx), int(cy)
            i_, ix1, iy1 = L(m, y, x), L(m, y, x + 1), L(m, y + 1, x)
            ix1y1, ix1_y1_ = L(m, y + 1, x + 1), L(m, y - 1, x - 1)
            ix1_, iy1_ = L(m, y, x - 1), L(m, y - 1, x)
        else:
            prev = flat[(q - 1) % (n * k)]
            i_ = ix1 = iy1 = ix1y1 = L(m, 0, 0)
            ix1_ = L(prev, h - 1, w - 1)          # padded[-1]
            iy1_ = L(prev, h - 1, 0)              # padded[-(w+2)]
            ix1_y1_ = L(prev, h - 1, w - 1)       # padded[-(w+3)] = row h (last real row), col w+1
        dx, dy = np.float32(0.5) * (ix1 - ix1_), np.float32(0.5) * (iy1 - iy1_)
        dxx = ix1 - 2 * i_ + ix1_
        dyy = iy1 - 2 * i_ + iy1_
        dxy = np.float32(0.5) * (ix1y1 - ix1 - iy1 + i_ + i_ - ix1_ - iy1_ + ix1_y1_)
        hess = np.array([[dxx, dxy], [dxy, dyy]], np.float64) + eps * np.eye(2)
        step = np.linalg.inv(hess) @ np.array([dx, dy], np.float64)
        out.reshape(-1, 2)[q] = (np.array([cx, cy], np.float32) - step).astype(np.float32)
    for i in range(n):
        c, s = box_to_center_and_scale(boxes[i], image_width=crop_hw[1], image_height=crop_hw[0])
        out[i] = transform_preds(out[i], c, s, [h, w])
    return out.astype(np.float32), scores[..., 0].astype(np.float32), idx.astype(np.int64)


# ----------------------------------------------------------------------------------------------
# a13 argmax + quarter offset (gluoncv / Simple-Baselines) — PARITY UNPINNED
# ----------------------------------------------------------------------------------------------

def quarter_offset_decode(heatmaps: np.ndarray, centers: np.ndarray, scales: np.ndarray):
    """``get_final_preds(heatmaps, center, scale)`` as called at module_v2.py:216-220, restated from the
    published Simple-Baselines code: ``get_max_pred`` (flat arg-max, coordinates zeroed where the
    maximum is <= 0), then for interior peaks ``coord += 0.25 * sign(right-left, down-up)``, then the
    inverse of the crop's similarity transform (``transform_preds`` with scale in pixels:
    ``img = coord * scale_w / W_hm + center - scale_w/2 * (1, H_hm/W_hm)``).
    Returns (preds[P,K,2] fp32, maxvals[P,K] fp32, argmax_idx[P,K])."""
    n, k, h, w = heatmaps.shape
    flat = heatmaps.reshape(n, k, -1)
    idx = np.argmax(flat, 2)
    maxvals = np.amax(flat, 2).astype(np.float32)
    coords = np.stack([idx % w, idx // w], -1).astype(np.float32)
    coords *= (maxvals > 0.0)[..., None]
    for i in range(n):
        for j in range(k):
            px = int(np.floor(coords[i, j, 0] + 0.5))
            py = int(np.floor(coords[i, j, 1] + 0.5))
            if 1 < px < w - 1 and 1 < py < h - 1:
                m = heatmaps[i, j]
                d = np.array([m[py, px + 1] - m[py, px - 1], m[py + 1, px] - m[py - 1, px]], np.float32)
                coords[i, j] += np.sign(d).astype(np.float32) * np.float32(0.25)
    preds = np.empty_like(coords)
    for i in range(n):
        r = np.float32(scales[i][0]) / np.float32(w)          # isotropic: W_crop/scale_w == W_hm*4/scale_w
        preds[i, :, 0] = coords[i, :, 0] * r + np.float32(centers[i][0]) - np.float32(scales[i][0]) * np.float32(0.5)
        preds[i, :, 1] = coords[i, :, 1] * r + np.float32(centers[i][1]) - r * np.float32(h) * np.float32(0.5)
    return preds, maxvals, idx.astype(np.int64)
