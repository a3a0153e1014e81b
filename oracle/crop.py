"""Oracle (TEST INFRASTRUCTURE): person-box affine crop to the ViTPose input, SURVEY.md §8a a9/a10.

Variant A (the parity oracle) restates HF ``VitPoseImageProcessor`` — third-party, un-vendored,
pinned by the reference as ``transformers>=4.48.1`` (requirements.txt:10); call sites in the
reference: scripts/modify_models.py:204, training/modify_models.py:335.  ``HF:`` below is
``transformers/models/vitpose/image_processing_vitpose.py`` of the installed 5.5.0.

Variant B restates the gluoncv/Simple-Baselines ``get_affine_transform`` crop used by
training/lightning/pose_estimation/datamodule_v2.py:119-129,213-226.  gluoncv is not installed,
not vendored and not pinned by the reference, and cv2.warpAffine quantises coordinates to 1/32 px:
PARITY UNPINNED — the restatement uses exact bilinear sampling on the published matrix.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def box_to_center_and_scale(box, image_width: int = 192, image_height: int = 256,
                            normalize_factor: float = 200.0, padding_factor: float = 1.25):
    """HF:68-109 — COCO (x, y, w, h) -> centre, aspect-fixed scale / 200 * 1.25 (both fp32).
    ``box`` entries are Python floats, as in the HF docstring (``list[list[list[float]]]``)."""
    x, y, w, h = (float(v) for v in box[:4])
    aspect = image_width / image_height
    center = np.array([x + w * 0.5, y + h * 0.5], dtype=np.float32)
    if w > aspect * h:
        h = w * 1.0 / aspect
    elif w < aspect * h:
        w = h * aspect
    scale = np.array([w / normalize_factor, h / normalize_factor], dtype=np.float32)
    scale = scale * padding_factor
    return center, scale


def warp_matrix(center: np.ndarray, scale: np.ndarray, out_w: int = 192, out_h: int = 256) -> np.ndarray:
    """HF:112-146 with theta = 0 as called from HF:393-396: ``get_warp_matrix(0, center*2,
    (W-1, H-1), scale*200)`` — UDP matrix, fp32 storage, same promotion order as the HF code."""
    size_input = center * 2.0                       # fp32
    size_dst = np.array((out_w, out_h)) - 1.0       # fp64
    size_target = scale * 200.0                     # fp32
    m = np.zeros((2, 3), dtype=np.float32)
    sx = size_dst[0] / size_target[0]
    sy = size_dst[1] / size_target[1]
    m[0, 0] = 1.0 * sx
    m[0, 1] = -0.0 * sx
    m[0, 2] = sx * (-0.5 * size_input[0] * 1.0 + 0.5 * size_input[1] * 0.0 + 0.5 * size_target[0])
    m[1, 0] = 0.0 * sy
    m[1, 1] = 1.0 * sy
    m[1, 2] = sy * (-0.5 * size_input[0] * 0.0 - 0.5 * size_input[1] * 1.0 + 0.5 * size_target[1])
    return m


def sample_bilinear_zero_outside(img: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """What ``scipy.ndimage.affine_transform(order=1, mode='constant', cval=0)`` computes for an
    axis-aligned map (HF:149-172): exact bilinear in fp64; the result is exactly 0 when the source
    coordinate lies outside [0, W-1] x [0, H-1] (strict; verified against scipy 1.18).
    ``img`` [C, H, W]; ``xs`` [Wout], ``ys`` [Hout] fp64 source coordinates.  Returns [C, Hout, Wout] fp64."""
    c, h, w = img.shape
    okx = (xs >= 0) & (xs <= w - 1)
    oky = (ys >= 0) & (ys <= h - 1)
    x0 = np.clip(np.floor(xs), 0, w - 1).astype(np.int64)
    y0 = np.clip(np.floor(ys), 0, h - 1).astype(np.int64)
    x1 = np.minimum(x0 + 1, w - 1)
    y1 = np.minimum(y0 + 1, h - 1)
    tx = np.where(okx, xs - x0, 0.0)
    ty = np.where(oky, ys - y0, 0.0)
    im = img.astype(np.float64)
    r0, r1 = im[:, y0, :], im[:, y1, :]                                   # [C, Hout, W]
    top = r0[:, :, x0] * (1 - tx) + r0[:, :, x1] * tx
    bot = r1[:, :, x0] * (1 - tx) + r1[:, :, x1] * tx
    out = top * (1 - ty)[None, :, None] + bot * ty[None, :, None]
    return out * (oky[None, :, None] & okx[None, None, :])


def source_coords(m: np.ndarray, out_w: int, out_h: int):
    """Invert the fp32 2x3 push matrix in fp64 (HF:160-170) and evaluate the (axis-aligned) source
    coordinates of every output column / row."""
    m3 = np.vstack([m.astype(np.float64), [0.0, 0.0, 1.0]])
    inv = np.linalg.inv(m3)
    xs = inv[0, 0] * np.arange(out_w, dtype=np.float64) + inv[0, 2]
    ys = inv[1, 1] * np.arange(out_h, dtype=np.float64) + inv[1, 2]
    return xs, ys


def fused_mean_std(mean=IMAGENET_MEAN, std=IMAGENET_STD, rescale_factor=None):
    """transformers/image_processing_backends.py:292-306: with rescale and normalise both on, HF folds
    the rescale into mean/std (fp32 tensor * (1/rescale)); otherwise plain mean/std."""
    m = np.asarray(mean, np.float32)
    s = np.asarray(std, np.float32)
    if rescale_factor is not None:
        m = (m * np.float32(1.0 / rescale_factor)).astype(np.float32)
        s = (s * np.float32(1.0 / rescale_factor)).astype(np.float32)
    return m, s


def crop_affine_hf(frames: np.ndarray, boxes: Sequence[Sequence[float]], frame_idx: Sequence[int],
                   out_hw: Tuple[int, int] = (256, 192), mean=IMAGENET_MEAN, std=IMAGENET_STD,
                   rescale_factor=None) -> np.ndarray:
    """HF:403-448 ``_preprocess`` — for every box: centre/scale, UDP warp, bilinear sample (fp64,
    rounded to fp32 as scipy does for a float32 channel), then ``(x - mean) / std`` in fp32.
    ``frames`` [B, 3, H, W] fp32 or uint8 (uint8: the warped crop is uint8 too, as in HF); returns
    ``[P, 3, out_h, out_w]`` fp32."""
    out_h, out_w = out_hw
    m_, s_ = fused_mean_std(mean, std, rescale_factor)
    res = np.empty((len(boxes), frames.shape[1], out_h, out_w), np.float32)
    for p, (box, fi) in enumerate(zip(boxes, frame_idx)):
        c, s = box_to_center_and_scale(box, out_w, out_h)
        xs, ys = source_coords(warp_matrix(c, s, out_w, out_h), out_w, out_h)
        v = sample_bilinear_zero_outside(frames[int(fi)], xs, ys)
        if frames.dtype == np.uint8:
            # scipy writes the fp64 sample into a uint8 output array: round half up, clamp (ni_interpolation.c)
            v = np.clip(np.floor(v + 0.5), 0, 255).astype(np.uint8)
        v = v.astype(np.float32)
        res[p] = (v - m_[:, None, None]) / s_[:, None, None]
    return res


# ----------------------------------------------------------------------------------------------
# Variant B — gluoncv / Simple-Baselines crop (PARITY UNPINNED, see module docstring)
# ----------------------------------------------------------------------------------------------

def center_scale_v2(box, out_w: int = 192, out_h: int = 256):
    """datamodule_v2.py:119-129 — centre = box centre, scale = (w, h) in pixels, then the reference's
    centre shift by aspect ratio (quirk Q5: it moves the crop instead of padding it)."""
    x, y, w, h = (float(v) for v in box[:4])
    center = np.array([x + w * 0.5, y + h * 0.5])
    scale = np.array([w, h])
    aspect = out_w / out_h
    if aspect > 1:
        center[0] = center[0] + (w * 0.5 * (aspect - 1))
    else:
        center[1] = center[1] + (h * 0.5 * (1 / aspect - 1))
    return center, scale


def affine_matrix_v2(center, scale, out_w: int = 192, out_h: int = 256, inv: bool = False) -> np.ndarray:
    """Simple-Baselines ``get_affine_transform(center, scale, rot=0, output_size=[W, H])`` as called at
    datamodule_v2.py:217 (scale in pixels, no x200): three point pairs — centre, centre shifted up by
    scale_w/2 (dst: W/2), and their right-angle completion — define an isotropic similarity
    ``dst = (src - center) * (W / scale_w) + (W/2, H/2)``.  Returns the 2x3 matrix (fp64)."""
    r = out_w / float(scale[0])
    m = np.array([[r, 0.0, out_w * 0.5 - r * center[0]],
                  [0.0, r, out_h * 0.5 - r * center[1]]])
    if inv:
        m = np.array([[1 / r, 0.0, center[0] - out_w * 0.5 / r],
                      [0.0, 1 / r, center[1] - out_h * 0.5 / r]])
    return m


def crop_affine_v2(frames: np.ndarray, boxes, frame_idx, out_hw=(256, 192), mean=IMAGENET_MEAN,
                   std=IMAGENET_STD) -> np.ndarray:
    """datamodule_v2.py:213-226,83-88 with exact bilinear sampling, border value 0
    (cv2.warpAffine's defaults: INTER_LINEAR, BORDER_CONSTANT 0 — but without its 1/32-px
    coordinate quantisation).  Pixels whose source coordinate is outside the frame are 0."""
    out_h, out_w = out_hw
    m_, s_ = fused_mean_std(mean, std, None)
    res = np.empty((len(boxes), frames.shape[1], out_h, out_w), np.float32)
    for p, (box, fi) in enumerate(zip(boxes, frame_idx)):
        c, s = center_scale_v2(box, out_w, out_h)
        mi = affine_matrix_v2(c, s, out_w, out_h, inv=True)
        xs = mi[0, 0] * np.arange(out_w, dtype=np.float64) + mi[0, 2]
        ys = mi[1, 1] * np.arange(out_h, dtype=np.float64) + mi[1, 2]
        v = sample_bilinear_zero_outside(frames[int(fi)], xs, ys).astype(np.float32)
        res[p] = (v - m_[:, None, None]) / s_[:, None, None]
    return res
