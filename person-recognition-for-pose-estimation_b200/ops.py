"""Low-level ops: torch CUDA tensors in, torch CUDA tensors out, work done by libspp.so.

PyTorch is only plumbing here (device memory, the current stream).  Every function validates its
arguments the way the reference surfaces errors — plain exceptions — and raises if the tensors are
not on a CUDA device: there is no CPU path.
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

MAX_WH = 7680.0     # training/yolopt/util.py:124
MAX_DET = 300       # util.py:125
MAX_NMS = 30000     # util.py:126

DECODE_MODES = {"dark": 0, "softargmax": 1, "quarter": 2, "_copy_only": 99}
CROP_VARIANTS = {"hf": 0, "udp": 0, "gluoncv": 1}


def _stream(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(name: str, *tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name}: expected torch.Tensor, got {type(t).__name__}")
        if not t.is_cuda:
            raise RuntimeError(f"{name}: tensors must live on a CUDA device (no CPU fallback); got {t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"{name}: tensors on different devices ({dev} vs {t.device})")
    return dev


def _f32c(name: str, t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------
# detection
# ------------------------------------------------------------------------------------------------

def _levels_args(levels: Sequence, strides: Sequence[float]):
    """``levels``: per pyramid level either the concatenated map ``[B, 64+nc, H, W]`` (nn.py:257) or the pair
    ``(box [B, 64, H, W], cls [B, nc, H, W])`` of conv outputs before the ``cat`` (nn.py:256-257).
    Returns (kept tensors, box ptrs, cls ptrs or None, hs, ws, strides, n, B, nc, A)."""
    if len(levels) < 1 or len(levels) > 4 or len(levels) != len(strides):
        raise ValueError("detection: need 1..4 levels and one stride per level")
    split = isinstance(levels[0], (tuple, list))
    if any(isinstance(l, (tuple, list)) != split for l in levels):
        raise ValueError("detection: levels must be all concatenated maps or all (box, cls) pairs")
    if split:
        bx = [_f32c("detection box level", l[0]) for l in levels]
        cl = [_f32c("detection class level", l[1]) for l in levels]
        b, nc = bx[0].shape[0], cl[0].shape[1]
        for x, c in zip(bx, cl):
            if x.dim() != 4 or c.dim() != 4 or x.shape[1] != 64 or c.shape[1] != nc or x.shape[0] != b or c.shape[0] != b \
                    or x.shape[2:] != c.shape[2:]:
                raise ValueError("detection: split levels must be ([B, 64, H, W], [B, nc, H, W]) with matching B, H, W")
        lv = bx
    else:
        lv = [_f32c("detection level", l) for l in levels]
        cl = None
        b, no = lv[0].shape[0], lv[0].shape[1]
        for l in lv:
            if l.dim() != 4 or l.shape[0] != b or l.shape[1] != no:
                raise ValueError("detection: every level must be [B, 64+nc, H, W] with the same B and channel count")
        nc = no - 64
    if nc < 1:
        raise ValueError("detection: need 64 box channels + nc >= 1 class channels")
    n = len(lv)
    ptrs = (ctypes.c_void_p * n)(*[l.data_ptr() for l in lv])
    cptrs = (ctypes.c_void_p * n)(*[c.data_ptr() for c in cl]) if split else None
    hs = (ctypes.c_int * n)(*[l.shape[2] for l in lv])
    ws = (ctypes.c_int * n)(*[l.shape[3] for l in lv])
    st = (ctypes.c_float * n)(*[float(s) for s in strides])
    a = sum(l.shape[2] * l.shape[3] for l in lv)
    return (lv, cl), ptrs, cptrs, hs, ws, st, n, b, nc, a


def _flat_levels(levels: Sequence) -> List[torch.Tensor]:
    return [t for l in levels for t in (l if isinstance(l, (tuple, list)) else (l,))]


def head_decode(levels: Sequence, strides: Sequence[float] = (8, 16, 32)) -> torch.Tensor:
    """``Head.forward`` eval branch (training/yolopt/nets/nn.py:255-270): raw per-level maps
    ``[B, 64+nc, H_l, W_l]`` (or ``(box, cls)`` pairs before the ``cat``) -> ``[B, 4+nc, A]`` (cx, cy, w, h in
    pixels; sigmoid scores)."""
    _need_cuda("head_decode", *_flat_levels(levels))
    keep, ptrs, cptrs, hs, ws, st, n, b, nc, a = _levels_args(levels, strides)
    out = torch.empty((b, 4 + nc, a), dtype=torch.float32, device=keep[0][0].device)
    if cptrs is None:
        _lib.check(_lib.lib().spp_head_decode(ptrs, hs, ws, st, n, b, nc, _ptr(out), _stream(out)), "spp_head_decode")
    else:
        _lib.check(_lib.lib().spp_head_decode_split(ptrs, cptrs, hs, ws, st, n, b, nc, _ptr(out), _stream(out)), "spp_head_decode_split")
    return out


class NmsResult:
    """Padded, device-resident NMS output: ``dets [B, max_det, 6]``, ``count [B]`` (negative = the candidate
    list overflowed ``max_candidates``; kept rows = ``~count``), ``keys [B, max_det]`` (anchor*nc + cls of every kept
    row, -1 padding)."""

    def __init__(self, dets, count, keys):
        self.dets, self.count, self.keys = dets, count, keys

    def kept(self) -> torch.Tensor:
        """Rows kept per image, overflow flag removed."""
        return torch.where(self.count < 0, ~self.count, self.count)

    def overflowed(self) -> torch.Tensor:
        return self.count < 0

    def to_list(self) -> List[torch.Tensor]:
        """The reference's return type: a list of ``[n_i, 6]`` tensors (one D2H of the counts)."""
        counts = self.kept().tolist()
        return [self.dets[i, :c] for i, c in enumerate(counts)]

    def keys_list(self) -> List[torch.Tensor]:
        counts = self.kept().tolist()
        return [self.keys[i, :c] for i, c in enumerate(counts)]


_ws_cache = {}
_ws_retired = []     # replaced scratch buffers stay referenced: a captured CUDA graph may still hold their address


def alloc_workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """A caller-owned, 1 KB-aligned scratch buffer (pass it as ``workspace=``).  ``SelectivePosePipeline`` owns its
    workspaces this way, so the addresses baked into its CUDA graph stay valid for the pipeline's life."""
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


def _workspace(dev: torch.device, nbytes: int, tag: str, given: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``given`` (caller-owned) if passed, else the grow-only per-(device, stream, tag) scratch buffer, 1 KB aligned.
    A buffer replaced by a larger one is retired, never freed: an earlier graph capture may replay with its address."""
    if given is not None:
        if given.dtype != torch.uint8 or not given.is_cuda or given.numel() < nbytes or given.data_ptr() % 1024:
            raise ValueError(f"workspace must be a 1 KB-aligned CUDA uint8 tensor of at least {nbytes} bytes")
        return given
    key = (dev, torch.cuda.current_stream(dev).cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ws_retired.append(buf)
        buf = alloc_workspace(dev, nbytes)
        _ws_cache[key] = buf
    return buf[:nbytes]


def nms_workspace_bytes(batch: int, num_anchors: int, nc: int, max_candidates: int = 0) -> int:
    return int(_lib.lib().spp_nms_workspace_bytes(batch, num_anchors, nc, max_candidates))


def nms_decoded(pred: torch.Tensor, conf_thres: float = 0.001, iou_thres: float = 0.65, max_det: int = MAX_DET,
                max_nms: int = MAX_NMS, max_wh: float = MAX_WH, max_candidates: int = 0,
                workspace: Optional[torch.Tensor] = None) -> NmsResult:
    """``non_max_suppression`` (training/yolopt/util.py:123-169) on a decoded ``[B, 4+nc, A]`` tensor."""
    _need_cuda("nms_decoded", pred)
    pred = _f32c("nms_decoded", pred)
    if pred.dim() != 3 or pred.shape[1] < 5:
        raise ValueError(f"nms_decoded: expected [B, 4+nc, A], got {tuple(pred.shape)}")
    b, nc, a = pred.shape[0], pred.shape[1] - 4, pred.shape[2]
    L = _lib.lib()
    dets = torch.empty((b, max_det, 6), dtype=torch.float32, device=pred.device)
    count = torch.empty((b,), dtype=torch.int32, device=pred.device)
    keys = torch.empty((b, max_det), dtype=torch.int32, device=pred.device)
    nbytes = L.spp_nms_workspace_bytes(b, a, nc, max_candidates)
    ws = _workspace(pred.device, nbytes, "nms", workspace)
    _lib.check(L.spp_nms_decoded(_ptr(pred), b, nc, a, conf_thres, iou_thres, max_det, max_nms, max_wh, max_candidates,
                                 _ptr(dets), _ptr(count), _ptr(keys), _ptr(ws), nbytes, _stream(pred)), "spp_nms_decoded")
    return NmsResult(dets, count, keys)


def set_decode_nms_mode(mode: str) -> str:
    """``"split"`` (default: scan, candidate decode, sort + NMS as three launches) or ``"fused"`` (one kernel, a CTA per
    image); identical results.  Returns the previous mode.  Call before a pipeline captures its CUDA graph."""
    prev = _lib.lib().spp_decode_nms_mode({"split": 0, "fused": 1}[mode])
    return "fused" if prev == 1 else "split"


def decode_nms(levels: Sequence, strides: Sequence[float] = (8, 16, 32), conf_thres: float = 0.001,
               iou_thres: float = 0.65, max_det: int = MAX_DET, max_nms: int = MAX_NMS, max_wh: float = MAX_WH,
               max_candidates: int = 0, out: Optional[NmsResult] = None, workspace: Optional[torch.Tensor] = None) -> NmsResult:
    """Fused ``Head.forward`` (eval) + ``non_max_suppression`` from the raw per-level maps (concatenated, or
    ``(box, cls)`` pairs straight from the head's conv stacks — no ``cat`` copy)."""
    _need_cuda("decode_nms", *_flat_levels(levels))
    keep, ptrs, cptrs, hs, ws_, st, n, b, nc, a = _levels_args(levels, strides)
    dev = keep[0][0].device
    L = _lib.lib()
    if out is None:
        out = NmsResult(torch.empty((b, max_det, 6), dtype=torch.float32, device=dev),
                        torch.empty((b,), dtype=torch.int32, device=dev),
                        torch.empty((b, max_det), dtype=torch.int32, device=dev))
    nbytes = L.spp_nms_workspace_bytes(b, a, nc, max_candidates)
    ws = _workspace(dev, nbytes, "nms", workspace)
    if cptrs is None:
        _lib.check(L.spp_decode_nms(ptrs, hs, ws_, st, n, b, nc, conf_thres, iou_thres, max_det, max_nms, max_wh,
                                    max_candidates, _ptr(out.dets), _ptr(out.count), _ptr(out.keys), _ptr(ws), nbytes,
                                    _stream(out.dets)), "spp_decode_nms")
    else:
        _lib.check(L.spp_decode_nms_split(ptrs, cptrs, hs, ws_, st, n, b, nc, conf_thres, iou_thres, max_det, max_nms, max_wh,
                                          max_candidates, _ptr(out.dets), _ptr(out.count), _ptr(out.keys), _ptr(ws), nbytes,
                                          _stream(out.dets)), "spp_decode_nms_split")
    return out


# ------------------------------------------------------------------------------------------------
# embeddings / gallery match
# ------------------------------------------------------------------------------------------------

def l2_normalize(x: torch.Tensor, mode: str = "backbone", eps: float = 1e-12, want_bf16: bool = False):
    """``mode='backbone'``: libs/net_adaface.py:334-337 -> (x/||x||, ||x|| [M,1]);
    ``mode='normalize'``: ``F.normalize`` (x / max(||x||, eps))."""
    _need_cuda("l2_normalize", x)
    x = _f32c("l2_normalize", x)
    if x.dim() != 2:
        raise ValueError("l2_normalize: expected [M, D]")
    m, d = x.shape
    out = torch.empty_like(x)
    norm = torch.empty((m, 1), dtype=torch.float32, device=x.device)
    ob = torch.empty((m, d), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    _lib.check(_lib.lib().spp_l2_normalize(_ptr(x), m, d, 0 if mode == "backbone" else 1, eps, _ptr(out), _ptr(norm),
                                           _ptr(ob), _stream(x)), "spp_l2_normalize")
    return (out, norm, ob) if want_bf16 else (out, norm)


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 (round to nearest even) — gallery enrolment."""
    _need_cuda("to_bf16", x)
    x = _f32c("to_bf16", x)
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().spp_f32_to_bf16(_ptr(x), x.numel(), _ptr(out), _stream(x)), "spp_f32_to_bf16")
    return out


def match_workspace_bytes(m: int, n: int, d: int = 512) -> int:
    return int(_lib.lib().spp_match_workspace_bytes(m, n, d))


def match_top1(emb: torch.Tensor, gallery_bf16: torch.Tensor, threshold: Optional[float] = None, id_offset: int = 0,
               want_keys: bool = False, _simt: bool = False, gallery_f32: Optional[torch.Tensor] = None,
               max_row_norm: float = 1.0, workspace: Optional[torch.Tensor] = None, out=None):
    """Cosine top-1 of every probe against a bf16 gallery ``[N, 512]`` (rows normalised at enrolment).
    ``gallery_f32``: the fp32 rows the bf16 copy was rounded from — when given, candidates are re-scored against them, so
    ids and similarities are those of the reference's fp32 ``F.linear(...).max(1)``; ``max_row_norm``: the largest gallery
    row norm (1 for a normalised gallery).  The id is the exact fp32 arg-max for any gallery content (include/spp.h).
    Returns ``(ids int32 [M], sims fp32 [M])`` (+ packed int64 keys for a cross-shard MAX reduce)."""
    _need_cuda("match_top1", emb, gallery_bf16, gallery_f32)
    emb = _f32c("match_top1", emb)
    if gallery_bf16.dtype != torch.bfloat16 or not gallery_bf16.is_contiguous():
        raise TypeError("match_top1: gallery must be a contiguous bfloat16 [N, 512] tensor (see enrol_gallery)")
    if emb.dim() != 2 or gallery_bf16.dim() != 2 or emb.shape[1] != gallery_bf16.shape[1]:
        raise ValueError(f"match_top1: shape mismatch {tuple(emb.shape)} vs {tuple(gallery_bf16.shape)}")
    if gallery_f32 is not None:
        if gallery_f32.dtype != torch.float32 or not gallery_f32.is_contiguous() or gallery_f32.shape != gallery_bf16.shape:
            raise TypeError("match_top1: gallery_f32 must be a contiguous float32 tensor of the bf16 gallery's shape")
        if _simt:
            raise ValueError("match_top1: the SIMT cross-check has no fp32-gallery mode")
    m, d = emb.shape
    n = gallery_bf16.shape[0]
    L = _lib.lib()
    if out is None:
        ids = torch.empty((m,), dtype=torch.int32, device=emb.device)
        sims = torch.empty((m,), dtype=torch.float32, device=emb.device)
        keys = torch.empty((m,), dtype=torch.int64, device=emb.device) if want_keys else None
    else:
        ids, sims, keys = out
    nbytes = L.spp_match_workspace_bytes(m, n, d)
    if nbytes == 0:
        raise ValueError(f"match_top1: unsupported shape M={m} N={n} D={d} (D must be 512, N >= 1)")
    ws = _workspace(emb.device, nbytes, "match", workspace)
    thr = float("nan") if threshold is None else float(threshold)
    if _simt:
        _lib.check(L.spp_debug_match_top1_simt(_ptr(emb), _ptr(gallery_bf16), m, n, d, thr, id_offset, _ptr(ids), _ptr(sims),
                                               _ptr(keys), _ptr(ws), nbytes, _stream(emb)), "spp_debug_match_top1_simt")
    else:
        _lib.check(L.spp_match_top1_ex(_ptr(emb), _ptr(gallery_bf16), _ptr(gallery_f32), float(max_row_norm), m, n, d, thr,
                                       id_offset, _ptr(ids), _ptr(sims), _ptr(keys), _ptr(ws), nbytes, _stream(emb)),
                   "spp_match_top1")
    return (ids, sims, keys) if want_keys else (ids, sims)


def match_unpack_keys(keys: torch.Tensor, threshold: Optional[float] = None):
    _need_cuda("match_unpack_keys", keys)
    m = keys.numel()
    ids = torch.empty((m,), dtype=torch.int32, device=keys.device)
    sims = torch.empty((m,), dtype=torch.float32, device=keys.device)
    thr = float("nan") if threshold is None else float(threshold)
    _lib.check(_lib.lib().spp_match_unpack_keys(_ptr(keys), m, thr, _ptr(ids), _ptr(sims), _stream(keys)),
               "spp_match_unpack_keys")
    return ids, sims


# ------------------------------------------------------------------------------------------------
# face -> person association (the "selective" step; not in the reference, see include/spp.h)
# ------------------------------------------------------------------------------------------------

def associate(face: "NmsResult", face_ids: torch.Tensor, person: "NmsResult", cap: int = 16):
    """Persons that contain a face with a matched identity.  ``face_ids [B, face_cap]`` int32 holds the
    identity of every face-detection row (-1 = unknown / gated / padding).  Returns
    ``(boxes [B, cap, 4] COCO xywh, ident [B, cap] int32, rows [B, cap] int32, count [B] int32)``."""
    _need_cuda("associate", face.dets, face.count, face_ids, person.dets, person.count)
    b, fc = face.dets.shape[0], face.dets.shape[1]
    pc = person.dets.shape[1]
    if face_ids.dtype != torch.int32 or tuple(face_ids.shape) != (b, fc):
        raise ValueError(f"associate: face_ids must be int32 [{b}, {fc}]")
    dev = face.dets.device
    boxes = torch.empty((b, cap, 4), dtype=torch.float32, device=dev)
    ident = torch.empty((b, cap), dtype=torch.int32, device=dev)
    rows = torch.empty((b, cap), dtype=torch.int32, device=dev)
    count = torch.empty((b,), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().spp_associate(_ptr(face.dets), _ptr(face.count), _ptr(face_ids.contiguous()), fc, _ptr(person.dets),
                                        _ptr(person.count), pc, b, cap, _ptr(boxes), _ptr(ident), _ptr(rows), _ptr(count),
                                        _stream(boxes)), "spp_associate")
    return boxes, ident, rows, count


# ------------------------------------------------------------------------------------------------
# crop
# ------------------------------------------------------------------------------------------------

# Test hook: False sends crop_affine through the one-CTA-per-item kernels (no workspace).
CROP_USE_WORKSPACE = True


def crop_affine(frames: torch.Tensor, boxes: torch.Tensor, frame_idx: torch.Tensor, out_hw: Tuple[int, int] = (256, 192),
                mean: Sequence[float] = (0.485, 0.456, 0.406), std: Sequence[float] = (0.229, 0.224, 0.225),
                variant: str = "hf", out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                planned: bool = False, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Bilinear affine crop of every box to ``out_hw`` with ``(x - mean) / std`` fused.
    ``frames [B,3,H,W]`` fp32 or uint8 (mean/std are in the frames' units: multiply ImageNet mean/std by 255 for
    uint8, as HF does when it folds the 1/255 rescale), ``boxes [P,4]`` COCO (x,y,w,h), ``frame_idx [P]`` int32.
    Runs the persistent plan + stream kernels on ``workspace`` (caller-owned, from :func:`crop_workspace_bytes` /
    :func:`alloc_workspace`) or on the per-stream scratch buffer.  ``planned=True``: the caller has already enqueued
    :func:`crop_plan` for these boxes on ``workspace`` (ordered before this call), only the stream kernel runs.
    ``out_dtype=torch.bfloat16``: the same fp32 arithmetic, the result rounded to bf16 on the store (an option for a pose
    backbone under bf16 autocast; the reference's processor returns fp32, which stays the default)."""
    _need_cuda("crop_affine", frames, boxes, frame_idx)
    u8 = frames.dtype == torch.uint8     # HF semantics for uint8 images: interpolate, round half-up to uint8, then normalise
    frames = (frames if frames.is_contiguous() else frames.contiguous()) if u8 else _f32c("crop_affine frames", frames)
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("crop_affine: out_dtype must be torch.float32 or torch.bfloat16")
    if planned and workspace is None:
        raise ValueError("crop_affine(planned=True) needs the workspace crop_plan wrote")
    boxes = _f32c("crop_affine boxes", boxes)
    if frames.dim() != 4 or frames.shape[1] != 3:
        raise ValueError(f"crop_affine: frames must be [B, 3, H, W], got {tuple(frames.shape)}")
    if boxes.dim() != 2 or boxes.shape[1] != 4 or frame_idx.shape[0] != boxes.shape[0]:
        raise ValueError("crop_affine: boxes must be [P, 4] with one frame index per box")
    if frame_idx.dtype != torch.int32:
        frame_idx = frame_idx.to(torch.int32)
    p = boxes.shape[0]
    oh, ow = out_hw
    if out is None:
        out = torch.empty((p, 3, oh, ow), dtype=out_dtype, device=frames.device)
    elif out.dtype != out_dtype or tuple(out.shape) != (p, 3, oh, ow) or not out.is_contiguous():
        raise ValueError(f"crop_affine: out must be a contiguous {out_dtype} tensor of shape {(p, 3, oh, ow)}")
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    ws = None
    if CROP_USE_WORKSPACE or planned:
        ws = _workspace(frames.device, max(crop_workspace_bytes(p, oh, ow, u8), 16), "crop", workspace)
    _lib.check(_lib.lib().spp_crop_affine_ex(_ptr(frames), 1 if u8 else 0, frames.shape[0], frames.shape[2], frames.shape[3], _ptr(boxes),
                                             _ptr(frame_idx.contiguous()), p, oh, ow, m3, s3, CROP_VARIANTS[variant], _ptr(out),
                                             1 if out_dtype == torch.bfloat16 else 0, _ptr(ws), 0 if ws is None else ws.numel(),
                                             1 if planned else 0, _stream(out)), "spp_crop_affine_ex")
    return out


def crop_plan(boxes: torch.Tensor, frame_idx: torch.Tensor, frames_shape: Sequence[int], frames_u8: bool, workspace: torch.Tensor,
              out_hw: Tuple[int, int] = (256, 192), variant: str = "hf") -> None:
    """First half of :func:`crop_affine` on its own: source map, coordinate tables and band layout of every box into
    ``workspace`` (on the current stream).  Follow with ``crop_affine(..., workspace=workspace, planned=True)``."""
    _need_cuda("crop_plan", boxes, frame_idx)
    boxes = _f32c("crop_plan boxes", boxes)
    if frame_idx.dtype != torch.int32:
        frame_idx = frame_idx.to(torch.int32)
    b, _, fh, fw = frames_shape
    _lib.check(_lib.lib().spp_crop_plan(1 if frames_u8 else 0, int(b), int(fh), int(fw), _ptr(boxes), _ptr(frame_idx.contiguous()),
                                        boxes.shape[0], out_hw[0], out_hw[1], CROP_VARIANTS[variant], _ptr(workspace), workspace.numel(),
                                        _stream(boxes)), "spp_crop_plan")


def crop_workspace_bytes(p: int, out_h: int = 256, out_w: int = 192, frames_u8: bool = False) -> int:
    """Bytes of scratch the persistent crop kernels need for ``p`` crops (planned tables, item descriptors, ticket counter)."""
    return int(_lib.lib().spp_crop_workspace_bytes(int(p), int(out_h), int(out_w), 1 if frames_u8 else 0))


# ------------------------------------------------------------------------------------------------
# heatmap decode
# ------------------------------------------------------------------------------------------------

FLAG_SCALE_SCORE = 1
FLAG_BACKPROJECT = 2
FLAG_CENTER_SCALE = 4
FLAG_HF_F32_INDEX = 8     # DARK: reproduce HF's float32 flat tap index (reference quirk Q6, include/spp.h)


def heatmap_decode(hm: torch.Tensor, hm_flipped: Optional[torch.Tensor] = None, perm: Optional[torch.Tensor] = None,
                   boxes: Optional[torch.Tensor] = None, mode: str = "dark", kernel: int = 11, flags: int = 0,
                   crop_hw: Tuple[int, int] = (256, 192), out=None):
    """One pass over the heatmaps: flip-average (optional), arg-max, refinement, back-projection.
    ``hm`` / ``hm_flipped``: fp32, or bf16 (half the HBM bytes; widened to fp32 on load, then the same arithmetic).
    Returns ``(keypoints [P,K,2] fp32, scores [P,K] fp32, argmax [P,K] int32)``."""
    _need_cuda("heatmap_decode", hm, hm_flipped, perm, boxes)
    bf16 = hm.dtype == torch.bfloat16
    if bf16:
        hm = hm if hm.is_contiguous() else hm.contiguous()
    else:
        hm = _f32c("heatmap_decode", hm)
    if hm.dim() != 4:
        raise ValueError(f"heatmap_decode: expected [P, K, H, W], got {tuple(hm.shape)}")
    if hm_flipped is not None:
        if hm_flipped.dtype != hm.dtype:
            raise TypeError("heatmap_decode: heatmaps and flipped heatmaps must have the same dtype")
        hm_flipped = hm_flipped.contiguous() if bf16 else _f32c("heatmap_decode flipped", hm_flipped)
        if hm_flipped.shape != hm.shape:
            raise ValueError("heatmap_decode: flipped heatmaps must have the same shape")
    p, k, h, w = hm.shape
    if perm is not None:
        if perm.dtype != torch.int32 or perm.numel() != k:
            raise ValueError("heatmap_decode: perm must be int32 [K]")
    if boxes is not None:
        boxes = _f32c("heatmap_decode boxes", boxes)
        if boxes.shape != (p, 4):
            raise ValueError("heatmap_decode: boxes must be [P, 4]")
    if mode not in DECODE_MODES:
        raise ValueError(f"heatmap_decode: unknown mode {mode!r}")
    if out is None:
        kp = torch.empty((p, k, 2), dtype=torch.float32, device=hm.device)
        sc = torch.empty((p, k), dtype=torch.float32, device=hm.device)
        am = torch.empty((p, k), dtype=torch.int32, device=hm.device)
    else:
        kp, sc, am = out
    fn = _lib.lib().spp_heatmap_decode_bf16 if bf16 else _lib.lib().spp_heatmap_decode
    _lib.check(fn(_ptr(hm), _ptr(hm_flipped), _ptr(perm), p, k, h, w, _ptr(boxes), DECODE_MODES[mode], flags, kernel, crop_hw[0],
                  crop_hw[1], _ptr(kp), _ptr(sc), _ptr(am), _stream(hm)), "spp_heatmap_decode")
    return kp, sc, am


# ------------------------------------------------------------------------------------------------
# COCO keypoint result rows + OKS
# ------------------------------------------------------------------------------------------------

def pose_results(keypoints: torch.Tensor, scores: torch.Tensor, boxes_xyxy: Optional[torch.Tensor] = None,
                 keypoint_thresh: float = 0.3) -> Tuple[torch.Tensor, torch.Tensor]:
    """``keypoints [P,K,2]`` (normalised when ``boxes_xyxy [P,4]`` is given, image pixels otherwise) and
    ``scores [P,K]`` -> ``(rows [P,K,3] = (x, y, v), instance_score [P])`` as module.py:534-549 builds them."""
    _need_cuda("pose_results", keypoints, scores, boxes_xyxy)
    keypoints, scores = _f32c("pose_results keypoints", keypoints), _f32c("pose_results scores", scores)
    if keypoints.dim() != 3 or keypoints.shape[2] != 2 or tuple(scores.shape) != tuple(keypoints.shape[:2]):
        raise ValueError("pose_results: keypoints must be [P, K, 2] and scores [P, K]")
    p, k = scores.shape
    if boxes_xyxy is not None:
        boxes_xyxy = _f32c("pose_results boxes", boxes_xyxy)
        if tuple(boxes_xyxy.shape) != (p, 4):
            raise ValueError("pose_results: boxes must be [P, 4] (x1, y1, x2, y2)")
    rows = torch.empty((p, k, 3), dtype=torch.float32, device=keypoints.device)
    inst = torch.empty((p,), dtype=torch.float32, device=keypoints.device)
    _lib.check(_lib.lib().spp_pose_results(_ptr(keypoints), _ptr(scores), _ptr(boxes_xyxy), p, k, float(keypoint_thresh), _ptr(rows),
                                           _ptr(inst), _stream(rows)), "spp_pose_results")
    return rows, inst


def pose_oks(pred: torch.Tensor, gt: torch.Tensor, gt_area: torch.Tensor, sigmas: torch.Tensor,
             gt_boxes_xywh: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Object keypoint similarity of pairs: ``pred [P,K,2|3]``, ``gt [P,K,3]`` (x, y, v), ``gt_area [P]``,
    ``sigmas [K]`` -> ``oks [P]`` (COCOeval.computeOks; boxes are used only for pairs without a labelled joint)."""
    _need_cuda("pose_oks", pred, gt, gt_area, sigmas, gt_boxes_xywh)
    pred, gt = _f32c("pose_oks pred", pred), _f32c("pose_oks gt", gt)
    gt_area, sigmas = _f32c("pose_oks area", gt_area), _f32c("pose_oks sigmas", sigmas)
    if pred.dim() != 3 or pred.shape[2] not in (2, 3) or gt.dim() != 3 or gt.shape[2] != 3 or gt.shape[:2] != pred.shape[:2]:
        raise ValueError("pose_oks: pred must be [P, K, 2|3] and gt [P, K, 3]")
    p, k = gt.shape[0], gt.shape[1]
    if gt_area.numel() != p or sigmas.numel() != k:
        raise ValueError("pose_oks: gt_area must be [P] and sigmas [K]")
    if gt_boxes_xywh is not None:
        gt_boxes_xywh = _f32c("pose_oks boxes", gt_boxes_xywh)
    out = torch.empty((p,), dtype=torch.float32, device=pred.device)
    _lib.check(_lib.lib().spp_pose_oks(_ptr(pred), int(pred.shape[2]), _ptr(gt), _ptr(gt_boxes_xywh), _ptr(gt_area), _ptr(sigmas), p, k,
                                       _ptr(out), _stream(out)), "spp_pose_oks")
    return out


# ------------------------------------------------------------------------------------------------
# detection evaluation: compute_metric / compute_ap of training/yolopt/util.py
# ------------------------------------------------------------------------------------------------

def det_match_targets(dets: torch.Tensor, det_count: torch.Tensor, targets: torch.Tensor, target_count: torch.Tensor,
                      iou_v: Sequence[float]) -> torch.Tensor:
    """``compute_metric`` (util.py:99-120) for a batch: ``dets [B, cap, 6]`` + ``det_count [B]`` (an ``NmsResult``),
    ``targets [B, tcap, 5]`` (cls, x1, y1, x2, y2) + ``target_count [B]`` -> ``correct [B, cap, len(iou_v)]`` bool."""
    _need_cuda("det_match_targets", dets, det_count, targets, target_count)
    dets, targets = _f32c("det_match_targets dets", dets), _f32c("det_match_targets targets", targets)
    if dets.dim() != 3 or dets.shape[2] != 6 or targets.dim() != 3 or targets.shape[2] != 5 or targets.shape[0] != dets.shape[0]:
        raise ValueError("det_match_targets: dets must be [B, cap, 6] and targets [B, tcap, 5]")
    b, cap, tcap = dets.shape[0], dets.shape[1], targets.shape[1]
    t = len(iou_v)
    correct = torch.zeros((b, cap, t), dtype=torch.uint8, device=dets.device)
    if b == 0 or cap == 0:
        return correct.bool()
    if tcap == 0:
        return correct.bool()
    iv = (ctypes.c_float * t)(*[float(v) for v in iou_v])
    _lib.check(_lib.lib().spp_det_match_targets(_ptr(dets), _ptr(det_count.to(torch.int32).contiguous()), cap, _ptr(targets),
                                                _ptr(target_count.to(torch.int32).contiguous()), tcap, iv, t, b, _ptr(correct),
                                                _stream(dets)), "spp_det_match_targets")
    return correct.bool()


def det_average_precision(tp: torch.Tensor, conf: torch.Tensor, pred_cls: torch.Tensor, target_cls: torch.Tensor,
                          nc_max: int = 128, eps: float = 1e-16):
    """``compute_ap`` (util.py:225-300) on the device.  Returns a dict: ``classes [nc]`` int32, ``ap [nc, T]`` fp64,
    ``tp / fp / p / r [nc]`` fp64 at the max-F1 confidence, and the scalars ``m_pre, m_rec, map50, mean_ap, index``."""
    _need_cuda("det_average_precision", tp, conf, pred_cls, target_cls)
    n, t = int(tp.shape[0]), int(tp.shape[1])
    tp8 = tp.to(torch.uint8).contiguous()
    conf, pred_cls, target_cls = _f32c("conf", conf.float()), _f32c("pred_cls", pred_cls.float()), _f32c("target_cls", target_cls.float())
    dev = tp.device
    L = _lib.lib()
    nbytes = L.spp_det_ap_workspace_bytes(n, t, nc_max)
    if nbytes == 0:
        raise ValueError("det_average_precision: unsupported sizes")
    ws = _workspace(dev, nbytes, "det_ap")
    classes = torch.empty((nc_max,), dtype=torch.int32, device=dev)
    num = torch.zeros((1,), dtype=torch.int32, device=dev)
    ap = torch.empty((nc_max, t), dtype=torch.float64, device=dev)
    stats = torch.empty((nc_max, 4), dtype=torch.float64, device=dev)
    summary = torch.empty((6,), dtype=torch.float64, device=dev)
    _lib.check(L.spp_det_average_precision(_ptr(tp8), _ptr(conf), _ptr(pred_cls), n, _ptr(target_cls), int(target_cls.numel()), t, nc_max,
                                           float(eps), _ptr(classes), _ptr(num), _ptr(ap), _ptr(stats), _ptr(summary), _ptr(ws), nbytes,
                                           _stream(tp8)), "spp_det_average_precision")
    nc = int(num.item())
    s = summary.tolist()
    return dict(classes=classes[:nc], ap=ap[:nc], tp=stats[:nc, 0], fp=stats[:nc, 1], p=stats[:nc, 2], r=stats[:nc, 3],
                m_pre=s[0], m_rec=s[1], map50=s[2], mean_ap=s[3], index=int(s[4]))
