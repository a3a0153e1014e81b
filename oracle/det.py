"""Oracle (TEST INFRASTRUCTURE): detection-head decode + class-aware NMS, SURVEY.md §8a rows a1-a4.

Arithmetic authority: torch CPU fp32 ops, exactly the calls the reference makes, and a plain
restatement of torchvision's CPU ``nms`` loop (third-party; the reference calls it at
``training/yolopt/util.py:162``).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

MAX_WH = 7680      # training/yolopt/util.py:124
MAX_DET = 300      # util.py:125
MAX_NMS = 30000    # util.py:126
DFL_CH = 16        # training/yolopt/nets/nn.py:233


def make_anchors(shapes: Sequence[Tuple[int, int]], strides: Sequence[float], offset: float = 0.5):
    """util.py:85-96 — per level, anchor centres (x+0.5, y+0.5) in grid units, row-major (index =
    y*W + x), levels concatenated in the given order; plus the per-anchor stride column."""
    pts, strs = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        strs.append(torch.full((h * w, 1), float(s), dtype=torch.float32))
    return torch.cat(pts), torch.cat(strs)


def dfl_expectation(box_logits: torch.Tensor) -> torch.Tensor:
    """nn.py:222-225 — ``[B, 64, A]`` side-major (channel 16*side + bin) -> softmax over the 16 bins
    -> expectation sum_j j*p_j (the reference's frozen 1x1 conv with weight arange(16))."""
    b, c, a = box_logits.shape
    p = box_logits.view(b, 4, DFL_CH, a).transpose(2, 1).softmax(1)            # [B, 16, 4, A]
    w = torch.arange(DFL_CH, dtype=torch.float32).view(1, DFL_CH, 1, 1)
    return torch.nn.functional.conv2d(p, w).view(b, 4, a)


def head_decode(levels: Sequence[torch.Tensor], strides: Sequence[float] = (8, 16, 32)) -> torch.Tensor:
    """nn.py:261-270 (eval branch of ``Head.forward`` after the per-level conv stacks): flatten and
    concatenate the levels, DFL, dist2bbox against the anchor grid, (cx, cy, w, h) * stride and
    sigmoid class scores.  Returns ``[B, 4+nc, A]``."""
    bsz, no = levels[0].shape[:2]
    nc = no - 4 * DFL_CH
    anchors, strs = make_anchors([tuple(l.shape[2:]) for l in levels], strides)
    anchors, strs = anchors.t(), strs.t()                     # [2, A], [1, A]
    x = torch.cat([l.reshape(bsz, no, -1) for l in levels], 2)
    box, cls = x.split((4 * DFL_CH, nc), 1)
    lt, rb = dfl_expectation(box).chunk(2, 1)
    a = anchors.unsqueeze(0) - lt
    b = anchors.unsqueeze(0) + rb
    box = torch.cat(((a + b) / 2, b - a), 1)
    return torch.cat((box * strs, cls.sigmoid()), 1)


def nms_greedy_np(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """torchvision CPU ``nms`` restated (csrc/ops/cpu/nms_kernel.cpp in torchvision 0.26): visit boxes
    by score, descending and stable; a visited, un-suppressed box i suppresses every later box j with
    ``inter / (area_i + area_j - inter) > thr`` (strict), all in fp32 without fused multiply-add."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    x1, y1, x2, y2 = (boxes[order, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, bool)
    keep = []
    zero = np.float32(0)
    thr = np.float32(thr)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(order[i])
        w = np.maximum(zero, np.minimum(x2[i], x2[i + 1:]) - np.maximum(x1[i], x1[i + 1:]))
        h = np.maximum(zero, np.minimum(y2[i], y2[i + 1:]) - np.maximum(y1[i], y1[i + 1:]))
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[i + 1:] - inter)
        dead[i + 1:] |= ovr > thr
    return np.asarray(keep, np.int64)


_C = None


def _load_c():
    """The same loop in plain C (oracle/nms_greedy.c), built by oracle/Makefile; optional."""
    global _C
    if _C is None:
        path = os.path.join(os.path.dirname(__file__), "_build", "liboracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.oracle_nms_greedy.restype = ctypes.c_int
            lib.oracle_nms_greedy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                              ctypes.c_void_p]
            _C = lib
        else:
            _C = False
    return _C


def nms_greedy(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    lib = _load_c()
    if not lib:
        return nms_greedy_np(boxes, scores, thr)
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    order = np.ascontiguousarray(np.argsort(-scores.astype(np.float32), kind="stable").astype(np.int32))
    keep = np.empty(n, np.int32)
    k = lib.oracle_nms_greedy(boxes.ctypes.data, order.ctypes.data, n, float(thr), keep.ctypes.data)
    return keep[:k].astype(np.int64)


def non_max_suppression(outputs: torch.Tensor, confidence_threshold: float = 0.001, iou_threshold: float = 0.65,
                        return_index: bool = False, nms_fn=nms_greedy):
    """util.py:123-169 without the wall-clock bail-out (:133-134,:166-167, load-dependent — SURVEY §5).

    ``outputs`` is ``[B, 4+nc, A]`` (cx, cy, w, h, class probabilities).  Returns the reference's
    ``list`` of ``[n_i, 6]`` rows (x1, y1, x2, y2, conf, cls); with ``return_index`` also the list
    of flat candidate keys ``anchor*nc + cls`` of the kept rows (the "keep indices").
    Ties: the reference's ``argsort`` (:157) is unstable, so equal scores have no defined order
    there; this restatement is stable, earlier candidate first.
    """
    bs, nc = outputs.shape[0], outputs.shape[1] - 4
    out, keys = [], []
    for b in range(bs):
        x = outputs[b].transpose(0, 1)                                   # [A, 4+nc]
        anchor_ids = torch.arange(x.shape[0])
        cand = x[:, 4:4 + nc].amax(1) > confidence_threshold             # :130
        x, anchor_ids = x[cand], anchor_ids[cand]
        if x.shape[0] == 0:
            out.append(torch.zeros((0, 6)))
            keys.append(torch.zeros((0,), dtype=torch.int64))
            continue
        box, cls = x.split((4, nc), 1)
        xy = box.clone()                                                 # wh2xy, :76-82
        xy[:, 0] = box[:, 0] - box[:, 2] / 2
        xy[:, 1] = box[:, 1] - box[:, 3] / 2
        xy[:, 2] = box[:, 0] + box[:, 2] / 2
        xy[:, 3] = box[:, 1] + box[:, 3] / 2
        if nc > 1:                                                       # multi-label, :147-148
            i, j = (cls > confidence_threshold).nonzero(as_tuple=False).T
            rows = torch.cat((xy[i], x[i, 4 + j, None], j[:, None].float()), 1)
            key = anchor_ids[i] * nc + j
        else:                                                            # best class, :150-151
            conf, j = cls.max(1, keepdim=True)
            sel = conf.view(-1) > confidence_threshold
            rows = torch.cat((xy, conf, j.float()), 1)[sel]
            key = (anchor_ids * nc + j.view(-1))[sel]
        if rows.shape[0] == 0:
            out.append(torch.zeros((0, 6)))
            keys.append(torch.zeros((0,), dtype=torch.int64))
            continue
        order = torch.argsort(rows[:, 4], descending=True, stable=True)[:MAX_NMS]   # :157
        rows, key = rows[order], key[order]
        off = rows[:, 5:6] * MAX_WH                                      # :160
        keep = nms_fn((rows[:, :4] + off).numpy(), rows[:, 4].numpy(), iou_threshold)[:MAX_DET]   # :161-163
        keep = torch.from_numpy(np.asarray(keep, np.int64))
        out.append(rows[keep])
        keys.append(key[keep])
    return (out, keys) if return_index else out


def near_threshold_pairs(rows: torch.Tensor, thr: float, eps: float = 1e-5) -> int:
    """Number of box pairs (same class offset applied) whose IoU lies within ``eps`` of ``thr`` —
    the parity harness re-seeds a synthetic set when this is non-zero (hard parts)."""
    if rows.shape[0] < 2:
        return 0
    b = rows[:, :4] + rows[:, 5:6] * MAX_WH
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.maximum(b[:, None, :2], b[None, :, :2])
    rb = torch.minimum(b[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    iou = inter / (area[:, None] + area[None, :] - inter)
    near = (iou - thr).abs() < eps
    return int(near.triu(1).sum())
