"""``torch.ops.spp.*`` — the ops registered with the PyTorch dispatcher.

Schemas + CUDA implementations (thin calls into ``ops.py`` -> C ABI) + fake ("meta") kernels for shape
inference, so the ops can be traced / exported and show up by name in profiler timelines.  Only the CUDA
dispatch key has an implementation: calling an op with CPU tensors fails in the dispatcher — there is no
CPU path.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops

_LIB = torch.library.Library("spp", "DEF")

_LIB.define("head_decode(Tensor[] levels, float[] strides) -> Tensor")
_LIB.define("nms_decoded(Tensor pred, float conf_thres, float iou_thres, int max_det) -> (Tensor, Tensor, Tensor)")
_LIB.define("decode_nms(Tensor[] levels, float[] strides, float conf_thres, float iou_thres, int max_det) -> (Tensor, Tensor, Tensor)")
_LIB.define("l2_normalize(Tensor x) -> (Tensor, Tensor)")
_LIB.define("match_top1(Tensor emb, Tensor gallery_bf16, float? threshold, int id_offset) -> (Tensor, Tensor)")
_LIB.define("crop_affine(Tensor frames, Tensor boxes, Tensor frame_idx, int out_h, int out_w, float[] mean, float[] std) -> Tensor")
_LIB.define("heatmap_decode(Tensor hm, Tensor? hm_flipped, Tensor? perm, Tensor? boxes, str mode, int kernel, int flags) "
            "-> (Tensor, Tensor, Tensor)")
_LIB.define("pose_results(Tensor keypoints, Tensor scores, Tensor? boxes_xyxy, float keypoint_thresh) -> (Tensor, Tensor)")
_LIB.define("pose_oks(Tensor pred, Tensor gt, Tensor gt_area, Tensor sigmas, Tensor? gt_boxes_xywh) -> Tensor")


def _head_decode(levels: List[torch.Tensor], strides: List[float]) -> torch.Tensor:
    return ops.head_decode(levels, strides)


def _nms_decoded(pred, conf_thres, iou_thres, max_det):
    r = ops.nms_decoded(pred, conf_thres, iou_thres, max_det)
    return r.dets, r.count, r.keys


def _decode_nms(levels, strides, conf_thres, iou_thres, max_det):
    r = ops.decode_nms(levels, strides, conf_thres, iou_thres, max_det)
    return r.dets, r.count, r.keys


def _l2_normalize(x):
    return ops.l2_normalize(x, mode="backbone")


def _match_top1(emb, gallery_bf16, threshold, id_offset):
    return ops.match_top1(emb, gallery_bf16, threshold, id_offset)


def _crop_affine(frames, boxes, frame_idx, out_h, out_w, mean, std):
    return ops.crop_affine(frames, boxes, frame_idx, (out_h, out_w), mean, std)


def _heatmap_decode(hm, hm_flipped, perm, boxes, mode, kernel, flags):
    return ops.heatmap_decode(hm, hm_flipped, perm, boxes, mode, kernel, flags)


def _pose_results(keypoints, scores, boxes_xyxy, keypoint_thresh):
    return ops.pose_results(keypoints, scores, boxes_xyxy, keypoint_thresh)


def _pose_oks(pred, gt, gt_area, sigmas, gt_boxes_xywh):
    return ops.pose_oks(pred, gt, gt_area, sigmas, gt_boxes_xywh)


for _name, _fn in (("pose_results", _pose_results), ("pose_oks", _pose_oks), ("head_decode", _head_decode), ("nms_decoded", _nms_decoded), ("decode_nms", _decode_nms),
                   ("l2_normalize", _l2_normalize), ("match_top1", _match_top1), ("crop_affine", _crop_affine),
                   ("heatmap_decode", _heatmap_decode)):
    _LIB.impl(_name, _fn, "CUDA")


# ---- fake kernels: shapes / dtypes only ------------------------------------------------------------

def _anchors(levels) -> int:
    return sum(l.shape[2] * l.shape[3] for l in levels)


@torch.library.register_fake("spp::head_decode")
def _(levels, strides):
    l0 = levels[0]
    return l0.new_empty((l0.shape[0], l0.shape[1] - 60, _anchors(levels)))


def _nms_out(ref: torch.Tensor, b: int, max_det: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return (ref.new_empty((b, max_det, 6)), ref.new_empty((b,), dtype=torch.int32), ref.new_empty((b, max_det), dtype=torch.int32))


@torch.library.register_fake("spp::nms_decoded")
def _(pred, conf_thres, iou_thres, max_det):
    return _nms_out(pred, pred.shape[0], max_det)


@torch.library.register_fake("spp::decode_nms")
def _(levels, strides, conf_thres, iou_thres, max_det):
    return _nms_out(levels[0], levels[0].shape[0], max_det)


@torch.library.register_fake("spp::l2_normalize")
def _(x):
    return x.new_empty(x.shape), x.new_empty((x.shape[0], 1))


@torch.library.register_fake("spp::match_top1")
def _(emb, gallery_bf16, threshold, id_offset):
    return emb.new_empty((emb.shape[0],), dtype=torch.int32), emb.new_empty((emb.shape[0],))


@torch.library.register_fake("spp::crop_affine")
def _(frames, boxes, frame_idx, out_h, out_w, mean, std):
    return boxes.new_empty((boxes.shape[0], 3, out_h, out_w))


@torch.library.register_fake("spp::heatmap_decode")
def _(hm, hm_flipped, perm, boxes, mode, kernel, flags):
    p, k = hm.shape[0], hm.shape[1]
    return hm.new_empty((p, k, 2)), hm.new_empty((p, k)), hm.new_empty((p, k), dtype=torch.int32)


@torch.library.register_fake("spp::pose_results")
def _(keypoints, scores, boxes_xyxy, keypoint_thresh):
    p, k = scores.shape
    return scores.new_empty((p, k, 3)), scores.new_empty((p,))


@torch.library.register_fake("spp::pose_oks")
def _(pred, gt, gt_area, sigmas, gt_boxes_xywh):
    return pred.new_empty((pred.shape[0],))
