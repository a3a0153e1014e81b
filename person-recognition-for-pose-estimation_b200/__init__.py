"""spp — B200-native selective-pose glue path (detection decode + NMS, gallery match, affine crop,
heatmap decode) behind the reference's own call signatures.

The directory name carries the project name (``person-recognition-for-pose-estimation_b200``) and is
therefore imported by string; ``import spp`` (repo-root shim) gives the same package.
"""
from . import _lib, ops, shims, synth, hostmath, pipeline, dist, torch_ops  # noqa: F401
from ._lib import SppError, build  # noqa: F401
from .ops import (  # noqa: F401
    associate, crop_affine, crop_workspace_bytes, decode_nms, det_average_precision, det_match_targets, head_decode, heatmap_decode, l2_normalize, match_top1, match_unpack_keys, nms_decoded,
    pose_oks, pose_results, to_bf16, NmsResult,
)
from .shims import (  # noqa: F401
    Gallery, VitPoseImageProcessor, backbone_tail, detect, flip_test_keypoints, get_final_preds,
    coco_keypoint_results, compute_ap, compute_metric, get_keypoints_from_heatmaps, head_conv_outputs, head_eval_forward, head_forward, l2_norm,
    non_max_suppression,
)

__all__ = [
    "SppError", "build", "ops", "shims", "synth",
    "associate", "crop_affine", "crop_workspace_bytes", "decode_nms", "det_average_precision", "det_match_targets", "compute_ap", "compute_metric", "head_decode", "heatmap_decode", "l2_normalize", "match_top1", "match_unpack_keys",
    "nms_decoded", "pose_oks", "pose_results", "to_bf16", "NmsResult",
    "Gallery", "VitPoseImageProcessor", "backbone_tail", "detect", "flip_test_keypoints", "get_final_preds",
    "coco_keypoint_results", "get_keypoints_from_heatmaps", "head_conv_outputs", "head_eval_forward", "head_forward", "l2_norm", "non_max_suppression",
]
