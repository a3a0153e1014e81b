"""Oracle (TEST INFRASTRUCTURE): COCO keypoint result rows and OKS, SURVEY.md §8a a15 / §8f-3.

``coco_results`` restates the result loop of ``PoseEstimationModule.validation_step``
(training/lightning/pose_estimation/module.py:505-560); pinned by tests/golden/pose_results.npz, which
oracle/gen_golden.py produces by running that very method of the reference.

``compute_oks`` restates ``pycocotools.cocoeval.COCOeval.computeOks`` (the evaluator the reference drives
at module.py:598-615).  pycocotools is third-party, not vendored, not pinned by the reference
(requirements.txt lists no version) and not installed here: PARITY UNPINNED, the published algorithm is
restated for explicit (detection, ground truth) pairs.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

COCO_SIGMAS = np.array([.026, .025, .025, .035, .035, .079, .079, .072, .072, .062, .062,
                        .107, .107, .087, .087, .089, .089], dtype=np.float32)   # datamodule.py:37-40


def coco_results(pred_coords: np.ndarray, pred_scores: np.ndarray, boxes: np.ndarray, areas: np.ndarray,
                 masks: np.ndarray, is_crowd: np.ndarray, image_ids, keypoint_thresh: float = 0.3) -> List[dict]:
    """module.py:505-560.  ``pred_coords [B,K,2]`` normalised, ``pred_scores [B,K]``, ``boxes [B,N,4]``
    xyxy, ``areas [B,N]``, ``masks [B,N]`` bool, ``is_crowd [B,N]`` bool.  All arithmetic fp32 (a Python
    float times a 0-dim fp32 tensor stays fp32), as in the reference."""
    f32 = np.float32
    out = []
    for b in range(len(image_ids)):
        for n in range(int(masks[b].sum())):
            if is_crowd[b, n]:
                continue
            box = boxes[b, n].astype(f32)
            bw, bh = f32(box[2] - box[0]), f32(box[3] - box[1])
            kps = []
            for kpt, score in zip(pred_coords[b].astype(f32), pred_scores[b].astype(f32)):
                x = float(f32(f32(kpt[0] * bw) + box[0]))
                y = float(f32(f32(kpt[1] * bh) + box[1]))
                v = 2 if float(score) > keypoint_thresh else 1
                kps.extend([x, y, int(v)])
            out.append({"image_id": int(image_ids[b]), "category_id": 1, "keypoints": kps,
                        "score": float(pred_scores[b].astype(f32).mean()), "bbox": [float(v) for v in box],
                        "area": float(areas[b, n])})
    return out


def compute_oks(pred_xy: np.ndarray, gt_xyv: np.ndarray, area: float, sigmas: np.ndarray = COCO_SIGMAS,
                gt_box_xywh: Optional[np.ndarray] = None) -> float:
    """COCOeval.computeOks for one (detection, ground truth) pair, numpy fp64 as in pycocotools."""
    sig = np.asarray(sigmas, np.float64)
    var = (sig * 2) ** 2
    xg, yg, vg = (np.asarray(gt_xyv, np.float64)[:, i] for i in range(3))
    xd, yd = np.asarray(pred_xy, np.float64)[:, 0], np.asarray(pred_xy, np.float64)[:, 1]
    k1 = int(np.count_nonzero(vg > 0))
    if k1 > 0:
        dx, dy = xd - xg, yd - yg
    else:
        bb = np.asarray(gt_box_xywh, np.float64)
        x0, x1 = bb[0] - bb[2], bb[0] + bb[2] * 2
        y0, y1 = bb[1] - bb[3], bb[1] + bb[3] * 2
        z = np.zeros(len(sig))
        dx = np.max((z, x0 - xd), axis=0) + np.max((z, xd - x1), axis=0)
        dy = np.max((z, y0 - yd), axis=0) + np.max((z, yd - y1), axis=0)
    e = (dx ** 2 + dy ** 2) / var / (float(area) + np.spacing(1)) / 2
    if k1 > 0:
        e = e[vg > 0]
    return float(np.sum(np.exp(-e)) / e.shape[0])
