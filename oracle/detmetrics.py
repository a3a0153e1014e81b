"""Oracle (TEST INFRASTRUCTURE): detection evaluation — the true-positive matrix and the AP / precision / recall
summary of the reference's test loop (training/yolopt/main.py:199-234), SURVEY.md §8f-3.

Restated in numpy, each function citing the reference lines it follows; pinned by tests/golden/det_metrics.npz, which
holds the outputs of the reference's own ``compute_metric`` / ``compute_ap`` (oracle/gen_golden.py)."""
from __future__ import annotations

import numpy as np


def box_iou_matrix(target_boxes: np.ndarray, det_boxes: np.ndarray) -> np.ndarray:
    """training/yolopt/util.py:100-106 — ``iou[label, det] = inter / (area_label + area_det - inter + 1e-7)`` in fp32,
    operations in the reference's order."""
    a1, a2 = target_boxes[:, None, :2].astype(np.float32), target_boxes[:, None, 2:4].astype(np.float32)
    b1, b2 = det_boxes[None, :, :2].astype(np.float32), det_boxes[None, :, 2:4].astype(np.float32)
    wh = np.clip(np.minimum(a2, b2) - np.maximum(a1, b1), 0, None)
    inter = wh[..., 0] * wh[..., 1]
    area_a = (a2 - a1)[..., 0] * (a2 - a1)[..., 1]
    area_b = (b2 - b1)[..., 0] * (b2 - b1)[..., 1]
    return inter / (((area_a + area_b) - inter) + np.float32(1e-7))


def compute_metric(output: np.ndarray, target: np.ndarray, iou_v: np.ndarray) -> np.ndarray:
    """training/yolopt/util.py:99-120.  ``output [n, 6]`` (x1, y1, x2, y2, conf, cls) rows of one image in NMS order,
    ``target [m, 5]`` (cls, x1, y1, x2, y2), ``iou_v [T]`` -> ``correct [n, T]`` bool.

    Per threshold: pairs (label, det) with IoU >= thr and equal class; sorted by IoU descending; first occurrence per
    detection (:116), then — the array being ordered by detection index after that ``unique`` — first occurrence per
    label (:117), i.e. a label goes to the LOWEST-index detection among those whose best label it is."""
    n, t = output.shape[0], iou_v.shape[0]
    correct = np.zeros((n, t), dtype=bool)
    if n == 0 or target.shape[0] == 0:
        return correct
    iou = box_iou_matrix(target[:, 1:5], output[:, :4])
    same = target[:, 0:1].astype(np.float32) == output[None, :, 5].astype(np.float32)
    for i in range(t):
        ok = (iou >= np.float32(iou_v[i])) & same
        best = np.where(ok, iou, -np.inf)
        lab = best.argmax(0)                         # best label of every detection (ties: undefined in the reference too)
        has = ok.any(0)
        winner = {}
        for d in range(n):                           # ascending detection index: the first one claims the label
            if has[d] and lab[d] not in winner:
                winner[lab[d]] = d
        for d in winner.values():
            correct[d, i] = True
    return correct


def smooth(y: np.ndarray, f: float = 0.1) -> np.ndarray:
    """training/yolopt/util.py:172-177 — box filter of fraction f with edge padding."""
    nf = round(len(y) * f * 2) // 2 + 1
    p = np.ones(nf // 2)
    yp = np.concatenate((p * y[0], y, p * y[-1]), 0)
    return np.convolve(yp, np.ones(nf) / nf, mode="valid")


def compute_ap(tp: np.ndarray, conf: np.ndarray, pred_cls: np.ndarray, target_cls: np.ndarray, eps: float = 1e-16):
    """training/yolopt/util.py:225-300 without the plotting.  Returns ``(tp, fp, m_pre, m_rec, map50, mean_ap)`` exactly as
    the reference, plus ``extra`` = dict(ap [nc, T], p, r, f1 [nc, 1000], index, classes) for the tests."""
    i = np.argsort(-conf)
    tp, conf, pred_cls = tp[i], conf[i], pred_cls[i]
    unique_classes, nt = np.unique(target_cls, return_counts=True)
    nc = unique_classes.shape[0]
    p = np.zeros((nc, 1000))
    r = np.zeros((nc, 1000))
    ap = np.zeros((nc, tp.shape[1]))
    px = np.linspace(start=0, stop=1, num=1000)
    for ci, c in enumerate(unique_classes):
        sel = pred_cls == c
        nl, no = nt[ci], sel.sum()
        if no == 0 or nl == 0:
            continue
        fpc = (1 - tp[sel]).cumsum(0)
        tpc = tp[sel].cumsum(0)
        recall = tpc / (nl + eps)
        r[ci] = np.interp(-px, -conf[sel], recall[:, 0], left=0)
        precision = tpc / (tpc + fpc)
        p[ci] = np.interp(-px, -conf[sel], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            m_rec = np.concatenate(([0.0], recall[:, j], [1.0]))
            m_pre = np.concatenate(([1.0], precision[:, j], [0.0]))
            m_pre = np.flip(np.maximum.accumulate(np.flip(m_pre)))
            x = np.linspace(start=0, stop=1, num=101)
            y = np.interp(x, m_rec, m_pre)
            ap[ci, j] = ((x[1:] - x[:-1]) * (y[1:] + y[:-1]) / 2.0).sum()        # numpy.trapz(y, x)
    f1 = 2 * p * r / (p + r + eps)
    idx = smooth(f1.mean(0), 0.1).argmax()
    pi, ri, f1i = p[:, idx], r[:, idx], f1[:, idx]
    tpn = (ri * nt).round()
    fpn = (tpn / (pi + eps) - tpn).round()
    ap50, apm = ap[:, 0], ap.mean(1)
    extra = dict(ap=ap, p=p, r=r, f1=f1, index=int(idx), classes=unique_classes, p_at=pi, r_at=ri)
    return tpn, fpn, pi.mean(), ri.mean(), ap50.mean(), apm.mean(), extra
