"""Oracle (TEST INFRASTRUCTURE): face -> person association, SURVEY.md §8f-2.

The reference has no such step (its inference glue is a TODO at scripts/modify_models.py:71-76): the
rule is builder-defined and PARITY IS UNPINNED.  This is the plain restatement of the rule documented in
include/spp.h (``spp_associate``), fp32 arithmetic, Python loops.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def associate(face_rows: Sequence[np.ndarray], face_ids: Sequence[np.ndarray], person_rows: Sequence[np.ndarray], cap: int = 16):
    """Per frame: ``face_rows[b]`` / ``person_rows[b]`` are ``[n, 6]`` NMS rows, ``face_ids[b]`` ``[n]`` ints.
    Returns per-frame lists of (boxes xywh [k,4] fp32, identities [k], person rows [k])."""
    out_boxes: List[np.ndarray] = []
    out_ident: List[np.ndarray] = []
    out_rows: List[np.ndarray] = []
    f32 = np.float32
    for fr, fid, pr in zip(face_rows, face_ids, person_rows):
        fr, pr = np.asarray(fr, f32).reshape(-1, 6), np.asarray(pr, f32).reshape(-1, 6)
        pick = np.full(len(fr), -1, np.int64)
        for f in range(len(fr)):
            if fid[f] < 0:
                continue
            fx1, fy1, fx2, fy2 = fr[f, :4]
            cx, cy = (fx1 + fx2) * f32(0.5), (fy1 + fy2) * f32(0.5)
            farea = (fx2 - fx1) * (fy2 - fy1)
            bestv = f32(-1)
            for r in range(len(pr)):
                px1, py1, px2, py2 = pr[r, :4]
                if cx < px1 or cx > px2 or cy < py1 or cy > py2:
                    continue
                w = max(f32(0), min(fx2, px2) - max(fx1, px1))
                h = max(f32(0), min(fy2, py2) - max(fy1, py1))
                v = (w * h) / farea if farea > 0 else f32(0)
                if v > bestv:
                    bestv, pick[f] = v, r
        boxes, ident, rows = [], [], []
        for r in range(len(pr)):
            faces = np.nonzero(pick == r)[0]
            if len(faces) and len(boxes) < cap:
                x1, y1, x2, y2 = pr[r, :4]
                boxes.append([x1, y1, x2 - x1, y2 - y1])
                ident.append(int(fid[faces[0]]))
                rows.append(r)
        out_boxes.append(np.asarray(boxes, f32).reshape(-1, 4))
        out_ident.append(np.asarray(ident, np.int64))
        out_rows.append(np.asarray(rows, np.int64))
    return out_boxes, out_ident, out_rows
