"""Small host-side restatements needed by the shims' return values (not on the hot path)."""
from __future__ import annotations

import numpy as np
import torch


def hf_center_scale(boxes_xywh: torch.Tensor, image_width: int = 192, image_height: int = 256,
                    normalize_factor: float = 200.0, padding_factor: float = 1.25) -> torch.Tensor:
    """HF ``box_to_center_and_scale`` (image_processing_vitpose.py:68-109) for a ``[P,4]`` tensor of COCO
    boxes -> ``[P,4]`` (cx, cy, scale_x, scale_y) fp32; same arithmetic order (Python floats, fp32 storage)."""
    out = np.zeros((boxes_xywh.shape[0], 4), np.float32)
    aspect = image_width / image_height
    for i, (x, y, w, h) in enumerate(boxes_xywh.double().tolist()):
        cx, cy = np.float32(x + w * 0.5), np.float32(y + h * 0.5)
        if w > aspect * h:
            h = w * 1.0 / aspect
        elif w < aspect * h:
            w = h * aspect
        s = np.array([w / normalize_factor, h / normalize_factor], dtype=np.float32) * padding_factor
        out[i] = (cx, cy, s[0], s[1])
    return torch.from_numpy(out)
