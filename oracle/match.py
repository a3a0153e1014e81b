"""Oracle (TEST INFRASTRUCTURE): AdaFace embedding L2-norm + cosine gallery top-1, SURVEY.md §8a a5-a8."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def backbone_tail(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """libs/net_adaface.py:334-337 — ``norm = ||x||_2`` over dim 1 (keepdim), ``out = x / norm`` (no eps)."""
    norm = torch.norm(x, 2, 1, True)
    return torch.div(x, norm), norm


def l2_norm(t: torch.Tensor, axis: int = 1) -> torch.Tensor:
    """libs/head_adaface.py:39-42."""
    return torch.div(t, torch.norm(t, 2, axis, True))


def enrol_gallery(kernel: torch.Tensor, quirk_q3: bool = False) -> torch.Tensor:
    """Gallery normalisation (off the hot path, done once at enrolment).

    ``kernel`` is the AdaFace classifier kernel ``[512, N]`` whose COLUMNS are identities
    (libs/head_adaface.py:79 ``l2_norm(kernel, axis=0)``).  ``quirk_q3`` reproduces
    training/lightning/face_recognition/module.py:137, which calls ``F.normalize(kernel)`` and so
    normalises dim 1 (the wrong axis — SURVEY.md quirk Q3).  Returns row-major ``[N, 512]``."""
    kn = F.normalize(kernel) if quirk_q3 else l2_norm(kernel, axis=0)
    return kn.t().contiguous()


def match_top1(embeddings: torch.Tensor, gallery: torch.Tensor, threshold: Optional[float] = None,
               scale: float = 64.0):
    """training/lightning/face_recognition/module.py:136-145 — ``cos = F.linear(F.normalize(emb),
    kernel.t())`` with ``kernel.t()`` = the row-major gallery, ``out = cos * s``, ``pred = out.max(1)[1]``
    (first maximum wins).  Returns ``(pred[M] int64, sim[M] fp32 = cos at pred)``.

    ``threshold`` (a8) is NOT in the reference — parity unpinned, builder-defined:
    ``pred = -1 where sim < threshold``."""
    cos = F.linear(F.normalize(embeddings), gallery)
    out = cos * scale
    pred = out.max(1)[1]
    sim = cos.gather(1, pred[:, None]).squeeze(1)
    if threshold is not None:
        pred = torch.where(sim >= threshold, pred, torch.full_like(pred, -1))
    return pred, sim


def top2_gap(embeddings: torch.Tensor, gallery: torch.Tensor) -> torch.Tensor:
    """Gap between best and second-best cosine per probe (used to skip probes whose top-1 is not
    defined at fp32 resolution when asserting bit-exact identity ids)."""
    cos = F.linear(F.normalize(embeddings), gallery)
    if cos.shape[1] < 2:
        return torch.full((cos.shape[0],), float("inf"))
    t = cos.topk(2, dim=1).values
    return t[:, 0] - t[:, 1]
