"""Tolerance study for SURVEY.md 8f-1 ("ViTPose decoder emits bf16 heatmaps"): what the reference's decode returns when
its fp32 heatmaps are rounded to bf16 first.  CPU only (oracle = test infrastructure):  python tools/study_bf16_heatmaps.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
synth = importlib.import_module("person-recognition-for-pose-estimation_b200.synth")
from oracle import pose as opose  # noqa: E402

hs = synth.make_heatmaps(64, 17, seed=3)
boxes = synth.make_crop_set(8, 720, 1280, per_frame=8, seed=3).boxes.tolist()


def run(hm, fl):
    avg = opose.flip_average(hm, fl, hs.perm)
    kp, sc, idx = opose.hf_dark_decode(avg.numpy(), boxes)
    c, s = opose.soft_argmax_decode(avg)
    return kp, sc, idx, c.numpy(), s.numpy()


a = run(hs.heatmaps, hs.flipped)
b = run(hs.heatmaps.bfloat16().float(), hs.flipped.bfloat16().float())
valid = a[1] > 0
same = valid & (a[2] == b[2])
rel = (np.abs(a[0] - b[0]) / np.abs(a[0]))[same]
print(f"arg-max index changed: {(a[2] != b[2]).sum()} of {a[2].size}")
print(f"DARK keypoints (same arg-max): max {np.abs(a[0] - b[0])[same].max():.4f} px, rel max {rel.max():.2e}, beyond 1e-3: {(rel > 1e-3).mean():.4%}")
print(f"DARK scores rel max {(np.abs(a[1] - b[1]) / np.abs(a[1]))[valid].max():.2e}")
print(f"soft-argmax coords abs max {np.abs(a[3] - b[3]).max():.2e}, score rel max {(np.abs(a[4] - b[4]) / a[4]).max():.2e}")
