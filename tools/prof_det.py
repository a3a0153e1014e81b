"""Fused decode + NMS of one synthetic head for ncu captures:  python tools/prof_det.py [batch] [objects_per_frame]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_obj = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
hm = spp.synth.make_head_maps_fast(b, 736, 1280, n_obj=n_obj, nc=1, seed=0)
lv = [l.to(dev) for l in hm.levels]
for _ in range(3):
    res = spp.decode_nms(lv)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    spp.decode_nms(lv, out=res)
e1.record()
torch.cuda.synchronize()
print(f"decode_nms B={b} objects={n_obj}: {e0.elapsed_time(e1) * 100:.1f} us per call; kept per frame (first 8): {res.count[:8].tolist()}")
