"""Microbench of the detection ops at cfg2 (64 frames, 1280x736 letterbox, A=19320): python tools/micro_det.py"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
dev = torch.device("cuda:0")
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
hm = spp.synth.make_head_maps_fast(64, 736, 1280, n_obj=10, nc=1, seed=0)
levels = [l.to(dev) for l in hm.levels]
other = [l.clone() for l in levels]            # alternate inputs: 643 MB in flight > L2
dec = spp.head_decode(levels)
full_bytes = 64 * (65 * 19320 * 4 + 5 * 19320 * 4)


def timed(fn, reps=20):
    fn(levels)
    fn(other)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for r in range(reps):
        fn(levels if r & 1 else other)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


t = timed(lambda lv: spp.head_decode(lv))
print(f"head_decode (full [B,5,A] output): {t:8.1f} us  {full_bytes / t / 1e3:8.1f} GB/s  "
      f"{full_bytes / t / 1e3 / peak:.3f} of measured HBM peak")
t = timed(lambda lv: spp.decode_nms(lv))
print(f"decode_nms (fused, from raw maps):  {t:8.1f} us")
t = timed(lambda lv: spp.nms_decoded(dec))
print(f"nms_decoded (reference signature):  {t:8.1f} us")
dense = spp.synth.make_head_maps(4, 736, 1280, n_obj=10, nc=1, seed=1, dense=True)
dl = [l.to(dev) for l in dense.levels]
t = timed(lambda lv: spp.decode_nms(dl, conf_thres=0.001), reps=5)
res = spp.decode_nms(dl, conf_thres=0.001)
print(f"decode_nms dense (4 frames, every anchor a candidate, kept {res.count.tolist()}): {t:8.1f} us")
