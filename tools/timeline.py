"""Kernel timeline of one graph replay of the cfg2 step (CUPTI through torch.profiler; there is no nsys here):
start / end of every kernel relative to the first one, so that branch overlap and tails can be read off.
python tools/timeline.py [frames_dtype]"""
import importlib
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "person-recognition-for-pose-estimation_b200"
spp = importlib.import_module(PKG)
pipeline = importlib.import_module(PKG + ".pipeline")

dev = torch.device("cuda:0")
if os.environ.get("SPP_TL_L2GRAN"):      # experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes; default 64)
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    torch.zeros(1, device=dev)
    print("cudaDeviceSetLimit(L2 fetch granularity) ->", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["SPP_TL_L2GRAN"]))))
inp = pipeline.synthetic_inputs(64, 720, 1280, 10, 17, seed=0)
if os.environ.get("SPP_TL_DET_OBJ"):     # experiment: detection heads with fewer / more planted objects than crops per frame
    n_obj = int(os.environ["SPP_TL_DET_OBJ"])
    inp.face_levels = spp.synth.make_head_maps_fast(64, 736, 1280, n_obj=n_obj, nc=1, seed=0).levels
    inp.person_levels = spp.synth.make_head_maps_fast(64, 736, 1280, n_obj=n_obj, nc=1, seed=1).levels
ms = spp.synth.make_match_set(640, 10000, seed=1000)
inp.embeddings = ms.embeddings
if len(sys.argv) > 1 and sys.argv[1] == "u8":
    inp.frames = (inp.frames * 255.0).round().clamp(0, 255).to(torch.uint8)
pipe = pipeline.SelectivePosePipeline(inp, ms.gallery.to(torch.bfloat16), dev, use_graph=os.environ.get("SPP_TL_NOGRAPH", "0") in ("", "0"),
                                      det_max_candidates=int(os.environ.get("SPP_TL_MAXCAND", "512")), match_sms=int(os.environ.get("SPP_TL_MATCH_SMS", "0")),
                                      heatmap_first=os.environ.get("SPP_TL_CROP_FIRST", "0") in ("", "0"),
                                      det_after_heatmap=int(os.environ["SPP_TL_DET_LATE"]) if os.environ.get("SPP_TL_DET_LATE") else None,
                                      crop_free_ctas=int(os.environ.get("SPP_TL_CROP_FREE", "48")))
for _ in range(5):
    pipe.step()
pipe.stream.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        pipe.step()
    pipe.stream.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# split into replays: a gap before a kernel whose name repeats the first kernel's name
names = [e.name for e in ev]
first = names[0]
starts = [i for i, n in enumerate(names) if n == first]
per = len(ev) // 4 if len(ev) % 4 == 0 else None
print("kernels:", len(ev), "per replay:", per)
if per:
    rep = ev[2 * per:3 * per]
    t0 = rep[0].time_range.start
    for e in rep:
        print(f"{(e.time_range.start - t0):9.1f} -> {(e.time_range.end - t0):9.1f} us  ({e.time_range.end - e.time_range.start:7.1f})  {e.name[:90]}")
    nxt = ev[3 * per].time_range.start - t0
    print(f"next replay starts at {nxt:.1f} us; this replay ends at {max(e.time_range.end for e in rep) - t0:.1f} us")
