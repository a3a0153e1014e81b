"""Eager (no CUDA graph) passes of the cfg2 step for ncu captures:  python tools/prof_step.py [steps]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "person-recognition-for-pose-estimation_b200"
spp = importlib.import_module(PKG)
pipeline = importlib.import_module(PKG + ".pipeline")

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
cfg = {"cfg2": (64, 720, 1280, 10, 17, 10000), "small": (8, 720, 1280, 10, 17, 10000)}[workload]
b, h, w, pf, k, n = cfg
dev = torch.device("cuda:0")
inp = pipeline.synthetic_inputs(b, h, w, pf, k, seed=0)
ms = spp.synth.make_match_set(b * pf, n, seed=1000)
inp.embeddings = ms.embeddings
# the kernels bench.py's default graph holds (fused small-footprint detection kernels, heatmap decode first), launched eagerly
pipe = pipeline.SelectivePosePipeline(inp, ms.gallery.to(torch.bfloat16), dev, use_graph=False, concurrent=False,
                                      det_max_candidates=int(os.environ.get("SPP_PROF_MAXCAND", "512")))   # 2 warm-up passes inside
for _ in range(steps):
    pipe.step()
pipe.stream.synchronize()
print("ok", {k: tuple(v.shape) for k, v in pipe.out.items() if hasattr(v, "shape")})
