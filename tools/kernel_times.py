"""Per-kernel device durations of one op (CUPTI through torch.profiler): mean / min over `reps` eager calls.
python tools/kernel_times.py match [M] [N]  |  det [per_frame]"""
import importlib
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "person-recognition-for-pose-estimation_b200"
spp = importlib.import_module(PKG)
pipeline = importlib.import_module(PKG + ".pipeline")
dev = torch.device("cuda:0")


def run(fn, reps=20, flush=True):
    filler = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            if flush:
                filler.sum()
            fn()
        torch.cuda.synchronize()
    acc = defaultdict(list)
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and "reduce_kernel" not in e.name:
            acc[e.name].append(e.time_range.end - e.time_range.start)
    for k, v in acc.items():
        print(f"{sum(v) / len(v):9.2f} us mean  {min(v):9.2f} min  x{len(v) // reps:2d}/call  {k[:110]}")


what = sys.argv[1] if len(sys.argv) > 1 else "match"
if what == "match":
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
    ms = spp.synth.make_match_set(m, min(n, 20000), seed=1000)
    gal = ms.gallery.to(torch.bfloat16).to(dev)
    if n > gal.shape[0]:
        g = torch.Generator(device=dev).manual_seed(1)
        extra = torch.randn((n - gal.shape[0], 512), generator=g, device=dev)
        gal = torch.cat([gal, (extra / extra.norm(dim=1, keepdim=True)).to(torch.bfloat16)])
    emb = ms.embeddings.to(dev)
    run(lambda: spp.match_top1(emb, gal, 0.4))
elif what == "det":
    pf = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    hm = spp.synth.make_head_maps_fast(64, 736, 1280, n_obj=pf, nc=1, seed=0)
    lv = [l.to(dev) for l in hm.levels]
    mc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    run(lambda: spp.decode_nms(lv, max_candidates=mc))
