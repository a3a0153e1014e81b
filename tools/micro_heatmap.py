"""Microbench of the heatmap decode kernel (cfg5): crops sweep x modes; CUDA events around back-to-back launches.
python tools/micro_heatmap.py [P ...]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
Ps = [int(a) for a in sys.argv[1:]] or [640, 5120]
K = 17
for P in Ps:
    g = torch.Generator(device=dev).manual_seed(0)
    hm = torch.randn(P, K, 64, 48, device=dev, generator=g)
    fl = torch.randn(P, K, 64, 48, device=dev, generator=g)
    perm = spp.synth.flip_perm(K).to(dev)
    boxes = torch.rand(P, 4, device=dev) * 200 + 50
    for mode, flip in (("_copy_only", True), ("quarter", True), ("dark", True), ("softargmax", True), ("dark", False), ("_copy_only", False)):
        args = (hm, fl if flip else None, perm if flip else None, boxes, mode, 11)
        out = spp.heatmap_decode(*args)
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            spp.heatmap_decode(*args, out=out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        nbytes = P * K * 64 * 48 * 4 * (2 if flip else 1)
        print(f"P={P:7d} mode={mode:11s} flip={flip!s:5s} {us:9.2f} us  {nbytes / us / 1e3:8.1f} GB/s  {nbytes / us / 1e3 / peak:6.3f} of measured peak", flush=True)
