"""Microbench of the gallery match (cfg5): M x N sweep, CUDA events around back-to-back launches.
python tools/micro_match.py [--simt]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
dev = torch.device("cuda:0")
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
cases = [(64, 10_000), (640, 10_000), (640, 100_000), (640, 1_000_000), (5120, 100_000), (5120, 1_000_000)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    cases = [(5120, 1_000_000), (5120, 10_000_000), (64, 10_000_000)]
for m, n in cases:
    g = torch.Generator(device=dev).manual_seed(0)
    gal = torch.randn(n, 512, device=dev, generator=g)
    gal = (gal / gal.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    ids_true = torch.randint(0, n, (m,), device=dev, generator=g)
    emb = gal[ids_true].float() + 0.02 * torch.randn(m, 512, device=dev, generator=g)
    ids, sims = spp.match_top1(emb, gal, 0.4)
    ok = float((ids.long() == ids_true).float().mean())
    reps = 10 if n >= 1_000_000 else 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        spp.match_top1(emb, gal, 0.4)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    flops = 2.0 * m * n * 512
    gbytes = (n * 1024 + m * 1032) / 1e9
    print(f"M={m:5d} N={n:9d}  {us:10.1f} us  {flops / us / 1e6:8.1f} TFLOP/s ({flops / us / 1e6 / pk['bf16_tflops']:.3f} of bf16 peak)  "
          f"gallery stream {gbytes / (us * 1e-6):8.1f} GB/s ({gbytes / (us * 1e-6) / pk['hbm_gbs']:.3f} of HBM peak)  planted-id recall {ok:.3f}", flush=True)
    del gal, emb
    torch.cuda.empty_cache()
