"""BASELINE.json configs[4] — kernel microbench sweep: heatmap decode over 1k..1M crops (17 x 64 x 48) and gallery match
over 1k..10M ids, each point with its roofline fraction against MEASURED_PEAKS.json.
    python tools/sweep_cfg5.py heatmap > profiles/r2_sweep_heatmap.json
    python tools/sweep_cfg5.py match   > profiles/r2_sweep_match.json
Timing: CUDA events around `reps` back-to-back launches after 2 warm-ups; every working set from 10k crops / 100k ids up
exceeds the 126 MB L2, the small points are additionally run behind a 256 MB L2-flushing read and reported as `us_cold`."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def timed(fn, reps, flush=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / reps
    ts = []
    for _ in range(reps):
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(1e3 * e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def fill_noise(t, chunk=1 << 28):
    flat = t.view(-1)
    for lo in range(0, flat.numel(), chunk):
        flat[lo:lo + chunk].normal_(0.0, 0.5)


def sweep_heatmap():
    K, H, W = 17, 64, 48
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    rows = []
    perm = spp.synth.flip_perm(K, spp.synth.COCO_FLIP_PAIRS).to(dev)
    plan = [  # (crops, resident crops, flip, dtype)
        (1_000, 1_000, True, torch.float32), (10_000, 10_000, True, torch.float32), (100_000, 100_000, True, torch.float32),
        (500_000, 500_000, False, torch.float32),            # 104 GB resident (with the flip twin it would be 209 GB > HBM)
        (250_000, 250_000, True, torch.float32),             # largest resident set WITH the flip test: 104 GB
        (1_000_000, 250_000, True, torch.float32),           # 1M crops streamed: 4 passes over a resident 250k-crop buffer
        (1_000_000, 500_000, False, torch.float32),          # 1M crops, no flip: 2 passes over 500k resident
        (1_000, 1_000, True, torch.bfloat16), (10_000, 10_000, True, torch.bfloat16), (100_000, 100_000, True, torch.bfloat16),
        (500_000, 500_000, True, torch.bfloat16),            # bf16 maps: 500k crops with the flip twin fit (104 GB)
        (1_000_000, 500_000, True, torch.bfloat16),
    ]
    for crops, resident, flip, dt in plan:
        es = 4 if dt == torch.float32 else 2
        hm = torch.empty((resident, K, H, W), dtype=dt, device=dev)
        fill_noise(hm)
        fl = None
        if flip:
            fl = torch.empty((resident, K, H, W), dtype=dt, device=dev)
            fill_noise(fl)
        boxes = torch.tensor([[100.0, 80.0, 90.0, 220.0]], device=dev).repeat(resident, 1).contiguous()
        out = (torch.empty((resident, K, 2), device=dev), torch.empty((resident, K), device=dev),
               torch.empty((resident, K), dtype=torch.int32, device=dev))
        passes = crops // resident

        def run():
            for _ in range(passes):
                spp.heatmap_decode(hm, fl, perm if flip else None, boxes, "dark", 11, out=out)
        reps = 20 if crops <= 10_000 else (5 if crops <= 100_000 else 2)
        us = timed(run, reps)
        byts = crops * (K * H * W * es * (2 if flip else 1) + K * 16)
        row = {"crops": crops, "resident_crops": resident, "streamed_passes": passes, "flip_test": flip, "maps_dtype": str(dt).split(".")[1],
               "mode": "dark+udp", "us": round(us, 1), "bytes": byts, "gbs": round(byts / us / 1e3, 1),
               "frac_hbm_peak": round(byts / us / 1e3 / peaks["hbm_gbs"], 3), "crops_per_s": round(crops / us * 1e6)}
        if byts < (200 << 20):
            row["us_cold"] = round(timed(run, 10, flush), 1)
            row["frac_hbm_peak_cold"] = round(byts / row["us_cold"] / 1e3 / peaks["hbm_gbs"], 3)
        rows.append(row)
        print(json.dumps(row), file=sys.stderr)
        del hm, fl, boxes, out
        torch.cuda.empty_cache()
    return {"kernel": "heatmap_decode (DARK + UDP back-projection, K=17, 64x48)", "peak_hbm_gbs": peaks["hbm_gbs"], "points": rows}


def sweep_match():
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    rows = []
    g = torch.Generator(device=dev).manual_seed(7)
    for n in (1_000, 10_000, 100_000, 1_000_000, 10_000_000):
        gal = torch.empty((n, 512), dtype=torch.bfloat16, device=dev)
        for lo in range(0, n, 1 << 18):
            x = torch.randn((min(n, lo + (1 << 18)) - lo, 512), generator=g, device=dev)
            gal[lo:lo + x.shape[0]] = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
        for m in (64, 640, 5120):
            ids = torch.randint(0, n, (m,), generator=g, device=dev)
            probes = gal[ids].float() + 0.3 * torch.randn((m, 512), generator=g, device=dev) / 512 ** 0.5
            unknown = torch.rand(m, generator=g, device=dev) < 0.1
            probes = torch.where(unknown[:, None], torch.randn((m, 512), generator=g, device=dev), probes).contiguous()

            def run():
                spp.match_top1(probes, gal, 0.4)
            reps = 20 if n <= 100_000 else (5 if n <= 1_000_000 else 2)
            us = timed(run, reps)
            flops = 2.0 * m * n * 512
            byts = n * 1024 + m * 1032
            got, _ = spp.match_top1(probes, gal, 0.4)
            row = {"m_probes": m, "n_ids": n, "us": round(us, 1), "tflops": round(flops / us / 1e6, 1),
                   "frac_bf16_burst_peak": round(flops / us / 1e6 / peaks["bf16_tflops"], 3),
                   "frac_bf16_sustained": round(flops / us / 1e6 / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]), 3),
                   "gallery_gbs": round(byts / us / 1e3, 1), "frac_hbm_peak": round(byts / us / 1e3 / peaks["hbm_gbs"], 3),
                   "bound": "tensor" if m >= 217 else "hbm (gallery stream)",
                   "planted_recovered": bool((got[~unknown] == ids[~unknown].int()).all().item())}
            if n <= 100_000:
                row["us_cold"] = round(timed(run, 10, flush), 1)
            rows.append(row)
            print(json.dumps(row), file=sys.stderr)
        del gal
        torch.cuda.empty_cache()
    return {"kernel": "match_top1 (normalise + tcgen05 GEMM/top-2 + fp32 re-score)", "peak_bf16_tflops": peaks["bf16_tflops"],
            "peak_bf16_tflops_sustained": peaks.get("bf16_tflops_sustained"), "peak_hbm_gbs": peaks["hbm_gbs"],
            "ridge_flop_per_byte": 217, "points": rows}


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "heatmap"
    print(json.dumps(sweep_heatmap() if what == "heatmap" else sweep_match(), indent=1))
