#!/usr/bin/env python
"""Benchmark of the post-backbone selective-pose glue path (BASELINE.json metric: frames/s at batch 64).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [--gpus N] ...                # the reference's CPU path (oracle port)

One "step" = one pass of decode+NMS (face head, person head), gallery match, crop and heatmap decode
over one batch of synthetic backbone outputs (config 2 of BASELINE.json: 64 frames of 1280x720,
10 faces + 10 persons per frame, 10k-identity gallery, ViTPose-B heatmaps with flip test).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

PKG = "person-recognition-for-pose-estimation_b200"

WORKLOADS = {
    # name: frames per GPU, frame H, W, persons(=faces) per frame, joints, gallery ids
    "cfg2": dict(batch=64, height=720, width=1280, per_frame=10, joints=17, gallery=10000,
                 desc="batch 64 synthetic 1280x720 frames, 10 faces and persons per frame, 10k-ID gallery, "
                      "ViTPose-B 17-joint 64x48 heatmaps with flip test"),
    "cfg1": dict(batch=1, height=640, width=640, per_frame=5, joints=17, gallery=100,
                 desc="one synthetic 640x640 frame, 5 faces, 100-ID gallery, 17-joint heatmaps"),
    "cfg4": dict(batch=64, height=720, width=1280, per_frame=100, joints=133, gallery=10000,
                 desc="crowd: 64 frames, 100 persons/frame, 133-joint heatmaps with flip test"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port) — used only as the timed baseline / checker
# ----------------------------------------------------------------------------------------------

def cpu_reference_step(inp, gallery_f32, n_frames: int, per_frame: int, threshold: float = 0.4):
    """The reference functions of SURVEY.md §8c on the first ``n_frames`` frames of the batch: torch CPU
    Head decode + non_max_suppression (torchvision nms) for both heads, F.normalize/F.linear/max match,
    HF VitPoseImageProcessor.preprocess crop, flip-average + HF post_process_pose_estimation."""
    import torchvision
    from transformers import VitPoseImageProcessor
    from transformers.models.vitpose.modeling_vitpose import VitPoseEstimatorOutput
    from oracle import det as odet, match as omatch, pose as opose

    def tv_nms(boxes, scores, thr):
        return torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()

    p = n_frames * per_frame
    out = {}
    for name, levels in (("face", inp.face_levels), ("person", inp.person_levels)):
        dec = odet.head_decode([l[:n_frames] for l in levels])
        out[name] = odet.non_max_suppression(dec, 0.001, 0.65, nms_fn=tv_nms)
    out["ids"], out["sims"] = omatch.match_top1(inp.embeddings[:p], gallery_f32, threshold)
    boxes = [[[float(v) for v in inp.boxes[f * per_frame + j]] for j in range(per_frame)] for f in range(n_frames)]
    proc = cpu_reference_step.proc = getattr(cpu_reference_step, "proc", None) or VitPoseImageProcessor()
    out["pixel_values"] = proc.preprocess([inp.frames[f] for f in range(n_frames)], boxes=boxes,
                                          do_rescale=inp.frames.dtype == torch.uint8, return_tensors="pt")["pixel_values"]
    hm = inp.heatmaps[:p]
    if inp.flipped is not None:
        hm = opose.flip_average(hm, inp.flipped[:p], inp.perm)
    out["poses"] = proc.post_process_pose_estimation(VitPoseEstimatorOutput(heatmaps=hm), boxes=boxes, kernel_size=11)
    return out


def time_cpu_reference(inp, gallery_f32, batch, per_frame, budget_s: float, reps: int = 1, warm: int = 0):
    """Time the CPU path on a bounded sample: a probe on 2 frames sizes the sample to ~budget_s."""
    torch.set_num_threads(os.cpu_count() or 1)
    cpu_reference_step(inp, gallery_f32, 1, per_frame)          # untimed: imports, thread pools, first-call setup
    t0 = time.perf_counter()
    cpu_reference_step(inp, gallery_f32, min(2, batch), per_frame)
    per_frame_s = (time.perf_counter() - t0) / min(2, batch)
    total = max(1, reps + warm)
    n = int(max(1, min(batch, budget_s / (per_frame_s * total))))
    times = []
    for i in range(total):
        t0 = time.perf_counter()
        cpu_reference_step(inp, gallery_f32, n, per_frame)
        if i >= warm:
            times.append(time.perf_counter() - t0)
    return n, times


# ----------------------------------------------------------------------------------------------

def roi_bytes(boxes, frame_h, frame_w, out_h=256, out_w=192):
    """Unique source bytes a crop has to read: the padded, aspect-fixed box clipped to the frame, 3 x fp32."""
    spp = importlib.import_module(PKG)
    cs = spp.hostmath.hf_center_scale(boxes, out_w, out_h)
    w, h = cs[:, 2] * 200.0, cs[:, 3] * 200.0
    x0 = (cs[:, 0] - w / 2).clamp(0, frame_w - 1)
    x1 = (cs[:, 0] + w / 2).clamp(0, frame_w - 1)
    y0 = (cs[:, 1] - h / 2).clamp(0, frame_h - 1)
    y1 = (cs[:, 1] + h / 2).clamp(0, frame_h - 1)
    return float(((x1 - x0 + 1) * (y1 - y0 + 1)).sum()) * 3 * 4


def roi_union_bytes(boxes, frame_idx, frame_h, frame_w, out_h=256, out_w=192):
    """Source bytes that have to cross HBM at least once: the UNION of the crops' source windows per frame (boxes of one
    frame overlap; `roi_bytes` counts every window), 3 x fp32."""
    import numpy as np
    spp = importlib.import_module(PKG)
    cs = spp.hostmath.hf_center_scale(boxes, out_w, out_h)
    w, h = cs[:, 2] * 200.0, cs[:, 3] * 200.0
    x0 = (cs[:, 0] - w / 2).clamp(0, frame_w - 1).floor().long().tolist()
    x1 = (cs[:, 0] + w / 2).clamp(0, frame_w - 1).ceil().long().tolist()
    y0 = (cs[:, 1] - h / 2).clamp(0, frame_h - 1).floor().long().tolist()
    y1 = (cs[:, 1] + h / 2).clamp(0, frame_h - 1).ceil().long().tolist()
    masks = {}
    for f, a, b, c, d in zip(frame_idx.tolist(), x0, x1, y0, y1):
        m = masks.setdefault(int(f), np.zeros((frame_h, frame_w), bool))
        m[c:d + 1, a:b + 1] = True
    return float(sum(int(m.sum()) for m in masks.values())) * 3 * 4


def shared_config(args, wl, world: int, shard: bool, collective: str) -> dict:
    """The `config` object both arms print (same keys, same values: the driver compares them)."""
    return {
        "workload": f"{args.workload}: {wl['desc']}",
        "frames_per_step": wl["batch"], "crops_per_step": wl["batch"] * wl["per_frame"], "gallery_ids": wl["gallery"],
        "frames_dtype": args.frames, "decode_mode": args.decode_mode, "n_gpus": world,
    }


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin this process (and hence the pinned buffers it first-touches and its copy-issuing thread) to the CPUs of the
    GPU's NUMA node.  Plumbing for the end-to-end region only; a no-op when the box exposes a single node."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes_on_box"] = len(nodes)
        if node >= 0 and len(nodes) > 1:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            os.sched_setaffinity(0, cpus)
            info.update(cpus=len(cpus), bound=True)
    except Exception as e:      # no sysfs entry / no permission: leave the affinity alone
        info["error"] = type(e).__name__
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="spp", choices=["spp", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--decode-mode", default="dark", choices=["dark", "softargmax", "quarter"])
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--frames", default="f32", choices=["f32", "u8"],
                    help="frame dtype: f32 in [0,1] (SURVEY 8d, default) or uint8 as a video decoder delivers them")
    ap.add_argument("--select-on-device", action="store_true",
                    help="crop list = persons containing a matched face (ops.associate on the NMS output) instead of the synthetic boxes")
    ap.add_argument("--gallery-ids", type=int, default=0, help="override the workload's gallery size")
    ap.add_argument("--shard-gallery", default="auto", choices=["auto", "yes", "no"],
                    help="N>1: shard the gallery by rows or replicate it; auto = shard above 100k ids")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="sharded gallery: exchange by the match kernels over NVLink peer memory (captured in the graph), or NCCL "
                         "all_gather + all_reduce(MAX) issued eagerly beside the graph")
    ap.add_argument("--capture-collectives", action="store_true", help="--collective nccl: capture the NCCL calls into the CUDA graph (hung in testing)")
    ap.add_argument("--serial", action="store_true", help="run the four chains back to back instead of on forked streams")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the cfg3 record (1M-id gallery, sharded at N>1)")
    ap.add_argument("--cfg3-ids", type=int, default=1_000_000)
    ap.add_argument("--det-max-candidates", type=int, default=-1,
                    help="bound on NMS candidates per image (<= 512 selects the small-footprint NMS kernel); -1 = 512 for cfg1/cfg2, unbounded else")
    ap.add_argument("--match-sms", type=int, default=0, help="SMs reserved for the match GEMM beside the heatmap decode; 0 = no split (measured best)")
    ap.add_argument("--det-after-heatmap", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="detection chains that wait for the heatmap decode and run beside the crop (pipeline default: -1)")
    ap.add_argument("--crop-free-ctas", type=int, default=-1, help="CTA slots the persistent crop leaves free (pipeline default: -1)")
    ap.add_argument("--crop-first", action="store_true", help="round-1 order: crop then heatmap decode on the main stream")
    ap.add_argument("--copy-streams", type=int, default=2, help="streams the per-step H2D copies of the e2e region are spread over")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "spp" else args.warmup

    # stdout carries exactly one JSON line: anything libraries print there (e.g. NCCL's version banner)
    # is diverted to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"

    def emit(line: dict) -> None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.gallery_ids > 0:
        wl["gallery"] = args.gallery_ids
        wl["desc"] += f" [gallery overridden to {args.gallery_ids} ids]"
    spp = importlib.import_module(PKG)
    pipeline = importlib.import_module(PKG + ".pipeline")
    if args.det_max_candidates < 0:
        args.det_max_candidates = 512 if args.workload in ("cfg1", "cfg2") else 0
    pipe_kw = dict(decode_mode=args.decode_mode, use_graph=not args.no_graph, concurrent=not args.serial,
                   det_max_candidates=args.det_max_candidates, match_sms=args.match_sms, heatmap_first=not args.crop_first)
    if args.det_after_heatmap >= 0:
        pipe_kw["det_after_heatmap"] = args.det_after_heatmap
    if args.crop_free_ctas >= 0:
        pipe_kw["crop_free_ctas"] = args.crop_free_ctas

    if args.impl == "reference":
        run_reference(args, wl, rank, world, pipeline, emit)
        return

    assert torch.cuda.is_available(), "bench.py (impl spp) needs a CUDA device: there is no CPU fallback"
    numa = bind_to_gpu_numa(local_rank)          # before any pinned allocation (first touch decides the node)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)    # NCCL kernels must not queue behind the crop
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    peaks = load_peaks()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms_val: float) -> float:
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms_val], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms_val

    # ---- inputs (per-rank shard: every rank owns `batch` frames; weak scaling) ------------------
    inp = pipeline.synthetic_inputs(wl["batch"], wl["height"], wl["width"], wl["per_frame"], wl["joints"], seed=rank)
    # One gallery for the whole job (same seed on every rank); with N > 1 it is sharded by rows.  Every rank's
    # probes are planted from the full gallery.
    ms = spp.synth.make_match_set(world * wl["batch"] * wl["per_frame"], wl["gallery"], seed=1000)
    m_local = wl["batch"] * wl["per_frame"]
    inp.embeddings = ms.embeddings[rank * m_local:(rank + 1) * m_local].contiguous()
    if args.frames == "u8":
        inp.frames = (inp.frames * 255.0).round().clamp(0, 255).to(torch.uint8)
    gallery_bf16 = ms.gallery.to(torch.bfloat16)
    shard_lo, shard_hi = spp.dist.shard_bounds(wl["gallery"], world, rank)
    matcher, peers = None, None
    # Pre-flight for the peer exchange (cudaMalloc + CUDA IPC between the ranks of the box): if ANY rank cannot map its peers'
    # buffers (a box without peer access, a container that forbids IPC handles), every rank falls back to the NCCL exchange
    # together, and the JSON line says so, instead of the run dying without a line.
    collective_fallback = None
    if world > 1 and args.collective == "peer":
        ok, why, probe = 1, "", None
        try:
            probe = spp.dist.PeerGroup(8)
        except Exception as e:          # noqa: BLE001 - whatever the failure, the decision has to be collective
            ok, why = 0, f"{type(e).__name__}: {e}"
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if probe is not None:
            probe.close(barrier=int(flag.item()) == 1)     # every rank is here only if every rank succeeded
        if int(flag.item()) == 0:
            args.collective = "nccl"
            collective_fallback = "peer exchange unavailable on this box (" + (why or "another rank failed") + "): NCCL all_gather + all_reduce(MAX) instead"
    # Placement policy: a gallery of <= 100k ids (<= 102 MB bf16) is replicated on every GPU and
    # the step has no collective at all; larger galleries are sharded by rows (cfg3 record below).
    shard = world > 1 and (args.shard_gallery == "yes" or (args.shard_gallery == "auto" and wl["gallery"] > 100_000))
    if shard:          # this rank holds ids [shard_lo, shard_hi)
        shard_rows = gallery_bf16[shard_lo:shard_hi].to(dev).contiguous()
        if args.collective == "peer":
            peers = spp.dist.PeerGroup(m_local)
            matcher = spp.dist.PeerShardedMatcher(peers, shard_rows, shard_lo, 0.4)
        else:
            matcher = spp.dist.gpu_matcher(shard_rows, shard_lo, 0.4)
    barrier()          # ranks enter the first exchange step together (a peer wait is bounded at 60 s)
    pipe = pipeline.SelectivePosePipeline(inp, gallery_bf16, dev, matcher=matcher, capture_collectives=args.capture_collectives,
                                          select_on_device=args.select_on_device, **pipe_kw)
    pipe.bind_host(inp, args.copy_streams)
    B, P, K, M = wl["batch"], wl["batch"] * wl["per_frame"], wl["joints"], wl["batch"] * wl["per_frame"]
    A = sum(l.shape[2] * l.shape[3] for l in inp.face_levels)
    nc = inp.face_levels[0].shape[1] - 64

    sampler = ClockSampler(local_rank)
    st = pipe.stream

    # ---- (1) device-resident timed region --------------------------------------------------------
    for _ in range(args.warmup):
        pipe.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        e0.record(st)
        for _ in range(args.steps):
            pipe.step()
        e1.record(st)
        barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = dev_ms / args.steps
    frames_per_s = world * B * args.steps / (dev_ms / 1e3)
    det_overflow = bool(pipe.out["_face"].overflowed().any().item() or pipe.out["_person"].overflowed().any().item())
    assert not det_overflow, "a frame has more NMS candidates than --det-max-candidates: results would be truncated"

    # ---- (2) end to end from pinned host buffers ---------------------------------------------------
    def time_e2e(p, n_steps):
        for _ in range(2):
            p.run_host()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler:
            a0.record(p.stream)
            for _ in range(n_steps):
                p.run_host()
            a1.record(p.stream)
            barrier()
        return max_over_ranks(a0.elapsed_time(a1))

    e2e_steps = max(3, min(args.steps, 20))
    e2e_ms = time_e2e(pipe, e2e_steps)
    e2e_fps = world * B * e2e_steps / (e2e_ms / 1e3)

    # ---- (2b) the same end-to-end pass with uint8 frames (what a video decoder delivers, and what HF's default
    #      do_rescale=True path expects): 4x fewer frame bytes over PCIe.  Reported beside the fp32 headline at every N;
    #      `bench.py --impl reference --frames u8` is the same-config CPU arm. -----
    e2e_u8 = None
    if args.frames == "f32" and not args.select_on_device and args.workload == "cfg2" and not shard:
        import copy
        inp8 = copy.copy(inp)
        inp8.frames = (inp.frames * 255.0).round().clamp(0, 255).to(torch.uint8)
        pipe8 = pipeline.SelectivePosePipeline(inp8, gallery_bf16, dev, **pipe_kw)
        pipe8.bind_host(inp8, args.copy_streams)
        n8 = max(3, min(args.steps, 10))
        ms8 = time_e2e(pipe8, n8) / n8
        e2e_u8 = {"value": round(world * B / (ms8 / 1e3), 1), "unit": "frames/s", "frames_dtype": "u8", "h2d_bytes_per_step": pipe8.h2d_bytes,
                  "d2h_bytes_per_step": pipe8.d2h_bytes, "ms_per_step": round(ms8, 3), "steps": n8,
                  "h2d_gbs_per_gpu": round(pipe8.h2d_bytes / (ms8 * 1e6), 2)}
        del pipe8, inp8
        torch.cuda.empty_cache()

    # ---- (3) per-kernel durations (eager launches, CUDA events around each op on its stream) ------
    ops = spp.ops
    i = pipe.inp
    u8_kw = ({"mean": [m * 255.0 for m in (0.485, 0.456, 0.406)], "std": [v * 255.0 for v in (0.229, 0.224, 0.225)]}
             if args.frames == "u8" else {})
    stages = {
        "decode_nms_face": lambda: ops.decode_nms(i.face_levels, out=pipe.out["_face"]),
        "decode_nms_person": lambda: ops.decode_nms(i.person_levels, out=pipe.out["_person"]),
        "match_top1": lambda: ops.match_top1(i.embeddings, pipe.gallery, 0.4),
        "crop_affine": lambda: ops.crop_affine(i.frames, i.boxes, i.frame_idx, out=pipe.out["pixel_values"], **u8_kw),
        "heatmap_decode": lambda: ops.heatmap_decode(i.heatmaps, i.flipped, i.perm, i.boxes, args.decode_mode, 11,
                                                     out=(pipe.out["keypoints"], pipe.out["scores"], pipe.out["argmax"])),
    }
    if args.decode_mode == "dark":
        # HF's float32 flat-index behaviour (reference quirk Q6) — the mode that reproduces ONE HF call on all 640 crops
        # literally, which is what the CPU arm below times.  Not the default; timed beside it.
        q_out = tuple(torch.empty_like(pipe.out[k]) for k in ("keypoints", "scores", "argmax"))
        stages["heatmap_decode_hf_f32_index"] = lambda: ops.heatmap_decode(i.heatmaps, i.flipped, i.perm, i.boxes, "dark", 11,
                                                                           flags=ops.FLAG_HF_F32_INDEX, out=q_out)
    # ONE captured graph holding, per stage, [L2 flush, event, stage, event]: kernel-to-kernel hand-over
    # inside a graph is ~1 us, so the events bracket device time, not host launch latency.  The flush is a
    # READ of a 256 MB buffer (2x the L2): it evicts the previous stage's lines without leaving dirty
    # lines whose write-back would be billed to the stage being timed.
    n_rep = max(10, min(args.steps, 30))
    filler = torch.empty(64 << 20, dtype=torch.float32, device=dev)      # 256 MB > 126 MB L2
    with torch.cuda.stream(st):
        for fn in stages.values():
            fn()
    st.synchronize()
    ev = {k: (torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))
          for k in stages}
    # Stages whose own input is more than twice the L2 (heatmap decode: 267 MB, crop: 1 GB) are launched 4x
    # back to back between the two events — every launch misses L2 anyway, and the event / launch overhead
    # (~3 us, comparable to 5 % of a 50 us kernel) is amortised; the small-input stages get one launch after
    # the flush so that they are timed cold.
    inner = {k: (4 if k.startswith(("heatmap_decode", "crop_affine")) else 1) for k in stages}
    tg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(tg, stream=st):
        for k, fn in stages.items():
            flush_sink = filler.sum()
            ev[k][0].record(st)
            for _ in range(inner[k]):
                fn()
            ev[k][1].record(st)
    samples = {k: [] for k in stages}
    with sampler:
        for r in range(n_rep):
            with torch.cuda.stream(st):
                tg.replay()
            st.synchronize()
            for k in stages:
                samples[k].append(ev[k][0].elapsed_time(ev[k][1]) / inner[k])
    kern_us = {k: 1e3 * statistics.median(v) for k, v in samples.items()}
    del filler

    traffic_all = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic_all = json.load(f).get(args.workload, {})
    traffic_note = "static: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/ncu_traffic.json), not measured in this run"

    hm_bytes = P * (K * 64 * 48 * 4 * (2 if i.flipped is not None else 1) + K * 16)
    det_full_bytes = B * ((64 + nc) * A * 4 + 300 * 6 * 4 + 4)
    crop_bytes = P * 3 * 256 * 192 * 4 + roi_bytes(inp.boxes, wl["height"], wl["width"]) * (0.25 if args.frames == "u8" else 1.0)
    match_flops = 2.0 * M * wl["gallery"] * 512
    kernels = {
        "heatmap_decode": dict(bound="hbm", us=kern_us["heatmap_decode"], bytes=hm_bytes),
        # `bytes` = SURVEY 8(d)'s algorithmic figure: output + every crop's own source window.  Boxes of one frame overlap, and
        # the kernel keeps the crops of a frame in flight together so that shared rows come from L2: `frac` can exceed 1.
        # `min_bytes` = output + the UNION of the windows, what has to cross HBM at least once.
        "crop_affine": dict(bound="hbm", us=kern_us["crop_affine"], bytes=crop_bytes,
                            min_bytes=P * 3 * 256 * 192 * 4 + roi_union_bytes(inp.boxes, inp.frame_idx, wl["height"], wl["width"]) *
                            (0.25 if args.frames == "u8" else 1.0),
                            note="frac = algorithmic bytes (every crop's own source window) / time / peak: above 1 because overlapping boxes of a "
                                 "frame share source rows through L2; frac_min = (output + union of the windows) / time / peak; frac_dram = ncu DRAM "
                                 "bytes (static) / time / peak"),
        # decode+NMS reads the class planes of every anchor and the 64 DFL planes of candidate anchors only: a
        # latency-bound chain of three small launches, not an HBM stream.  `survey_bytes` (every plane of the
        # head, what the reference's Head.forward reads) is given for scale only — no fraction of peak is derived from it.
        "decode_nms_face": dict(bound="latency", us=kern_us["decode_nms_face"], survey_bytes=det_full_bytes),
        "decode_nms_person": dict(bound="latency", us=kern_us["decode_nms_person"], survey_bytes=det_full_bytes),
        "match_top1": dict(bound="tensor", us=kern_us["match_top1"], flops=match_flops),
    }
    for k, d in kernels.items():
        tr = traffic_all.get(k)
        if d["bound"] == "hbm":
            d["achieved"] = d["bytes"] / (d["us"] * 1e-6) / 1e9
            d["peak"], d["unit"] = peaks["hbm"], "GB/s"
            d["frac"] = d["achieved"] / d["peak"]
            if tr:
                d["traffic"] = tr
                d["frac_dram"] = tr / (d["us"] * 1e-6) / 1e9 / peaks["hbm"]
            if "min_bytes" in d:
                d["frac_min"] = d["min_bytes"] / (d["us"] * 1e-6) / 1e9 / peaks["hbm"]
        elif d["bound"] == "tensor":
            d["achieved"] = d["flops"] / (d["us"] * 1e-6) / 1e12
            d["peak"], d["unit"] = peaks["bf16"], "TFLOP/s"
            d["frac"] = d["achieved"] / d["peak"]
        else:
            if tr:
                d["bytes_read_actual"] = tr
                d["achieved"] = tr / (d["us"] * 1e-6) / 1e9
                d["peak"], d["unit"] = peaks["hbm"], "GB/s"
                d["frac"] = d["achieved"] / d["peak"]
            d["note"] = "latency-bound (3 launches, one CTA per frame in the NMS); frac = actual DRAM bytes / time / peak, from the static ncu capture"
    dominant = max((k for k in kernels if kernels[k]["bound"] != "latency"), key=lambda k: kernels[k]["us"])
    dk = kernels[dominant]
    roofline = dict(kernel=dominant, bound=dk["bound"], achieved=round(dk["achieved"], 1), peak=dk["peak"], unit=dk["unit"],
                    frac=round(dk["frac"], 4), traffic=dk.get("traffic"), frac_dram=round(dk["frac_dram"], 4) if "frac_dram" in dk else None,
                    traffic_source=traffic_note, peak_source=peaks["source"],
                    launch_us=round(dk["us"], 2), algorithmic=dk.get("bytes", dk.get("flops")),
                    **({"frac_min": round(dk["frac_min"], 4), "min_bytes": dk["min_bytes"], "note": dk["note"]} if "frac_min" in dk else {}),
                    step_share=round(dk["us"] / sum(v["us"] for v in kernels.values()), 3))
    step_floor = None
    if traffic_all:
        tot = sum(traffic_all.get(k, 0) for k in kernels)
        step_floor = {"dram_bytes_per_step": tot, "floor_us": round(tot / peaks["hbm"] / 1e3, 1),
                      "frac_of_floor": round(tot / peaks["hbm"] / 1e3 / (ms_per_step * 1e3), 3), "source": traffic_note}

    # ---- (3b) cfg3: 1M-id gallery, local at N=1, row-sharded at N>1 (north-star config 3) --------------
    cfg3 = None
    if not args.no_cfg3 and args.workload == "cfg2" and not args.select_on_device:
        cfg3 = run_cfg3(args, spp, pipeline, pipe, dev, rank, world, B, wl["per_frame"], barrier, max_over_ranks, peaks, pipe_kw)

    # ---- (4) CPU baseline on this host (rank 0, N=1 only) ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n, times = time_cpu_reference(inp, ms.gallery, B, wl["per_frame"], args.cpu_budget)
        cpu = dict(value=round(n / statistics.median(times), 3), unit="frames/s", cores=os.cpu_count(), kind="port",
                   sample=f"{n} of {B} frames of {args.workload} ({n * wl['per_frame']} crops), 1 timed pass. " + CPU_ARM_TEXT,
                   seconds=round(statistics.median(times), 3))

    if rank == 0:
        cfg = shared_config(args, wl, world, shard, args.collective)
        line = {
            "metric": "frames/sec post-backbone selective-pose pipeline (batch 64, 1/2/4/8 B200)",
            "value": round(frames_per_s, 1), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "impl_detail": {
                "frames_per_gpu": B, "crops_per_gpu": P,
                "arithmetic": "fp32 throughout; gallery match = bf16 tcgen05 candidate search, every candidate inside the bf16 error band "
                              "re-scored in exact fp32 (the id is the fp32 arg-max for any gallery content)",
                "decode_quirk": "intended" if args.decode_mode == "dark" else None,
                "decode_quirk_note": "DARK taps at their true positions (= HF called on <= 299 crops at a time). The CPU arm calls HF once on all "
                                     "crops, which from crop 299 on reads its taps through a rounded float32 index (reference quirk Q6); "
                                     "kernels.heatmap_decode_hf_f32_index times the mode that reproduces that literally",
                "parallelism": f"dp{world}: frames/crops/heatmaps sharded with no collective" + (
                    f"; the {wl['gallery']}-id gallery sharded by rows ({wl['gallery'] // world} per GPU), exchange = {args.collective}"
                    if shard else (f"; the {wl['gallery']}-id gallery replicated per GPU (policy: shard above 100k ids; the sharded case is the cfg3 record)" if world > 1 else "")),
                "l2": "step region: inputs (1.6 GB per step) are larger than the 126 MB L2, no flush; "
                      "per-kernel region: a 256 MB read before each stage evicts L2 (cold, clean); heatmap decode and crop "
                      "(inputs > 2x L2) are the mean of 4 back-to-back launches after the flush",
                "cuda_graph": not args.no_graph, "crop_boxes": "selected on the device from the detections" if args.select_on_device else "synthetic input boxes",
                "streams": "serial" if args.serial else ("4 forked chains (" + ("crop -> heatmap decode" if args.crop_first else "heatmap decode -> crop") +
                                                           " | det face | det person | match)"),
                "det_max_candidates": args.det_max_candidates or "unbounded",
                "det_overflow": det_overflow,
                "sm_split": (f"match GEMM on {args.match_sms} SMs beside the heatmap decode on the others" if args.match_sms else "none"),
                **({"collective_fallback": collective_fallback} if collective_fallback else {}),
                "host_numa": numa,
            },
            "crops_per_s": round(world * P * args.steps / (dev_ms / 1e3), 1),
            "e2e": {"value": round(e2e_fps, 1), "unit": "frames/s", "h2d_bytes_per_step": pipe.h2d_bytes,
                    "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": round(e2e_ms / e2e_steps, 3), "steps": e2e_steps,
                    "h2d_gbs_per_gpu": round(pipe.h2d_bytes / (e2e_ms / e2e_steps * 1e6), 2), "copy_streams": pipe.copy_streams},
            "e2e_u8_frames": e2e_u8,
            "gpu_launches": pipe.launches_per_step * args.steps,
            "roofline": roofline,
            "step_dram_floor": step_floor,
            "kernels": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in d.items()} for k, d in kernels.items()},
            "cfg3": cfg3,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
        }
        if "heatmap_decode_hf_f32_index" in kern_us:
            line["kernels"]["heatmap_decode_hf_f32_index"] = {"us": round(kern_us["heatmap_decode_hf_f32_index"], 3),
                                                              "note": "SPP_DECODE_FLAG_HF_F32_INDEX: literal HF float32 tap index (quirk Q6)"}
        emit(line)
    if peers is not None:
        peers.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


CPU_ARM_TEXT = ("Detection decode + NMS and the match are the ORACLE PORT of the reference (oracle.det.head_decode = nn.py:255-270 in torch, "
                "oracle.det.non_max_suppression = util.py:123-169 with the real torchvision.ops.nms, oracle.match.match_top1 = "
                "F.normalize / F.linear / max as face_recognition/module.py:136-145) because /root/reference does not exist on the GPU box; "
                "crop and DARK decode (~95 % of the time) are the REAL HF VitPoseImageProcessor.preprocess / post_process_pose_estimation calls. "
                "All host threads.")


def make_gallery_on_device(n: int, dev, seed: int = 4242, chunk: int = 131072):
    """[n, 512] unit rows, generated on the GPU in fixed chunks so that every rank of a job builds the same matrix."""
    g = torch.Generator(device=dev).manual_seed(seed)
    out = torch.empty((n, 512), dtype=torch.bfloat16, device=dev)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        x = torch.randn((hi - lo, 512), generator=g, device=dev)
        out[lo:hi] = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    return out


def run_cfg3(args, spp, pipeline, pipe2, dev, rank, world, B, per_frame, barrier, max_over_ranks, peaks, pipe_kw):
    """North-star config 3 at this N: the cfg2 step per GPU (64 frames) against a 1M-id gallery — held locally at N=1,
    row-sharded over the ranks at N>1 with the exchange done by the match kernels over NVLink peer memory (or NCCL with
    --collective nccl).  Reports the whole step and the three phases of the match; at N>1 rank 0 also runs the replicated
    match on the full gallery and compares ids and similarities bit for bit."""
    import copy
    d = spp.dist
    n_ids, m_local = args.cfg3_ids, B * per_frame
    full = make_gallery_on_device(n_ids, dev)                       # 1 GB bf16; every rank builds the same rows
    g = torch.Generator(device=dev).manual_seed(977 + rank)
    ids = torch.randint(0, n_ids, (m_local,), generator=g, device=dev)
    unknown = torch.rand(m_local, generator=g, device=dev) < 0.1
    noise = torch.randn((m_local, 512), generator=g, device=dev) / 512 ** 0.5
    probes = torch.where(unknown[:, None], torch.randn((m_local, 512), generator=g, device=dev), full[ids].float() + 0.3 * noise)
    probes = (probes / probes.norm(dim=1, keepdim=True) * (5.0 + 25.0 * torch.rand((m_local, 1), generator=g, device=dev))).contiguous()
    lo, hi = d.shard_bounds(n_ids, world, rank)
    inp3 = copy.copy(pipe2.inp)                                     # same device-resident frames / head maps / heatmaps
    inp3.embeddings = probes
    matcher = peers = None
    if world > 1:
        shard_rows = full[lo:hi].contiguous()
        if args.collective == "peer":
            peers = d.PeerGroup(m_local)
            matcher = d.PeerShardedMatcher(peers, shard_rows, lo, 0.4)
        else:
            matcher = d.gpu_matcher(shard_rows, lo, 0.4)
        if rank != 0:
            del full
            full = None
            torch.cuda.empty_cache()
    barrier()
    # A 1M-id match is the step's longest kernel (one persistent CTA per SM, 225 KB of shared memory: nothing co-resides), so
    # here the step is a sum of whole-machine kernels and the chains should be SHORT rather than small: three-launch
    # detection chains on the whole machine, crop -> heatmap decode on the main stream, no SM split.
    kw3 = dict(pipe_kw)
    kw3.update(match_sms=0, det_max_candidates=0, det_fused=False, heatmap_first=False)
    pipe3 = pipeline.SelectivePosePipeline(inp3, full if world == 1 else torch.empty((1, 512), dtype=torch.bfloat16), dev,
                                           matcher=matcher, capture_collectives=args.capture_collectives, **kw3)
    st = pipe3.stream
    steps = max(5, min(args.steps, 20))
    for _ in range(3):
        pipe3.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        pipe3.step()
    e1.record(st)
    barrier()
    step_ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    ids_step, sims_step = pipe3.out["ids"].clone(), pipe3.out["sims"].clone()

    # phases of the match alone (eager launches on one stream, events between the phases, ranks aligned by a barrier)
    reps = 5
    ph = {"allgather_us": [], "match_us": [], "allreduce_us": []}
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(reps + 1):
        barrier()
        with torch.cuda.stream(st):
            if world == 1:
                evs[0].record(st); evs[1].record(st)
                spp.ops.match_top1(probes, full, 0.4)
                evs[2].record(st); evs[3].record(st)
            elif args.collective == "peer":
                evs[0].record(st)
                matcher.match(probes, d.STAGE_PUSH | d.STAGE_WAIT)
                evs[1].record(st)
                matcher.match(probes, d.STAGE_SEARCH)
                evs[2].record(st)
                matcher.match(probes, d.STAGE_FINALIZE | d.STAGE_REDUCE)
                evs[3].record(st)
            else:
                import torch.distributed as dist
                evs[0].record(st)
                gath = torch.empty((world * m_local, 512), dtype=probes.dtype, device=dev)
                dist.all_gather_into_tensor(gath, probes)
                evs[1].record(st)
                keys = matcher.local_match(gath)
                evs[2].record(st)
                dist.all_reduce(keys, op=dist.ReduceOp.MAX)
                spp.ops.match_unpack_keys(keys[rank * m_local:(rank + 1) * m_local].contiguous(), 0.4)
                evs[3].record(st)
        st.synchronize()
        ph["allgather_us"].append(1e3 * evs[0].elapsed_time(evs[1]))
        ph["match_us"].append(1e3 * evs[1].elapsed_time(evs[2]))
        ph["allreduce_us"].append(1e3 * evs[2].elapsed_time(evs[3]))
    phase = {k: max_over_ranks(statistics.median(v[1:])) for k, v in ph.items()}

    equal = None
    if world > 1:
        ok = torch.ones(1, device=dev)
        if rank == 0:       # replicated run on rank 0: its own probes against the whole gallery
            ref_ids, ref_sims = spp.ops.match_top1(probes, full, 0.4)
            ok[0] = float(torch.equal(ref_ids, ids_step) and torch.equal(ref_sims, sims_step))
        import torch.distributed as dist
        dist.broadcast(ok, 0)
        equal = bool(ok.item() > 0)
    planted_ok = bool((torch.where(unknown, torch.full_like(ids, -1), ids).int() == ids_step)[~unknown].all().item())
    flops = 2.0 * (world * m_local) * (hi - lo) * 512
    rec = {
        "workload": f"cfg3: {world * B} frames over {world} GPU(s) ({B} per GPU), {n_ids}-id bf16 gallery " +
                    ("held locally" if world == 1 else f"sharded by rows ({hi - lo} ids per GPU), every rank scores all {world * m_local} probes against its shard"),
        "exchange": "none (N=1)" if world == 1 else ("peer: probes and (sim,id) keys written by the match kernels into the peers' buffers over NVLink, 5 kernels captured in the step's CUDA graph, no NCCL call"
                                                       if args.collective == "peer" else "nccl: all_gather_into_tensor + all_reduce(MAX) on packed keys, eager beside the graph"),
        "ms_per_step": round(step_ms, 4), "frames_per_s": round(world * B / (step_ms / 1e3), 1), "steps": steps,
        "match_us": round(phase["match_us"], 1), "allgather_us": round(phase["allgather_us"], 1), "allreduce_us": round(phase["allreduce_us"], 1),
        "phase_note": "match alone, eager, after a barrier: allgather = normalise + push + wait for all ranks' probes; match = tcgen05 GEMM + top-2 "
                      "over the local shard; allreduce = fp32 re-score + key push + wait + max + unpack (N=1: the whole local op is match_us)",
        "match_tflops_per_gpu": round(flops / (phase["match_us"] * 1e-6) / 1e12, 1) if phase["match_us"] > 0 else None,
        "match_frac_bf16_peak": round(flops / (phase["match_us"] * 1e-6) / 1e12 / peaks["bf16"], 3) if phase["match_us"] > 0 else None,
        "sharded_ids_equal": equal, "planted_ids_recovered": planted_ok,
        "limiter": max((("match GEMM", phase["match_us"]), ("probe exchange", phase["allgather_us"]), ("re-score + key reduce", phase["allreduce_us"])),
                       key=lambda kv: kv[1])[0] + f" ({max(phase.values()):.0f} us of the {1e3 * step_ms:.0f} us step)",
    }
    del pipe3
    if peers is not None:
        barrier()
        peers.close()
    torch.cuda.empty_cache()
    return rec


def run_reference(args, wl, rank, world, pipeline, emit):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the same torch /
    torchvision / HF calls the reference makes) on this box's host cores.  Rank 0 only."""
    if rank != 0:
        return
    spp = importlib.import_module(PKG)
    inp = pipeline.synthetic_inputs(wl["batch"], wl["height"], wl["width"], wl["per_frame"], wl["joints"], seed=0)
    ms = spp.synth.make_match_set(wl["batch"] * wl["per_frame"], wl["gallery"], seed=1000)
    inp.embeddings = ms.embeddings
    if args.frames == "u8":
        inp.frames = (inp.frames * 255.0).round().clamp(0, 255).to(torch.uint8)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    n, times = time_cpu_reference(inp, ms.gallery, wl["batch"], wl["per_frame"], budget_s=150.0, reps=steps, warm=warm)
    total = sum(times)
    fps = n * len(times) / total
    line = {
        "impl": "reference",
        "metric": "frames/sec post-backbone selective-pose pipeline (batch 64, 1/2/4/8 B200)",
        "value": round(fps, 3), "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": round(1e3 * total / len(times), 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": shared_config(args, wl, world, False, "none"),
        "impl_detail": {"frames_timed_per_step": n, "crops_timed_per_step": n * wl["per_frame"]},
        "cpu_baseline": {"value": round(fps, 3), "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"each step = the first {n} of {wl['batch']} frames of {args.workload} "
                                   f"({n * wl['per_frame']} crops) through the reference's CPU calls. " + CPU_ARM_TEXT},
        "e2e": {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


if __name__ == "__main__":
    main()
