"""Host -> device copy ceiling of the box with N GPUs copying at once: every rank copies a pinned
1 GiB buffer to its GPU `reps` times on `streams` streams; prints per-rank and aggregate GB/s.  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_ceiling.py
This is what bounds bench.py's end-to-end number at N=8 (1.6 GB of backbone outputs per GPU and step)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
out = {}
for streams in (1, 2, 4):
    n = 1 << 28                                   # 1 GiB of fp32
    host = torch.empty(n, dtype=torch.float32).pin_memory()
    host.fill_(1.0)
    devbuf = torch.empty(n, dtype=torch.float32, device=dev)
    ss = [torch.cuda.Stream(dev) for _ in range(streams)]
    chunk = n // streams
    def run():
        for j, s in enumerate(ss):
            with torch.cuda.stream(s):
                devbuf[j * chunk:(j + 1) * chunk].copy_(host[j * chunk:(j + 1) * chunk], non_blocking=True)
    run(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 8
    for _ in range(reps):
        run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([reps * n * 4 / dt / 1e9], device=dev, dtype=torch.float64)
    if world > 1:
        all_ = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(all_, gbs)
        vals = [float(x) for x in all_]
    else:
        vals = [float(gbs)]
    out[f"streams_{streams}"] = {"per_gpu_gbs": [round(v, 1) for v in vals], "sum_gbs": round(sum(vals), 1), "min_gbs": round(min(vals), 1)}
    del host, devbuf
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_pinned_1GiB": out, "cpus": os.cpu_count(),
                      "numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])}))
if world > 1:
    dist.destroy_process_group()
