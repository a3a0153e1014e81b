"""torchrun probe: time the sharded-gallery match chain alone and beside the graph (N >= 2)."""
import importlib, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "person-recognition-for-pose-estimation_b200"
spp = importlib.import_module(PKG); pipeline = importlib.import_module(PKG + ".pipeline")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ms = spp.synth.make_match_set(640, 10000, seed=rank)
gal = ms.gallery.to(torch.bfloat16).to(dev)
emb = ms.embeddings.to(dev)
matcher = spp.dist.gpu_matcher(gal, rank * 10000, 0.4)
st = torch.cuda.Stream(dev)
def timed(fn, n=50):
    with torch.cuda.stream(st):
        for _ in range(5): fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(st)
        for _ in range(n): fn()
        e1.record(st); cpu = (time.perf_counter() - t0) / n * 1e6
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, cpu
g = torch.empty(world * 640, 512, device=dev)
keys = torch.zeros(world * 640, dtype=torch.int64, device=dev)
res = {}
res["all_gather"] = timed(lambda: dist.all_gather_into_tensor(g, emb))
res["all_reduce"] = timed(lambda: dist.all_reduce(keys, op=dist.ReduceOp.MAX))
res["local_match"] = timed(lambda: spp.match_top1(g, gal, None, 0, want_keys=True))
res["chain"] = timed(lambda: matcher.match(emb))
if rank == 0:
    for k, (gpu, cpu) in res.items(): print(f"{k:12s} gpu {gpu:8.1f} us/iter   host {cpu:8.1f} us/iter", flush=True)
dist.destroy_process_group()
