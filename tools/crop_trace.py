"""Event trace of one CTA of the persistent crop kernel: when the producer issues each band copy and when the consumers wait
for / get each band and start / finish each item.  Needs the trace build of the library:
    make -C person-recognition-for-pose-estimation_b200/csrc clean all EXTRA=-DSPP_CROP_TRACE && python tools/crop_trace.py
(rebuild without EXTRA afterwards: the hooks are compiled out of the shipped library)."""
import importlib, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
dev = torch.device("cuda:0")
cs = spp.synth.make_crop_set(64, 720, 1280, per_frame=10, seed=2, smooth=False)
frames, boxes, idx = cs.frames.to(dev), cs.boxes.to(dev), cs.frame_idx.to(dev)
out = spp.crop_affine(frames, boxes, idx)
tr = torch.zeros(1 + 2 * 4000, dtype=torch.int64, device=dev)
for _ in range(3):
    spp.crop_affine(frames, boxes, idx, out=out)
torch.cuda.synchronize()
os.environ["SPP_CROP_TRACE_PTR"] = str(tr.data_ptr())
spp.crop_affine(frames, boxes, idx, out=out)
torch.cuda.synchronize()
t = tr.cpu().numpy()
n = int(t[0]); ev = t[1:1 + 2 * min(n, 4000)].reshape(-1, 2)
ev = ev[np.argsort(ev[:, 1])]
t0 = ev[0, 1]
names = {1: "P issue band", 2: "P bands done", 10: "C item start", 11: "C wait full", 12: "C got band", 13: "C item done"}
for code, ts in ev[:160]:
    ty, k, b = code >> 48, (code >> 24) & 0xffffff, code & 0xffffff
    print(f"{(ts - t0) / 1e3:8.2f} us  {names.get(ty, ty):24s} item {k} band/kind {b}")
