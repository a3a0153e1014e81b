"""Microbench of the crop kernel at cfg2 (640 crops from 64 frames 1280x720): python tools/micro_crop.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
dev = torch.device("cuda:0")
cs = spp.synth.make_crop_set(64, 720, 1280, per_frame=int(os.environ.get("MICRO_CROP_PER_FRAME", "10")), seed=2, smooth=False)
frames, boxes, idx = cs.frames.to(dev), cs.boxes.to(dev), cs.frame_idx.to(dev)
out = spp.crop_affine(frames, boxes, idx)
for dtype in ("f32", "u8"):
    fr = frames if dtype == "f32" else (frames * 255).round().to(torch.uint8)
    kw = {} if dtype == "f32" else {"mean": [m * 255 for m in (0.485, 0.456, 0.406)], "std": [s * 255 for s in (0.229, 0.224, 0.225)]}
    spp.crop_affine(fr, boxes, idx, out=out, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        spp.crop_affine(fr, boxes, idx, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"stage_kb={os.environ.get('SPP_CROP_STAGE_KB','40')} frames={dtype}: {e0.elapsed_time(e1) * 50:8.1f} us per launch", flush=True)
