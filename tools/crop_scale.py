import importlib, os, sys, torch
sys.path.insert(0, "/root/repo")
spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
dev = torch.device("cuda:0")
for nf in (16, 32, 64, 128, 256):
    cs = spp.synth.make_crop_set(nf, 720, 1280, per_frame=10, seed=2, smooth=False)
    frames, boxes, idx = cs.frames.to(dev), cs.boxes.to(dev), cs.frame_idx.to(dev)
    out = spp.crop_affine(frames, boxes, idx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        spp.crop_affine(frames, boxes, idx, out=out)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 50
    print(f"frames={nf} crops={nf*10}: {t:8.1f} us  per 640 crops {t*64/nf:7.1f} us", flush=True)
    del frames, out
