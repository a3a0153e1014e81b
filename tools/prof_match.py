"""One large gallery match for ncu captures:  python tools/prof_match.py [M] [N]   (default 5120 x 1 000 000)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
m = int(sys.argv[1]) if len(sys.argv) > 1 else 5120
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
gal = torch.randn(n, 512, device=dev, generator=g)
gal = (gal / gal.norm(dim=1, keepdim=True)).to(torch.bfloat16)
emb = gal[torch.randint(0, n, (m,), device=dev, generator=g)].float() + 0.02 * torch.randn(m, 512, device=dev, generator=g)
for _ in range(3):
    ids, sims = spp.match_top1(emb, gal, 0.4)
torch.cuda.synchronize()
print("ok", ids.shape, float(sims.mean()))
