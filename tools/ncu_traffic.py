"""profiles/ncu_traffic.json from the per-kernel summary of a full-set ncu capture of one eager step:
    python tools/ncu_summary.py step.ncu-rep > profiles/rX_full_summary.txt && python tools/ncu_traffic.py profiles/rX_full_summary.txt "<how it was captured>"
"""
import json
import os
import re
import sys

txt = open(sys.argv[1]).read()
how = sys.argv[2] if len(sys.argv) > 2 else ""
mul = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
stages, dets = {}, []
for b in txt.split("----- ")[1:]:
    name = b.splitlines()[0]
    tot = 0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = re.search(m.replace(".", r"\.") + r"\s+([\d.]+)\s+(\w+)", b).groups()
        tot += int(float(v) * mul[u])
    if "nms_kernel<2" in name or "cand_" in name or "nms_kernel<1" in name:
        dets.append(tot)
    elif "heatmap" in name:
        stages["heatmap_decode"] = stages.get("heatmap_decode", 0) + tot
    elif "crop" in name:
        stages["crop_affine"] = stages.get("crop_affine", 0) + tot
    else:
        stages["match_top1"] = stages.get("match_top1", 0) + tot
assert len(dets) == 2, "expected the two fused detection kernels"
stages["decode_nms_face"], stages["decode_nms_person"] = dets
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch, from " + how + " (summary: " + os.path.basename(sys.argv[1]) +
       "), cfg2, one B200; per stage = sum over its kernels. Static: bench.py reads this file, it does not measure traffic in the run.",
       "cfg2": {k: stages[k] for k in ("crop_affine", "heatmap_decode", "decode_nms_face", "decode_nms_person", "match_top1")}}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
open(path, "w").write(json.dumps(out, indent=2) + "\n")
print(out["cfg2"], sum(out["cfg2"].values()))
