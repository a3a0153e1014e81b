"""Aggregate warp-stall samples per CUDA source line from
`ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:K --launch-count 1`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.defaultdict(lambda: collections.Counter())
srcline = {}
hdr = None
fname = ""
for r in rows:
    if r and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr):
        key = (fname, r[0])
        if r[1]:
            srcline[key] = r[1]
        def num(x):
            try:
                return int(x)
            except ValueError:
                return 0
        si = hdr.index("Warp Stall Sampling (All Samples)")
        agg[key]["samples"] += num(r[si])
        agg[key]["inst"] += num(r[hdr.index("Instructions Executed")])
        for i, n in enumerate(hdr):
            if n.startswith("stall_") and "Not Issued" not in n:
                agg[key][n] += num(r[i])
tot = sum(v["samples"] for v in agg.values())
print("total samples", tot)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    reasons = sorted(((n, c) for n, c in v.items() if n.startswith("stall_") and c), key=lambda x: -x[1])[:3]
    rs = " ".join(f"{n[6:]}={c}" for n, c in reasons)
    print(f"{v['samples']:5d} {100.0 * v['samples'] / max(tot, 1):5.1f}% inst={v['inst']:8d} {key[0][:18]}:{key[1]:>4s} [{rs}]  {srcline.get(key, '')[:80].strip()}")
