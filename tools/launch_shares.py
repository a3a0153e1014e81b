"""Per-kernel count / mean duration / share of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file F`):
python tools/launch_shares.py profiles/r1_e_launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.defaultdict(list)
for r in rows:
    if r and r[0] == "ID":
        hdr = r
    elif hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
            agg[d["Kernel Name"][:70]].append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} n={len(v):4d} mean={sum(v) / len(v):8.1f} us  share={100 * sum(v) / tot:5.1f}%")
