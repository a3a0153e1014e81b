"""Print the hottest SASS lines (stall samples) of one kernel from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) == len(rows[h]) and r[0].startswith("0x")]
si, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot, "warp-instr", sum(int(r[ie] or 0) for r in data), "sass lines", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:top]
for i in sorted(order):
    print(f"{i:5d} {int(data[i][si] or 0):6d} {100.0 * int(data[i][si] or 0) / max(tot, 1):5.1f}% {data[i][ie]:>8s}  {data[i][1][:100]}")
