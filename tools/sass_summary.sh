#!/bin/bash
# Per-kernel SASS evidence: which Blackwell instruction families each kernel of libspp.so contains.
#   bash tools/sass_summary.sh > profiles/sass_summary.txt
LIB=person-recognition-for-pose-estimation_b200/libspp.so
echo "# cuobjdump -sass $LIB | per-function instruction counts (sm_100a)"
echo "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load, UBLKCP = bulk TMA (cp.async.bulk),"
echo "# SYNCS = mbarrier ops, UTCATOM/UTCCP etc. as emitted; HMMA/IMMA/wgmma must be absent."
cuobjdump -sass "$LIB" | awk '
/Function : / { fn=$3; sub(/^_ZN3spp[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_/, "", fn); names[fn]=1; next }
fn != "" {
  total[fn]++
  if ($0 ~ /UTCHMMA/) a[fn,"UTCHMMA"]++
  if ($0 ~ /LDTM/) a[fn,"LDTM"]++
  if ($0 ~ /UTCBAR/) a[fn,"UTCBAR"]++
  if ($0 ~ /UTMALDG/) a[fn,"UTMALDG"]++
  if ($0 ~ /UBLKCP/) a[fn,"UBLKCP"]++
  if ($0 ~ /SYNCS/) a[fn,"SYNCS"]++
  if ($0 ~ /(^|[^C])HMMA|IMMA|WGMMA|HGMMA/) a[fn,"legacy_mma"]++
  if ($0 ~ /VOTE|VOTEU/) a[fn,"VOTE"]++
  if ($0 ~ /SHFL/) a[fn,"SHFL"]++
  if ($0 ~ /DFMA|DADD|DMUL/) a[fn,"FP64"]++
}
END {
  printf "%-9s %-8s %-6s %-7s %-8s %-7s %-6s %-5s %-5s %-5s %-10s  %s\n", "lines", "UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "VOTE", "SHFL", "FP64", "legacy_mma", "kernel"
  for (f in names) printf "%-9d %-8d %-6d %-7d %-8d %-7d %-6d %-5d %-5d %-5d %-10d  %s\n", total[f], a[f,"UTCHMMA"], a[f,"LDTM"], a[f,"UTCBAR"], a[f,"UTMALDG"], a[f,"UBLKCP"], a[f,"SYNCS"], a[f,"VOTE"], a[f,"SHFL"], a[f,"FP64"], a[f,"legacy_mma"], f
}' | (read -r hdr; echo "$hdr"; sort -k12)
