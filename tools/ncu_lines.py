#!/usr/bin/env python
"""Top source lines by warp-stall samples from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K --launch-count 1`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = None
hdr = None
data = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            s = int(d["# Samples"])
        except Exception:
            continue
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        data.append((s, fname, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"] or 0), stalls))
tot = sum(d[0] for d in data) or 1
print("total samples", tot)
agg = {}
for d in data:
    for k, v in d[5].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for d in sorted(data, reverse=True)[:top]:
    st = ",".join(f"{k}:{v}" for k, v in sorted(d[5].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * d[0] / tot:5.1f}% {d[1]}:{d[2]:<4} inst={d[4]:<9} [{st}] {d[3]}")
