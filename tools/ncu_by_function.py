#!/usr/bin/env python
"""Aggregate the source page of an ncu capture per FUNCTION of the source file: warp-level instructions executed and
stall samples, so that a pipeline kernel's stages can be ranked.
usage: ncu_by_function.py <source-page csv> <source file .cu/.cuh> [more source files]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
srcs = sys.argv[2:]
# function start lines per source file (crude: lines that look like a definition at column 0)
funcs = {}
for path in srcs:
    name = path.split("/")[-1]
    starts = []
    for ln, text in enumerate(open(path), 1):
        m = re.match(r"^(?:static\s+)?(?:__device__|__global__|template|extern \"C\").*?([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
        if m and not text.startswith(" "):
            starts.append((ln, m.group(1)))
    funcs[name] = starts


def owner(fname, line):
    best = "?"
    for ln, fn in funcs.get(fname, []):
        if ln <= line:
            best = fn
        else:
            break
    return f"{fname}:{best}"


agg = {}
fname = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            s = int(d["# Samples"]); ins = int(d["Instructions Executed"] or 0)
        except Exception:
            continue
        k = owner(fname, int(r[0]))
        a = agg.setdefault(k, [0, 0])
        a[0] += s; a[1] += ins
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{100 * a[0] / ts:5.1f}% samples  {100 * a[1] / ti:5.1f}% inst  {a[1]:>11}  {k}")
