#!/usr/bin/env python
"""Code footprint of a kernel per device FUNCTION: joins the SASS page of an ncu capture (executed count and stall
samples per instruction) with `nvdisasm -g` of the same cubin (which out-of-line device function an instruction
belongs to), and reports per function the 128-byte instruction lines that hold an instruction executed at least
`thr` times (for K4: about once per two pops), its share of the executed warp instructions and of the samples.

  cuobjdump -xelf all headland_trajectory_planning_b200/csrc/hl_astar.o && nvdisasm -g -c hl_astar.sm_100a.cubin > astar.sass
  ncu -i k4.ncu-rep --page source --print-source cuda,sass --csv > k4_src.csv
  tools/ncu_code_footprint.py k4_src.csv astar.sass _Z16k_hybrid_astar_s [thr=100000]
The ncu page's own per-file sections mis-attribute inlined lines, hence the join on the instruction offset."""
import collections
import csv
import re
import sys

src_csv, sass_path, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 100e3

# ---- executed count / samples per instruction address (ncu)
rows = list(csv.reader(open(src_csv)))
hdr = None
ins = {}
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and r[0] == "" and len(r) > ii and r[2].startswith("0x") and r[ii] not in ("-", ""):
        a, e = int(r[2], 16), int(r[ii])
        s = int(r[si]) if r[si].isdigit() else 0
        if a not in ins or e > ins[a][0]:
            ins[a] = (e, s)
base = min(ins)

# ---- function of every instruction offset (nvdisasm): sub-function labels look like $<kernel>$<mangled callee>:
lines = open(sass_path).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kernel))
end = next((i for i, l in enumerate(lines) if i > start and l.startswith(".text.")), len(lines))
fn = "kernel body"
owner = {}
for l in lines[start:end]:
    m = re.match(r"^\$?(\S+):$", l)
    if m and not l.startswith(".L_x") and not l.startswith("\t"):
        name = m.group(1)
        if name.startswith(kernel):
            callee = name.split("$")[-1] if "$" in name else ""
            m2 = re.search(r"(\d+)([A-Za-z_]\w*)", callee[callee.rfind("_INTERNAL"):] if "_INTERNAL" in callee else callee)
            if "_INTERNAL" in callee:              # _ZN42_INTERNAL_<hash>_<file><len><name>E...: the last <len><name> pair
                m3 = re.search(r"_cu_[0-9a-f]{8}(\d+)(.*)", callee)
                fn = m3.group(2)[:int(m3.group(1))] if m3 else callee
            elif callee.startswith("_Z"):
                m4 = re.match(r"_Z(\d+)(.*)", callee)
                fn = m4.group(2)[:int(m4.group(1))] if m4 else callee
            else:
                fn = callee or "kernel body"
        elif name.startswith("__internal") or name.startswith("_internal"):
            fn = name
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+", l)
    if m:
        owner[int(m.group(1), 16)] = fn

hot = collections.Counter()
size = collections.Counter()
dyn = collections.Counter()
smp = collections.Counter()
by_line = collections.defaultdict(list)
for a, (e, s) in ins.items():
    o = a - base
    f = owner.get(o, "?")
    size[f] += 1
    dyn[f] += e
    smp[f] += s
    by_line[o // 128].append((e, f))
for vs in by_line.values():
    if max(v[0] for v in vs) >= thr:
        hot[collections.Counter(v[1] for v in vs).most_common(1)[0][0]] += 1
td, ts = sum(dyn.values()) or 1, sum(smp.values()) or 1
print(f"{len(ins)} instructions ({len(ins) * 16 / 1024:.1f} KB); lines with an instruction executed >= {thr:.0f} times: "
      f"{sum(hot.values()) * 128 / 1024:.1f} KB")
print(f"{'function':36s} hot KB  size KB  executed %  samples %")
for f in sorted(size, key=lambda k: -dyn[k]):
    if dyn[f]:
        print(f"{f[:36]:36s} {hot[f] * 128 / 1024:6.1f}  {size[f] * 16 / 1024:7.1f}  {100 * dyn[f] / td:10.1f}  {100 * smp[f] / ts:9.1f}")
