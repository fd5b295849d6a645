#!/usr/bin/env python
"""Hot instruction footprint of a kernel from an ncu source-page csv: SASS instructions executed at least
`frac` x the most-executed one, in KB (16 B per instruction) -- to compare with the 32 KB L1.5 instruction cache."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]
ii = hdr.index('Instructions Executed')
ex = {}
for r in rows[3:]:
    if len(r) > ii and r[0] == '' and r[2].startswith('0x') and r[ii] not in ('-', ''):
        ex[r[2]] = max(ex.get(r[2], 0), int(r[ii]))
mx = max(ex.values())
for frac in (0.5, 0.1, 0.02, 0.005, 0.0):
    n = sum(1 for v in ex.values() if v > frac * mx or (frac == 0.0 and v > 0))
    print(f"executed > {frac:5.3f} x max ({mx}): {n} instructions = {n * 16 / 1024:.1f} KB")
print(f"all SASS of the kernel: {len(ex) * 16 / 1024:.1f} KB")
