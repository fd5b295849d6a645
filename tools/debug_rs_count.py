import os, sys, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import rs_port
from headland_trajectory_planning_b200 import ops
import test_offset_golden as T
np.set_printoptions(precision=17)
k = 0
for sp, r0, r1, side, start, end, leave, enter in T._start_end_cases():
    car_curv = np.tan(0.55) / 1.9
    maxc = 1.0 / (1.0 / car_curv)
    ref = rs_port.calc_all_paths(start[0], start[1], start[2], end[0], end[1], end[2], maxc, 0.1)
    sg = np.array([[*start, *end]])
    words, count, _ = ops.rs_all_paths(sg, maxc, 0.1, want_order=False)
    w = ops.rs_words_to_host(words)
    for j, p in enumerate(ref):
        if w["npts"][0, j] != len(p.x):
            print("case", k, "start", start, "end", end, "word", j, "cand", p.cand, "".join(p.ctypes), "gpu npts", w["npts"][0, j], "port", len(p.x))
            print("   lens port", p.lengths, "gpu", w["len"][0, j, :len(p.lengths)], "nlen", w["nlen"][0, j, :len(p.lengths)])
            print("   lens/step*maxc", [l * maxc / (0.1 * maxc) for l in p.lengths])
    k += 1
print("cases", k)
