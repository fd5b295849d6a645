"""Mirror of ``path_planner/utils/reeds_shepp.py``'s public entry points
(``calc_all_paths``, ``calc_optimal_path``, ``PATH``) backed by the CUDA word
evaluator (``hl_rs_all_paths``) and sampler (``hl_rs_sample``).  Single queries are
batches of one; sweeps should call ``ops.rs_all_paths`` directly."""
import numpy as np

from .. import ops

STEP_SIZE = 0.2
MAX_LENGTH = 1000.0
_LETTERS = "SLR"
# letters of the 46 candidate rows in evaluation order (reeds_shepp.py:131-141,
# 163-236, 286-319, 353-422, 443-468)
_ROW_LETTERS = (["SLS", "SRS"]
                + ["LSL", "LSL", "RSR", "RSR"] + ["LSR", "LSR", "RSL", "RSL"]
                + ["LRL", "LRL", "RLR", "RLR"] * 2
                + ["LRLR", "LRLR", "RLRL", "RLRL"] * 2
                + ["LRSL", "LRSL", "RLSR", "RLSR"] + ["LRSR", "LRSR", "RLSL", "RLSL"]
                + ["LSRL", "LSRL", "RSLR", "RSLR"] + ["RSRL", "RSRL", "LSLR", "LSLR"]
                + ["LRSLR", "LRSLR", "RLSRL", "RLSRL"])
assert len(_ROW_LETTERS) == 46


class PATH:
    def __init__(self, lengths, ctypes, L, x, y, yaw, cs, directions):
        self.lengths = lengths
        self.ctypes = ctypes
        self.L = L
        self.x = x
        self.y = y
        self.yaw = yaw
        self.directions = directions
        self.cs = cs


def row_letters(cand):
    return list(_ROW_LETTERS[cand])


def calc_all_paths(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=STEP_SIZE):
    sg = np.array([[sx, sy, syaw, gx, gy, gyaw]], dtype=np.float64)
    words, count, _ = ops.rs_all_paths(sg, maxc, step_size, want_order=False)
    n = int(count.cpu().numpy()[0])
    if n < 0:
        raise AssertionError("path.L >= 0.01")       # reeds_shepp.py:84
    w = ops.rs_words_to_host(words)[0, :n]
    if n == 0:
        return []
    off, x, y, yaw, cs, dr = ops.rs_sample(np.repeat(sg[:, :3], n, axis=0), w, maxc, step_size)
    paths = []
    for k in range(n):
        a, b = off[k], off[k + 1]
        ns = int(w["n_seg"][k])
        paths.append(PATH([float(v) for v in w["len"][k, :ns]], row_letters(int(w["cand"][k])),
                          float(w["L"][k]), x[a:b].tolist(), y[a:b].tolist(), yaw[a:b].tolist(),
                          [c if c != 0 else 0 for c in cs[a:b].tolist()], [int(d) for d in dr[a:b]]))
    return paths


def calc_optimal_path(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=STEP_SIZE):
    """reeds_shepp.py:26-36 (``<=``: the LAST of equally short paths wins)."""
    paths = calc_all_paths(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=step_size)
    min_l, mini = paths[0].L, 0
    for i, p in enumerate(paths):
        if p.L <= min_l:
            min_l, mini = p.L, i
    return paths[mini]
