"""path_planner/utils/path_utils.py (calculate_path_length, get_projection_point, angle_wrap) of the reference run
unmodified against the product mirror and the oracle's restatement: same bits on seeded inputs (live when the
reference checkout is present; the golden travels)."""
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import planner as OP                                            # noqa: E402
from oracle import ref_loader                                               # noqa: E402
from headland_trajectory_planning_b200.utils import path_utils as PU        # noqa: E402

GOLD = os.path.join(HERE, "golden", "path_utils_golden.npz")


def run(mod):
    rng = np.random.default_rng(3)
    lengths, wraps, projs = [], [], []
    for n in (2, 3, 17, 200):
        xs, ys = np.cumsum(rng.normal(0, 0.3, n)), np.cumsum(rng.normal(0, 0.3, n))
        lengths.append(float(mod.calculate_path_length(xs, ys)))
    a = np.concatenate([rng.uniform(-20, 20, 500), [0.0, math.pi, -math.pi, 2 * math.pi, -2 * math.pi, 3 * math.pi, 1e-300, -1e-300]])
    wraps.append(np.asarray(mod.angle_wrap(a), dtype=np.float64))
    wraps.append(np.array([mod.angle_wrap(float(v)) for v in a[:64]], dtype=np.float64))          # the scalar form the search uses
    for _ in range(64 if hasattr(mod, "get_projection_point") else 0):      # not on the hot path: the oracle has none
        v = rng.uniform(-10, 10, 6)
        pp, yaw = mod.get_projection_point(v[0], v[1], v[2] % 3.0, v[3] * 0.05, v[4], v[5])
        projs.append(np.array([pp[0], pp[1], yaw], dtype=np.float64))
    return np.array(lengths), np.concatenate(wraps), (np.concatenate(projs) if projs else None)


@pytest.mark.parametrize("mod", [PU, OP], ids=["mirror", "oracle"])
def test_equal_reference_golden(mod):
    g = np.load(GOLD)
    for a, k in zip(run(mod), ("lengths", "wraps", "projs")):
        if a is not None:
            assert np.array_equal(a, g[k]), k


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_live_reference_equals_golden():
    g = np.load(GOLD)
    for a, k in zip(run(ref_loader.load("path_utils")), ("lengths", "wraps", "projs")):
        assert np.array_equal(a, g[k]), k


if __name__ == "__main__":
    lengths, wraps, projs = run(ref_loader.load("path_utils"))
    np.savez_compressed(GOLD, lengths=lengths, wraps=wraps, projs=projs)
    print("path_utils golden:", lengths.shape, wraps.shape, projs.shape)
