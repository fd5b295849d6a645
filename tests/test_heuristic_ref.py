"""Guide polyline, per-segment search lengths and calculate_state_cost of ReferenceLineHeuristic
(path_planner/reference_line_heuristic.py:50-96, :131-158) pinned on the reference's OWN class, run unmodified with
shapely reduced to inert stand-ins (these three are numpy only; the lane predicates stay unpinned).  The oracle's
restatement and the product's host mirror must give the same bits; the device's state cost is compared with the
oracle's in tests/test_astar_gpu.py.  Golden: tests/golden/heuristic_ref_golden.npz (written by running this file)."""
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import planner as OP                                            # noqa: E402
from oracle import ref_loader                                               # noqa: E402

GOLD = os.path.join(HERE, "golden", "heuristic_ref_golden.npz")


def cases():
    rng = np.random.default_rng(11)
    out = []
    for k in range(24):
        n = int(rng.integers(2, 8))                                 # 1 .. 6 segments (the long-search rule needs > 4)
        pts = np.cumsum(rng.uniform(-1, 1, (n, 2)) * rng.uniform(2, 12), axis=0) + rng.uniform(-20, 20, 2)
        poses = np.column_stack([pts[rng.integers(0, n, 40)] + rng.normal(0, 1.5, (40, 2)), rng.uniform(-math.pi, math.pi, 40)])
        poses[:4, :2] += 30.0                                       # far from the guide: the 100-cost branch
        out.append((pts, np.array([pts[-1, 0], pts[-1, 1], 0.0]), poses))
    return out


def run(cls, car, with_costs=True):
    guide, lengths, costs = [], [], []
    for pts, goal, poses in cases():
        h = cls(pts, goal, car)
        guide.append(np.asarray(h.guided_path, dtype=np.float64).reshape(-1))
        lengths.append(np.asarray(h.search_lengths, dtype=np.float64))
        if with_costs:
            costs.append(np.array([h.calculate_state_cost(p) for p in poses]))
    return np.concatenate(guide), np.concatenate(lengths), (np.concatenate(costs) if with_costs else None)


def _check(got):
    g = np.load(GOLD)
    for a, k in zip(got, ("guide", "lengths", "costs")):
        if a is not None:
            assert np.array_equal(a, g[k]), k                      # bit for bit: the same numpy calls in the same order


def test_oracle_heuristic_equals_reference_golden():
    _check(run(OP.ReferenceLineHeuristic, OP.CarModel(max_steer=0.55, axle_to_back=0.55, width=1.48)))


def test_mirror_heuristic_equals_reference_golden():
    from headland_trajectory_planning_b200.car_model import CarModel
    from headland_trajectory_planning_b200.reference_line_heuristic import ReferenceLineHeuristic
    # the mirror builds the guide and the lengths on the host and leaves calculate_state_cost to the search kernel
    _check(run(ReferenceLineHeuristic, CarModel(max_steer=0.55, axle_to_back=0.55, width=1.48), with_costs=False))


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_live_reference_heuristic_equals_golden():
    ref = ref_loader.load_reference_line_heuristic()
    _check(run(ref.ReferenceLineHeuristic, None))


if __name__ == "__main__":
    ref = ref_loader.load_reference_line_heuristic()
    guide, lengths, costs = run(ref.ReferenceLineHeuristic, None)
    np.savez_compressed(GOLD, guide=guide, lengths=lengths, costs=costs)
    print("heuristic golden:", guide.shape, lengths.shape, costs.shape)
