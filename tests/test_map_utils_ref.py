"""The synthetic orchard and the row leave / enter poses (path_planner/utils/map_utils.py:45-61, :228-271) are what every
scenario of the benchmark is built from.  The reference's own map_utils.py runs here unmodified (shapely / matplotlib
stubbed: `box` and the plot helper are never called by these two functions); the product mirror and the oracle's
restatement must give the same bits.  Golden: tests/golden/map_utils_golden.npz (generator: the live test's own case
list, written by `python tests/test_map_utils_ref.py`)."""
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import planner as OP                                            # noqa: E402
from oracle import ref_loader                                               # noqa: E402
from headland_trajectory_planning_b200.utils import map_utils as MU         # noqa: E402

GOLD = os.path.join(HERE, "golden", "map_utils_golden.npz")


def cases():
    rng = np.random.default_rng(5)
    out = []
    for k in range(40):
        row_num = int(rng.integers(3, 10))
        lengths = float(rng.uniform(10, 30)) if k % 3 else [float(v) for v in rng.uniform(10, 30, row_num)]
        out.append(dict(seed=int(rng.integers(0, 1 << 30)), row_num=row_num, row_width=float(rng.uniform(2.2, 4.0)),
                        lengths=lengths, slope=float(rng.uniform(0.0, math.radians(15))), l_std=(0.0, 0.5, 1.0)[k % 3],
                        offset=float(rng.uniform(0.0, 3.0))))
    out.append(dict(seed=1, row_num=8, row_width=2.5, lengths=20, slope=math.radians(10), l_std=1.0, offset=0.0))   # the notebooks' orchard
    return out


def run(mod):
    """rows of every case, then every (row, side, pose type) base pose, flattened."""
    rows_all, poses_all = [], []
    for c in cases():
        np.random.seed(c["seed"])
        rows = mod.create_tree_rows(c["row_num"], c["row_width"], c["lengths"], slope_angle=c["slope"], l_std=c["l_std"])
        rows_all.append(np.asarray(rows).reshape(-1))
        for r in range(c["row_num"] - 1):
            for side in (mod.NEAR_SIDE, mod.FAR_SIDE):
                for kind in (mod.LEAVE_POSE, mod.ENTER_POSE):
                    poses_all.append(np.asarray(mod.get_base_pose(r, rows, c["offset"], side=side, pose_type=kind), dtype=np.float64))
    return np.concatenate(rows_all), np.array(poses_all)


@pytest.mark.parametrize("mod", [MU, OP], ids=["mirror", "oracle"])
def test_rows_and_base_poses_equal_reference_golden(mod):
    g = np.load(GOLD)
    rows, poses = run(mod)
    assert np.array_equal(rows, g["rows"])
    assert np.array_equal(poses, g["poses"])                       # bit for bit: same numpy calls in the same order


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_live_reference_map_utils_equals_golden():
    ref = ref_loader.load_planner("map_utils")
    g = np.load(GOLD)
    rows, poses = run(ref)
    assert np.array_equal(rows, g["rows"]) and np.array_equal(poses, g["poses"])


if __name__ == "__main__":                                         # regenerate the golden from the reference itself
    ref = ref_loader.load_planner("map_utils")
    rows, poses = run(ref)
    np.savez_compressed(GOLD, rows=rows, poses=poses)
    print("map_utils golden:", rows.shape, poses.shape)
