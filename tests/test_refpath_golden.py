"""SURVEY.md 8(f) rank 4 -- warm-start path -> OBCA initial guess (obca_py/util.py:62-113, cubic_spline.py:19-112).

``tests/golden/refpath_golden.npz`` = the REFERENCE's own ``get_init_ref_path`` (imported through
``oracle.ref_loader.load_obca_util``, generator ``oracle/gen_golden.py refpath``) on the golden planner paths plus
synthetic paths with several direction changes, repeated poses and 2- / 3-pose pieces.  The oracle restatement (scipy
CubicSpline like the reference) must reproduce it bit for bit; K8 solves the not-a-knot systems itself (Thomas
recurrence instead of LAPACK's pivoted banded solve) and must agree to 1e-9 on x / y / v / yaw and 1e-7 on steer
(second derivatives) -- far inside the 1e-3 tier the OBCA consumer is judged at."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import obca_util as OU  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refpath_golden.npz")


def _cases():
    g = np.load(GOLD)
    po = np.concatenate([[0], np.cumsum(g["path_len"])])
    to = np.concatenate([[0], np.cumsum(g["traj_len"])])
    return [(g["path"][po[i]:po[i + 1]], g["traj"][to[i]:to[i + 1]]) for i in range(len(g["path_len"]))]


def test_oracle_matches_reference_golden():
    n_raise = 0
    for p, want in _cases():
        try:
            got = OU.get_init_ref_path(1.9, p[:, 0], p[:, 1], p[:, 2], p[:, 3], p[:, 4])
        except ValueError:
            got = np.zeros((0, 5))
            n_raise += 1
        assert got.shape == want.shape and np.array_equal(got, want)
    assert n_raise >= 1


@pytest.mark.gpu
def test_gpu_ref_paths_match_reference_golden(built_library):
    from headland_trajectory_planning_b200 import ops
    cases = _cases()
    x = np.concatenate([p[:, 0] for p, _ in cases]); y = np.concatenate([p[:, 1] for p, _ in cases])
    d = np.concatenate([p[:, 4] for p, _ in cases]).astype(np.int8)
    off = np.concatenate([[0], np.cumsum([len(p) for p, _ in cases])])
    traj, out_off, status = ops.ref_path_batch(x, y, d, off, 1.9, 0.5, 0.1)
    traj = traj.cpu().numpy()
    worst = np.zeros(5)
    for i, (p, want) in enumerate(cases):
        got = traj[out_off[i]:out_off[i + 1]]
        assert (status[i] == 1) == (len(want) == 0), i
        assert got.shape == want.shape, (i, got.shape, want.shape)
        if len(want):
            worst = np.maximum(worst, np.abs(got - want).max(axis=0))
    assert (worst[:4] < 1e-9).all() and worst[4] < 1e-7, worst


@pytest.mark.gpu
def test_gpu_single_path_mirror(built_library):
    from headland_trajectory_planning_b200 import obca_util as GU
    from headland_trajectory_planning_b200.car_model import CarModel
    car = CarModel()
    p, want = next((p, w) for p, w in _cases() if len(w) and (np.diff(p[:, 4]) != 0).any())
    got = GU.get_init_ref_path(car, p[:, 0], p[:, 1], p[:, 2], p[:, 3], p[:, 4])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-7)
    bad = next(p for p, w in _cases() if len(w) == 0)
    with pytest.raises(ValueError):
        GU.get_init_ref_path(car, bad[:, 0], bad[:, 1], bad[:, 2], bad[:, 3], bad[:, 4])


@pytest.mark.gpu
def test_gpu_batch_from_search_output(built_library):
    """A sweep's paths go from hl_hybrid_astar_batch to OBCA-ready rows without leaving the GPU; every scenario's rows
    equal the oracle's get_init_ref_path of the same path."""
    from headland_trajectory_planning_b200 import obca_util as GU, ops, scenarios as SC, sweep
    from headland_trajectory_planning_b200.env_batch import EnvBatch
    from headland_trajectory_planning_b200.hybrid_a_star_search import unpack_path
    g = np.load(os.path.join(os.path.dirname(GOLD), "astar_golden.npz"))
    n = 64
    specs = [SC.scenario_spec(int(i)) for i in g["index"][:n]]
    scns = [SC.finalize(sp, f) for sp, f in zip(specs, g["feas"][:n])]
    recs, scen, car = sweep.build_records(scns)
    out = ops.hybrid_astar_batch(EnvBatch(recs), scen, sweep.search_params(car), path_capacity=2048 * n, to_host=False)
    traj, off, status = GU.get_init_ref_path_batch(out, car.WHEEL_BASE)
    traj = traj.cpu().numpy()
    host = dict(out)
    from headland_trajectory_planning_b200 import _lib
    host["results"] = out["results"].cpu().numpy().view(_lib.RESULT_DTYPE)
    for k in ("x", "y", "yaw", "k", "dir"):
        host[k] = out[k].cpu().numpy()
    checked = 0
    for i in range(n):
        if host["results"]["path_len"][i] < 2:
            assert off[i + 1] == off[i]
            continue
        x, y, yaw, dirs, ks = unpack_path(host, i)
        try:
            want = OU.get_init_ref_path(car.WHEEL_BASE, np.asarray(x), np.asarray(y), np.asarray(yaw), np.asarray(ks),
                                        np.asarray(dirs, dtype=np.float64))
        except ValueError:
            assert status[i] == 1
            continue
        got = traj[off[i]:off[i + 1]]
        assert got.shape == want.shape
        assert np.abs(got[:, :4] - want[:, :4]).max() < 1e-9 and np.abs(got[:, 4] - want[:, 4]).max() < 1e-7
        checked += 1
    assert checked >= 40
