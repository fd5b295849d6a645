"""SURVEY.md 8(f) rank 2 (first part) -- the offset-pose sweep (safety_forward_path_plan.py:248-283).

``tests/golden/offset_golden.npz`` holds the results of the REFERENCE's own ``get_offset_pose`` (imported
unmodified through ``oracle.ref_loader.load_planner`` together with the reference's
``CarModel.calculate_motion_path``; generator ``oracle/gen_golden.py offset``).  The oracle restatement must
reproduce them bit for bit; the CUDA path (``hl_arc_paths`` + K1 + reduce) must pick the same offset and give
the same arc within the north_star pose tolerance."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import planner as OP  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "offset_golden.npz")


def _cases():
    from headland_trajectory_planning_b200 import scenarios as SC
    g = np.load(GOLD)
    po = np.concatenate([[0], np.cumsum(g["path_len"])])
    for k in range(len(g["index"])):
        if g["pose_type"][k] >= 100:
            continue
        yield (k, SC.scenario_spec(int(g["index"][k])), g["init"][k], int(g["pose_type"][k]), g["turn"][k], g["steer"][k],
               g["dist"][k], g["pose"][k], g["path"][po[k]:po[k + 1]])


def _start_end_cases():
    """Rows written by get_start_end_pose_for_reeds_shepp of the reference (:300-364)."""
    from headland_trajectory_planning_b200 import scenarios as SC
    g = np.load(GOLD)
    po = np.concatenate([[0], np.cumsum(g["path_len"])])
    for k in np.nonzero(g["pose_type"] >= 100)[0]:
        r0, r1 = int(g["turn"][k]) // 10, int(g["turn"][k]) % 10
        yield (SC.scenario_spec(int(g["index"][k])), r0, r1, int(g["pose_type"][k]) - 100, g["init"][k], g["pose"][k],
               g["dist"][k], g["path"][po[k]][0])


def test_oracle_matches_reference_golden():
    dists = set()
    for k, sp, init, typ, turn, steer, dist, pose, path in _cases():
        env = OP.OrchardGeometryEnvironment(sp["rows"], [], tree_width=sp["tree_width"], headland_width=sp["headland_width"])
        car = OP.CarModel(**sp["car"])
        d, p, pa = OP.get_offset_pose(init, typ, turn, car, env, steer_angle=steer)
        assert d == dist and np.array_equal(p, pose) and np.array_equal(pa, path), k
        dists.add(round(float(d), 1))
    assert len(dists) >= 5            # a spread of offsets, not only 0.0


@pytest.mark.gpu
def test_gpu_offset_pose_matches_reference_golden(built_library):
    from headland_trajectory_planning_b200 import safety_forward_path_plan as SF
    from headland_trajectory_planning_b200.car_model import CarModel
    from headland_trajectory_planning_b200.orchard_geometry_environment import OrchardGeometryEnvironment
    for k, sp, init, typ, turn, steer, dist, pose, path in _cases():
        env = OrchardGeometryEnvironment(sp["rows"], [], tree_width=sp["tree_width"], headland_width=sp["headland_width"])
        car = CarModel(**sp["car"])
        with contextlib.redirect_stdout(io.StringIO()):
            d, p, pa = SF.get_offset_pose(init, typ, turn, car, env, steer_angle=steer)
        assert d == dist, (k, d, dist)                                   # chosen offset: exact
        assert np.array_equal(p, pose)
        assert pa.shape == path.shape
        np.testing.assert_allclose(pa[:, :3], path[:, :3], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(pa[:, 3], path[:, 3], rtol=1e-12)
        assert np.array_equal(pa[:, 4], path[:, 4])


@pytest.mark.gpu
def test_gpu_start_end_pose_for_reeds_shepp(built_library):
    """get_start_end_pose_for_reeds_shepp (two GPU offset sweeps + the common outmost x) and
    get_all_reeds_shepp_paths_full against the reference's own results / the pinned RS port."""
    import math
    from oracle import rs_port
    from headland_trajectory_planning_b200 import safety_forward_path_plan as SF
    from headland_trajectory_planning_b200.car_model import CarModel
    from headland_trajectory_planning_b200.orchard_geometry_environment import OrchardGeometryEnvironment
    n = 0
    for sp, r0, r1, side, start, end, leave, enter in _start_end_cases():
        env = OrchardGeometryEnvironment(sp["rows"], [], tree_width=sp["tree_width"], headland_width=sp["headland_width"])
        car = CarModel(**sp["car"])
        with contextlib.redirect_stdout(io.StringIO()):
            s2, e2, l2, n2 = SF.get_start_end_pose_for_reeds_shepp(sp["rows"], r0, r1, car, env, side=side)
        assert np.array_equal(s2, start) and np.array_equal(e2, end) and l2 == leave and n2 == enter
        if n < 4:
            got = SF.get_all_reeds_shepp_paths_full(s2, e2, 1.0 / car.curvature)
            maxc = 1.0 / (1.0 / car.curvature)            # what the reference passes (:375); differs from curvature by an ulp
            ref = rs_port.calc_all_paths(s2[0], s2[1], s2[2], e2[0], e2[1], e2[2], maxc, 0.1)
            assert len(got) == len(ref)
            for a, p in zip(got, ref):
                # These pose pairs are mirror images on purpose (same x, opposite headings), so the goal's x in the
                # start frame is 0 in exact arithmetic and the reference's "pop trailing points while px == 0.0"
                # (reeds_shepp.py:520-524) hinges on the last bit of libm's atan2 / acos / sin / cos: the final
                # sample (= the goal pose) may be kept by one implementation and dropped by the other.
                assert abs(a.shape[0] - len(p.x)) <= 1 and a.shape[1] == 5
                m = min(a.shape[0], len(p.x))
                np.testing.assert_allclose(a[:m, 0], p.x[:m], rtol=1e-5, atol=1e-6)
                np.testing.assert_allclose(a[:m, 1], p.y[:m], rtol=1e-5, atol=1e-6)
                assert np.array_equal(a[:m, 4], np.array(p.directions[:m], dtype=float))
        n += 1
    assert n >= 16


@pytest.mark.gpu
def test_gpu_arc_paths_match_host_rollout(built_library):
    """hl_arc_paths against the mirror's numpy CarModel.calculate_motion_path (<= 1e-12)."""
    from headland_trajectory_planning_b200 import ops
    from headland_trajectory_planning_b200.car_model import CarModel
    rng = np.random.default_rng(11)
    car = CarModel()
    rows, want = [], []
    for _ in range(200):
        pose = np.array([rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-3.14, 3.14)])
        steer = rng.choice([0.55, -0.55, 0.3, 0.0, -0.41])
        direction = rng.choice([1.0, -1.0])
        dyaw = rng.uniform(0.3, 1.6)
        rows.append([pose[0], pose[1], pose[2], steer, direction, dyaw / car.curvature, car.WHEEL_BASE, 0.0])
        want.append(car.calculate_motion_path(pose, [steer, direction], dyaw, 0.1)[:, :3])
    poses, offs = ops.arc_paths(np.array(rows), 0.1)
    poses = poses.cpu().numpy()
    for i, w in enumerate(want):
        got = poses[offs[i]:offs[i + 1]]
        assert got.shape == w.shape
        assert np.abs(got - w).max() < 1e-12
