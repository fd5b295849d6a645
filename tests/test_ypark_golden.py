"""SURVEY.md 8(f) rank 1 -- the Y-type parking sweep (headland_path_planning.py:382-527).

``tests/golden/ypark_golden.npz`` holds the results of the REFERENCE's own ``search_y_type_parking_path``
(imported unmodified through ``oracle.ref_loader.load_planner``, generator ``oracle/gen_golden.py ypark``)
on config-5 environments; the oracle restatement must reproduce the chosen parameters exactly and the
paths to 1e-12 (the reference's 4x4 BLAS product may round differently in the last bit), and so must the
CUDA path (K5 ``hl_ypark_paths`` + K1 + reduce) within the north_star tolerance (1e-5 relative on poses,
booleans / chosen candidate exact)."""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import planner as OP  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ypark_golden.npz")


def _cases(limit=None):
    from headland_trajectory_planning_b200 import scenarios as SC
    g = np.load(GOLD)
    po = np.concatenate([[0], np.cumsum(g["path_len"])])
    n = len(g["index"]) if limit is None else min(limit, len(g["index"]))
    for k in range(n):
        sp = SC.scenario_spec(int(g["index"][k]))
        ps = g["param_sets"][int(g["pset"][k])]
        yield k, sp, g["end"][k], g["bdir"][k], ps, bool(g["found"][k]), g["par"][k], g["path"][po[k]:po[k + 1]]


def test_oracle_matches_reference_golden():
    n_found = 0
    for k, sp, end, bdir, ps, found, par, path in _cases(limit=60):      # 20 scenarios x 3 parameter sets
        env = OP.OrchardGeometryEnvironment(sp["rows"], [], tree_width=sp["tree_width"], headland_width=sp["headland_width"])
        car = OP.CarModel(**sp["car"])
        assert OP.get_backward_steer_dir_for_y_type_parking(sp["start"], end) == bdir
        got_path, got_par = OP.search_y_type_parking_path(car, env, end, bdir, -bdir, *ps[:8], step_size=ps[8], debug=True)
        assert (len(got_par) > 0) == found, k
        if found:
            n_found += 1
            assert np.array_equal(np.array(got_par), par), (k, got_par, par)       # chosen candidate: exact
            assert got_path.shape == path.shape
            np.testing.assert_allclose(got_path, path, rtol=0, atol=1e-12)
    assert n_found >= 15


def test_candidate_enumeration_matches_loop_order():
    """The vectorised candidate table of the product equals the oracle's nested loops (and hence the
    reference's, :405-420), including the arange quirks (0.42 > max_steer_backward 0.35)."""
    from headland_trajectory_planning_b200 import headland_path_planning as HP
    for ps in np.load(GOLD)["param_sets"]:
        a = HP.y_park_candidates(*ps[:8])
        b = OP.y_park_candidates(1.0, -1.0, *ps[:8])
        assert np.array_equal(a, b)
    assert any(abs(v - 0.42) < 1e-12 for v in HP.y_park_candidates(0.35, 0.55, 2.0, 2.5, 1.4, 0.7, 0.22, 0.50)[:, 2])


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present (GPU box)")
def test_live_reference_sweep_and_search_loop():
    """Run the reference's OWN code here: the Y-park sweep of the notebook (cell 9 -> 1.70/2.00/0.00/0.50) and
    its Hybrid A* loop on two config-5 scenarios, against the oracle restatements."""
    import contextlib
    import io
    from headland_trajectory_planning_b200 import scenarios as SC
    from oracle import baseline as OB
    H = ref_loader.load_planner("headland_path_planning")
    np.random.seed(1)
    rows = OP.create_tree_rows(8, 2.5, 20, math.radians(10), l_std=0.0)
    env = OP.OrchardGeometryEnvironment(rows, [], tree_width=0.3, headland_width=6.0)
    car = OP.CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48)
    start = OP.get_base_pose(1, rows, -1.0, side=OP.NEAR_SIDE, pose_type=OP.LEAVE_POSE)
    end = OP.get_base_pose(3, rows, 3.66, side=OP.NEAR_SIDE, pose_type=OP.ENTER_POSE)
    bdir = H.get_backward_steer_dir_for_y_type_parking(start, end)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        path, par = H.search_y_type_parking_path(car, env, end, bdir, -bdir, 0.15, 0.55, 3.0, 2.0, 1.0, 1.0, 0.0, 0.5,
                                                 step_size=0.2, debug=True)
    assert "backward distance:1.70, forward distance:2.00, backward steer:0.00, forward steer:0.50" in buf.getvalue()
    p2, par2 = OP.search_y_type_parking_path(car, env, end, bdir, -bdir, 0.15, 0.55, 3.0, 2.0, 1.0, 1.0, 0.0, 0.5,
                                             step_size=0.2, debug=True)
    assert list(par) == list(par2)
    np.testing.assert_allclose(p2, path, rtol=0, atol=1e-12)
    R = ref_loader.load_planner("hybrid_a_star_search")
    for idx in (0, 6):
        spec = SC.scenario_spec(idx)
        scn = SC.finalize(spec, OB.candidate_feasibility(spec))
        e = OP.OrchardGeometryEnvironment(scn["rows"], [], tree_width=scn["tree_width"], headland_width=scn["headland_width"])
        c = OP.CarModel(**scn["car"])
        h = OP.ReferenceLineHeuristic(scn["waypoints"], scn["goal"], c)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = R.HybridAStarSearch(scn["start"], scn["goal"], e, c, h, motion_type="King",
                                      plan_resolution=scn["step_size"]).hybrid_a_star_search(max_nodes=400)
            ora = OP.HybridAStarSearch(scn["start"], scn["goal"], e, c, h, motion_type="King",
                                       plan_resolution=scn["step_size"]).hybrid_a_star_search(max_nodes=400)
        assert ref[5] == ora[5]
        for a, b in zip(ref[:5], ora[:5]):
            assert np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))


@pytest.mark.gpu
def test_gpu_sweep_matches_reference_golden(built_library):
    """Product path (mirror of the reference function, one K5 + K1 + reduce launch per sweep)."""
    import contextlib
    import io
    from headland_trajectory_planning_b200 import headland_path_planning as HP
    from headland_trajectory_planning_b200.car_model import CarModel
    from headland_trajectory_planning_b200.orchard_geometry_environment import OrchardGeometryEnvironment
    n_found = 0
    for k, sp, end, bdir, ps, found, par, path in _cases():
        env = OrchardGeometryEnvironment(sp["rows"], [], tree_width=sp["tree_width"], headland_width=sp["headland_width"])
        car = CarModel(**sp["car"])
        assert HP.get_backward_steer_dir_for_y_type_parking(sp["start"], end) == bdir
        with contextlib.redirect_stdout(io.StringIO()):
            got_path, got_par = HP.search_y_type_parking_path(car, env, end, bdir, -bdir, *ps[:8], step_size=ps[8], debug=True)
        assert (len(got_par) > 0) == found, k
        if found:
            n_found += 1
            assert np.array_equal(np.array(got_par), par), (k, got_par, par)
            assert got_path.shape == path.shape
            np.testing.assert_allclose(got_path[:, :3], path[:, :3], rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(got_path[:, 3], path[:, 3], rtol=1e-12)
            assert np.array_equal(got_path[:, 4], path[:, 4])
    assert n_found >= 20


@pytest.mark.gpu
def test_gpu_candidate_paths_match_numpy(built_library):
    """K5 against the host rollout of the same candidates: the poses agree to ~1 ulp (bit-exact wherever CUDA's
    cos/sin/tan agree with libm), including zero steer, both steer signs and the duplicated junction pose."""
    from headland_trajectory_planning_b200 import headland_path_planning as HP, ops
    rng = np.random.default_rng(5)
    car = OP.CarModel()
    rows = []
    for _ in range(300):
        rows.append([rng.uniform(0.7, 3.5), rng.uniform(1.0, 2.5), rng.choice([0.0, 0.22, -0.35, 0.4]),
                     rng.choice([0.3, -0.5, 0.55, 0.0]), rng.uniform(-20, 20), rng.uniform(-20, 20),
                     rng.uniform(-math.pi, math.pi), 1.9])
    rows = np.array(rows)
    for step in (0.1, 0.2):
        poses, offs = ops.ypark_paths(rows, step)
        poses = poses.cpu().numpy()
        worst = 0.0
        for i, r in enumerate(rows):
            want = OP.get_path_in_odom(r[4:7], OP.get_y_type_parking_path(car, r[0], r[2], r[1], r[3], step))[:, :3]
            got = poses[offs[i]:offs[i + 1]]
            assert got.shape == want.shape
            worst = max(worst, np.abs(got - want).max())
        assert worst < 1e-12, worst


@pytest.mark.gpu
def test_gpu_batch_sweep(built_library):
    """Many sweeps in one launch (the sweep stage that feeds hl_hybrid_astar_batch) = the per-problem answers."""
    from headland_trajectory_planning_b200 import headland_path_planning as HP, scenarios as SC, sweep
    from headland_trajectory_planning_b200.env_batch import EnvBatch, make_record
    from headland_trajectory_planning_b200.car_model import CarModel
    from headland_trajectory_planning_b200.orchard_geometry_environment import OrchardGeometryEnvironment
    g = np.load(GOLD)
    sel = [k for k in range(len(g["index"])) if g["pset"][k] == 1][:24]
    ps = g["param_sets"][1]
    recs, ends, bdirs, wbs = [], [], [], []
    for k in sel:
        sp = SC.scenario_spec(int(g["index"][k]))
        env = OrchardGeometryEnvironment(sp["rows"], [], tree_width=sp["tree_width"], headland_width=sp["headland_width"])
        car = CarModel(**sp["car"])
        recs.append(make_record(env, car))
        ends.append(g["end"][k]); bdirs.append(g["bdir"][k]); wbs.append(car.WHEEL_BASE)
    cands = HP.y_park_candidates(*ps[:8])
    first, feas, goal = HP.search_y_type_parking_path_batch(EnvBatch(recs), np.arange(len(sel)), np.array(ends), bdirs, wbs,
                                                            cands, step_size=ps[8])
    po = np.concatenate([[0], np.cumsum(g["path_len"])])
    for j, k in enumerate(sel):
        # note: the end pose itself is the last pose of every candidate, so a blocked end pose means "none"
        if g["found"][k]:
            assert first[j] >= 0 and np.array_equal(cands[first[j]], g["par"][k]), (j, k)
            np.testing.assert_allclose(goal[j], g["path"][po[k]][:3], rtol=1e-5, atol=1e-9)
        else:
            assert first[j] == -1 and not feas[j].any()
