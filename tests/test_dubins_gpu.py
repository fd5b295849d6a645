"""SURVEY.md 8(f) ranks 2-3 on the GPU: Dubins paths + spline course (K9), distance of a swept path to the field
boundary (K10), the corridor test (K11), the start / end pose sampling sweeps, circle-back turns and Pawn-mode Hybrid A*.

Goldens (``tests/golden/sampling_golden.npz``, ``pawn_golden.npz``; generator ``oracle/gen_golden.py sampling | pawn``)
come from the REFERENCE's own ``safety_forward_path_plan.py`` / ``hybrid_a_star_search.py`` run in the build container
with ``dubins`` := ``oracle.dubins_port`` (pydubins is un-vendored and un-pinned: parity unpinned at that boundary) and
the oracle's geometry."""
import math
import os

import numpy as np
import pytest

import hl_helpers as H
from oracle import dubins_port as DP
from oracle import geometry as geo
from oracle import planner as OP
from oracle.gen_golden import SAMPLING_CASES, _AUX
from oracle.obca_util import calc_spline_course

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _wrap(a):
    return (a + math.pi) % (2 * math.pi) - math.pi


def test_dubins_words_samples_and_course(built_library):
    """2000 random pose pairs: shortest word, length, every sample_many configuration and the spline course."""
    from headland_trajectory_planning_b200 import ops
    rng = np.random.default_rng(0)
    n = 2000
    pairs = np.stack([rng.uniform(-12, 12, n), rng.uniform(-12, 12, n), rng.uniform(-math.pi, math.pi, n),
                      rng.uniform(-12, 12, n), rng.uniform(-12, 12, n), rng.uniform(-math.pi, math.pi, n)], axis=1)
    rho = 3.098978705155902
    for append_goal in (False, True):
        rows, off, word, length, samples, slot_off = ops.dubins_course_batch(pairs, rho, step=0.2, ds=0.2, append_goal=append_goal,
                                                                            want_samples=True)
        rows, samples = rows.cpu().numpy(), samples.cpu().numpy()
        for k in range(n):
            p = DP.shortest_path(pairs[k, :3], pairs[k, 3:], rho)
            assert word[k] == p.path_type(), k
            assert abs(length[k] - p.path_length()) <= 1e-12 * max(1.0, p.path_length())
            qs, _ = p.sample_many(0.2)
            got = samples[slot_off[k]:slot_off[k + 1] - 1]
            assert len(got) == len(qs), k
            if len(qs):
                q = np.array(qs)
                np.testing.assert_allclose(got[:, :2], q[:, :2], rtol=0, atol=1e-9)
                assert np.abs(_wrap(got[:, 2] - q[:, 2])).max() < 1e-9
            if k % 10 == 0 and len(qs) >= 2:                     # the course of every 10th pair against scipy
                pts = np.array(qs)[:, :2]
                if append_goal:
                    pts = np.vstack([pts, pairs[k, 3:5]])
                rx, ry, ryaw, rk, _ = calc_spline_course(pts[:, 0], pts[:, 1], ds=0.2)
                r = rows[off[k]:off[k + 1]]
                assert len(r) == len(rx), k
                np.testing.assert_allclose(r[:, 0], rx, rtol=0, atol=1e-8)
                np.testing.assert_allclose(r[:, 1], ry, rtol=0, atol=1e-8)
                assert np.abs(_wrap(r[:, 2] - np.array(ryaw))).max() < 1e-6
                np.testing.assert_allclose(r[:, 3], rk, rtol=1e-5, atol=1e-6)
    # coincident poses: several words tie at a full circle, the course is still well defined; no crash, some word
    same = pairs[:2].copy()
    same[:, 3:] = same[:, :3]
    _, off2, word2, _ = ops.dubins_course_batch(same, rho, step=0.2, ds=0.2)
    assert (word2 >= 0).all() and len(off2) == 3
    # the drop-in module
    from headland_trajectory_planning_b200 import dubins
    p = dubins.shortest_path(tuple(pairs[5, :3]), tuple(pairs[5, 3:]), rho)
    o = DP.shortest_path(pairs[5, :3], pairs[5, 3:], rho)
    qs, ts = p.sample_many(0.1)
    oq, ot = o.sample_many(0.1)
    assert ts == ot and np.allclose(np.array(qs)[:, :2], np.array(oq)[:, :2], atol=1e-9) and p.path_type() == o.path_type()


def _case_objects(case):
    from headland_trajectory_planning_b200.car_model import CarModel
    from headland_trajectory_planning_b200.orchard_geometry_environment import OrchardGeometryEnvironment
    l_std, slope, rw, hw, aux, atf = case[:6]
    np.random.seed(1)
    rows = OP.create_tree_rows(8, rw, 20, slope_angle=math.radians(slope), l_std=l_std)
    kw = dict(max_steer=0.55, axle_to_front=atf, axle_to_back=0.55, width=1.48, aux_poly_features=_AUX[aux], with_aux=bool(_AUX[aux]))
    return rows, (OP.OrchardGeometryEnvironment(rows, [], tree_width=0.3, headland_width=hw), OP.CarModel(**kw)), \
        (OrchardGeometryEnvironment(rows, [], tree_width=0.3, headland_width=hw), CarModel(**kw))


def test_min_distance_to_boundary(built_library):
    g = np.load(os.path.join(GOLD, "sampling_golden.npz"))
    po = np.concatenate([[0], np.cumsum(g["dubins_path_len"])])
    for k, case in enumerate(SAMPLING_CASES):
        rows, (o_env, o_car), (g_env, g_car) = _case_objects(case)
        path = g["dubins_path"][po[k]:po[k + 1]]
        got = g_env.get_min_distance_to_boundary(g_car, path[:, :3], with_aux=True)
        assert abs(got - g["dubins_min_dist"][k]) < 1e-9, (k, got, g["dubins_min_dist"][k])
        assert g_env.check_path_feasibility(g_car, path[:, :3], boundary_check=False, aux_check=True) == bool(g["dubins_feasible"][k])
        # body only, and a path pushed across the boundary (negative distance)
        want = o_env.get_min_distance_to_boundary(o_car, path[::3, :3], with_aux=False)
        assert abs(g_env.get_min_distance_to_boundary(g_car, path[::3, :3], with_aux=False) - want) < 1e-9
        shifted = path[:, :3] + np.array([-9.0, 0.0, 0.0])
        want = o_env.get_min_distance_to_boundary(o_car, shifted, with_aux=True)
        got = g_env.get_min_distance_to_boundary(g_car, shifted, with_aux=True)
        assert want < 0 and abs(got - want) < 1e-9


def test_dubins_and_circle_back_paths_equal_reference(built_library):
    from headland_trajectory_planning_b200 import safety_forward_path_plan as SF
    g = np.load(os.path.join(GOLD, "sampling_golden.npz"))
    po = np.concatenate([[0], np.cumsum(g["dubins_path_len"])])
    co = np.concatenate([[0], np.cumsum(g["circle_path_len"])])
    for k, case in enumerate(SAMPLING_CASES):
        rows, _, (g_env, g_car) = _case_objects(case)
        pair = g["dubins_pair"][k]
        got = SF.get_dubins_path_full(pair[:3], pair[3:], 1.0 / g_car.curvature)
        want = g["dubins_path"][po[k]:po[k + 1]]
        assert got.shape == want.shape
        np.testing.assert_allclose(got[:, :2], want[:, :2], rtol=0, atol=1e-8)
        assert np.abs(_wrap(got[:, 2] - want[:, 2])).max() < 1e-6
        np.testing.assert_allclose(got[:, 3], want[:, 3], rtol=1e-5, atol=1e-6)
        start, end = SF.get_start_end_pose(rows, case[6], case[6] + 1)
        got = SF.get_circle_back_path_full(np.asarray(start, float), end, 1.0 / g_car.curvature, g_car)
        want = g["circle_path"][co[k]:co[k + 1]]
        assert got.shape == want.shape, k
        np.testing.assert_allclose(got[:, :2], want[:, :2], rtol=0, atol=1e-8)
        assert np.abs(_wrap(got[:, 2] - want[:, 2])).max() < 1e-6
        assert np.array_equal(got[:, 4], want[:, 4])


def test_pose_sampling_sweeps_equal_reference(built_library):
    """sample_start_end_pose_for_dubins / _reeds_shepp / _circle_back: the chosen sample of every golden case."""
    from headland_trajectory_planning_b200 import safety_forward_path_plan as SF
    g = np.load(os.path.join(GOLD, "sampling_golden.npz"))

    def pack(r):
        return np.full(8, np.nan) if r is None else np.concatenate([np.asarray(r[0], float), np.asarray(r[1], float), [r[2], r[3]]])
    for k, case in enumerate(SAMPLING_CASES):
        rows, _, (g_env, g_car) = _case_objects(case)
        s_row, e_row, acc_d, acc_r = case[6:]
        got = pack(SF.sample_start_end_pose_for_dubins(rows, s_row, e_row, g_car, g_env, accuracy=acc_d))
        np.testing.assert_allclose(got, g["dubins"][k], rtol=0, atol=1e-9, equal_nan=True, err_msg=f"dubins sweep {k}")
        got = pack(SF.sample_start_end_pose_for_reeds_shepp(rows, s_row, e_row, g_car, g_env, accuracy=acc_r))
        np.testing.assert_allclose(got, g["rs"][k], rtol=0, atol=1e-9, equal_nan=True, err_msg=f"rs sweep {k}")
        got = pack(SF.sample_start_end_pose_for_circle_back(rows, s_row, s_row + 1, g_car, g_env))
        np.testing.assert_allclose(got, g["circle"][k], rtol=0, atol=1e-9, equal_nan=True, err_msg=f"circle-back sweep {k}")


def test_corridor_and_classic_circle_back(built_library):
    """K11 against a brute-force statement of the buffered-polyline test, then classic_circle_back_turning_path on the
    notebook's fish-tail case (test/classic_planner.ipynb cell 15): a feasible turn that starts and ends on the poses."""
    from headland_trajectory_planning_b200 import ops, safety_forward_path_plan as SF
    rows = H.canonical_rows(l_std=1.0)
    (o_env, o_car, _), (g_env, g_car, _) = H.make_pair(rows, axle_to_front=2.85)
    rng = np.random.default_rng(3)
    lines, want = [], []
    quads = [geo.ccw(q) for q in o_env.obs_poly_list]
    for _ in range(300):
        a = np.array([rng.uniform(-6, 6), rng.uniform(0, 18)])
        d = rng.uniform(-math.pi, math.pi)
        n = int(rng.integers(2, 12))
        pts = [a]
        for _ in range(n - 1):
            d += rng.uniform(-0.5, 0.5)
            pts.append(pts[-1] + 0.4 * np.array([math.cos(d), math.sin(d)]))
        pts = np.array(pts)
        lines.append(pts)
        hit = False
        for k in range(len(pts) - 1):
            seg = pts[k + 1] - pts[k]
            yaw = math.atan2(seg[1], seg[0])
            pose = np.array([[pts[k, 0], pts[k, 1], yaw]])
            ext = (0.0, float(np.hypot(*seg)), -0.3, 0.3)
            hit |= any(geo.rects_hit_convex(pose, ext, q)[0] for q in quads)
        for k in range(1, len(pts) - 1):
            for q in quads:
                d2 = geo.signed_distance_to_ring(pts[k][None, :], q)[0]
                hit |= (d2 >= 0) or (-d2 <= 0.3)
        want.append(hit)
    off = np.concatenate([[0], np.cumsum([len(l) for l in lines])])
    got = ops.corridor_hits(g_env._env_batch(g_car), np.vstack(lines), off, 0.3).cpu().numpy().astype(bool)
    assert np.array_equal(got, np.array(want)) and 0.05 < np.mean(want) < 0.95
    sx = -1.30805046
    path = SF.classic_circle_back_turning_path([sx, 3.75, math.pi], [sx, 8.75, 0.0], g_env, g_car)
    assert path.shape[1] == 5 and len(path) > 50
    assert g_env.check_path_feasibility(g_car, path[:, :3], boundary_check=False)
    assert np.allclose(path[0, :2], [sx, 3.75], atol=1e-6) and np.allclose(path[-1, :2], [sx, 8.75], atol=1e-3)


def test_pawn_mode_search_equals_reference(built_library):
    """motion_type="Pawn": forward-only primitives + Dubins goal extension, 48 config-5 scenarios in one launch against
    the golden the reference's own search loop produced (status, counter, expanded-key sequence, path)."""
    from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep
    from headland_trajectory_planning_b200.env_batch import EnvBatch
    from headland_trajectory_planning_b200.hybrid_a_star_search import unpack_path
    g = np.load(os.path.join(GOLD, "pawn_golden.npz"))
    n = len(g["index"])
    specs = [SC.scenario_spec(int(i)) for i in g["index"]]
    scns = [SC.finalize(sp, f) for sp, f in zip(specs, g["feas"])]
    recs, scen, car = sweep.build_records(scns)
    params = sweep.search_params(car, max_nodes=int(g["max_nodes"]), motion_type="Pawn")
    out = ops.hybrid_astar_batch(EnvBatch(recs), scen, params, path_capacity=4096 * n)
    res = out["results"]
    eo = np.concatenate([[0], np.cumsum(g["n_expanded"])])
    po = np.concatenate([[0], np.cumsum(g["path_len"])])
    bad = []
    for i in range(n):
        ok = (res["status"][i] == g["status"][i] and res["counter"][i] == g["counter"][i]
              and np.array_equal(ops.expanded_of(out, i), g["expanded"][eo[i]:eo[i + 1]])
              and res["path_len"][i] == g["path_len"][i])
        if ok and g["path_len"][i]:
            x, y, yaw, dirs, ks = unpack_path(out, i)
            want = g["path"][po[i]:po[i + 1]]
            ok = (np.allclose(x, want[:, 0], rtol=0, atol=1e-6) and np.allclose(y, want[:, 1], rtol=0, atol=1e-6)
                  and np.abs(_wrap(np.asarray(yaw) - want[:, 2])).max() < 1e-5
                  and np.allclose(ks, want[:, 3], rtol=1e-4, atol=1e-5) and np.array_equal(np.asarray(dirs, float), want[:, 4]))
        if not ok:
            bad.append(i)
    assert not bad, f"{len(bad)} of {n} Pawn scenarios differ from the reference: {bad[:10]}"
    assert (res["arrival"] == 1).sum() >= 10 and (res["status"] == 3).sum() >= 5
    # the drop-in class, API default motion type
    from headland_trajectory_planning_b200.hybrid_a_star_search import HybridAStarSearch
    env, car1, heur = SC.build_host_objects(scns[2])
    s = HybridAStarSearch(scns[2]["start"], scns[2]["goal"], env, car1, heur, plan_resolution=scns[2]["step_size"])
    assert s.motion_type == "Pawn"
    r = s.hybrid_a_star_search(max_nodes=int(g["max_nodes"]))
    assert r[5] == g["counter"][2] and len(r[0]) == g["path_len"][2]


def test_dubins_word_and_length_equal_reference_dubins_path(built_library):
    """K9 against the reference's OWN pure-Python Dubins planner (path_planner/utils/dubins_path.py run unmodified,
    tests/golden/dubins_ref_golden.npz, generator oracle/gen_golden.py dubins_ref): same word, same length.
    The first 100 cases go through the `dubins` drop-in as given; all 2000 in one batch scaled to a unit turning
    radius (a Dubins path scales with rho)."""
    from headland_trajectory_planning_b200 import dubins, ops
    g = np.load(os.path.join(GOLD, "dubins_ref_golden.npz"))
    cases, word, length = g["cases"], g["word"], g["length"]
    for c, w, ln in zip(cases[:100], word[:100], length[:100]):
        p = dubins.shortest_path(tuple(c[:3]), tuple(c[3:6]), float(c[6]))
        assert p.path_type() == int(w)
        assert abs(p.path_length() - ln) <= 1e-9                       # metres; tolerance as in tests/test_dubins_ref.py
    pairs = cases[:, :6].copy()
    pairs[:, [0, 1, 3, 4]] /= cases[:, 6:7]
    _, _, got_word, got_len = ops.dubins_course_batch(pairs, 1.0, step=1.0, ds=1.0)
    assert np.array_equal(np.asarray(got_word), word)
    assert np.allclose(np.asarray(got_len) * cases[:, 6], length, rtol=1e-12, atol=1e-9)
