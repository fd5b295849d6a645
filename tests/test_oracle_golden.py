"""CPU: the oracle against every golden vector that exists for this path --
fixtures generated from the reference's OWN modules (tests/golden/rs_golden.npz,
df_golden.npz; generator: oracle/gen_golden.py), the known-answer vectors of SURVEY.md
8(c), and the values printed in the reference's notebooks (test/obca.ipynb,
test/classic_planner.ipynb)."""
import math
import os

import numpy as np
import pytest

from oracle import distance_field as DF
from oracle import planner as OP
from oracle import ref_loader, rs_port
from oracle.heapdict_port import HeapDict

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAXC = math.tan(0.55) / 1.9
LET = "SLR"


def test_rs_port_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "rs_golden.npz"))
    off = np.concatenate([[0], np.cumsum(g["states_len"])])
    si = 0
    for i, (q, st) in enumerate(zip(g["sg"], g["steps"])):
        paths = rs_port.calc_all_paths(*q, MAXC, st)
        assert len(paths) == g["count"][i]
        for k, p in enumerate(paths):
            assert "".join(p.ctypes) == "".join(LET[c] for c in g["letters"][i, k] if c >= 0)
            assert np.array_equal(np.asarray(p.lengths), g["lens"][i, k, :len(p.lengths)])      # bit-for-bit
            assert p.L == g["L"][i, k]
            assert len(p.x) == g["npts"][i, k]
            if i < 43:
                s = g["states"][off[si]:off[si + 1]]
                si += 1
                assert np.array_equal(np.asarray(p.x), s[:, 0]) and np.array_equal(np.asarray(p.y), s[:, 1])
                assert np.array_equal(np.asarray(p.yaw), s[:, 2])
                assert np.array_equal(np.asarray(p.cs, dtype=np.float64), s[:, 3])
                assert np.array_equal(np.asarray(p.directions, dtype=np.float64), s[:, 4])


def test_rs_known_answers():
    """SURVEY.md 8(c) KATs (probed from the real module)."""
    p = rs_port.calc_all_paths(0, 0, 0, 3, 4, 1.0, MAXC, 0.1)
    assert ["".join(x.ctypes) for x in p] == ["SLS", "RLR", "RLR", "LRL", "RLRL"]
    assert [len(x.x) for x in p] == [76, 167, 61, 154, 112]
    np.testing.assert_allclose([x.L for x in p], [7.42092959201, 16.3724787623, 5.88553737345, 15.0839514994,
                                                   10.8944608623], rtol=1e-11)
    for x in p:
        assert abs(x.x[-1] - 3) < 1e-9 and abs(x.y[-1] - 4) < 1e-9 and abs(x.yaw[-1] - 1.0) < 1e-9
    p = rs_port.calc_all_paths(-1.30805046, 3.75, math.pi, -1.30805046, 8.75, 0, MAXC, 0.1)
    assert ["".join(x.ctypes) for x in p] == ["LRL", "RLR", "RLR"]
    assert [len(x.x) for x in p] == [101, 101, 101]
    np.testing.assert_allclose([x.L for x in p], [9.73572873375] * 3, rtol=1e-11)
    p = rs_port.calc_all_paths(1, 2, -2, -4, 1.5, 2.5, MAXC, 0.2)
    assert ["".join(x.ctypes) for x in p] == ["SRS", "LRL", "RLR", "RLRL", "LRLR", "LRSR"]
    assert [len(x.x) for x in p] == [34, 73, 59, 44, 48, 34]


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present (GPU box)")
def test_rs_port_matches_live_reference():
    ref = ref_loader.load("reeds_shepp")
    rng = np.random.default_rng(99)
    for i in range(300):
        q = list(rng.uniform(-10, 10, 2)) + [rng.uniform(-math.pi, math.pi)] + list(rng.uniform(-10, 10, 2)) + \
            [rng.uniform(-math.pi, math.pi)]
        a = ref.calc_all_paths(*q, MAXC, 0.1)
        b = rs_port.calc_all_paths(*q, MAXC, 0.1)
        assert len(a) == len(b)
        for pa, pb in zip(a, b):
            assert pa.ctypes == pb.ctypes and pa.lengths == pb.lengths and pa.x == pb.x and pa.y == pb.y \
                and pa.yaw == pb.yaw and pa.directions == pb.directions and pa.cs == pb.cs


def test_distance_field_port_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "df_golden.npz"))
    for occ, goal, mt, out in zip(g["grids"], g["goals"], g["motions"], g["outs"]):
        n = int(goal[0])
        got = DF.holonomic_costs_with_obstacles((int(goal[1]), int(goal[2])), occ[:n, :n], "King" if mt == 0 else "Pawn")
        assert np.array_equal(got, out[:n, :n])
    got = DF.holonomic_costs_with_obstacles(tuple(int(v) for v in g["big_goal"]), g["big_occ"], "King")
    assert np.array_equal(got, g["big_out"])


def test_distance_field_wrap_quirk_kat():
    """free 8x8 grid, goal (1,1): the reference returns 11.314 AT the goal
    cell (index wrap-around) and inf in column 0 for Pawn."""
    a = DF.holonomic_costs_with_obstacles((1, 1), np.zeros((8, 8), dtype=bool), "King")
    assert abs(a[1, 1] - 11.313708498984763) < 1e-12
    b = DF.holonomic_costs_with_obstacles((1, 1), np.zeros((8, 8), dtype=bool), "Pawn")
    assert np.isinf(b[:, 0]).all()


def test_heapdict_tie_order():
    """Equal priorities: the newer entry moves above the older one."""
    h = HeapDict()
    for k, v in (("a", 5.0), ("b", 5.0), ("c", 5004.0)):
        h[k] = v
    assert [h.popitem()[0] for _ in range(3)] == ["b", "a", "c"]
    h = HeapDict()
    for k in range(6):
        h[k] = 1.0
    h[2] = 1.0                                   # re-key: delete (bubble to root + pop) then append
    assert len(h) == 6
    order = [h.popitem()[0] for _ in range(6)]
    assert sorted(order) == list(range(6)) and order[0] == 2


def _notebook_orchard(l_std):
    np.random.seed(1)
    rows = OP.create_tree_rows(8, 2.5, 20, slope_angle=math.radians(10), l_std=l_std)
    env = OP.OrchardGeometryEnvironment(rows, [], tree_width=0.3, headland_width=6.0)
    return rows, env


def test_notebook_classic_planner_goldens():
    """test/classic_planner.ipynb cells 3-6, 10-11."""
    rows, env = _notebook_orchard(1.0)
    start = OP.get_base_pose(1, rows, -0.0, side=OP.NEAR_SIDE, pose_type=OP.LEAVE_POSE)
    np.testing.assert_allclose(start, [0.38166505, 3.75, 3.14159265], atol=5e-9)          # cell 11 stdout
    car = OP.CarModel(max_steer=0.55, axle_to_back=0.55, width=1.48)
    # cell 10 output: body polygon at the end pose (1.008, 8.75, 0) printed to 3 decimals
    end = OP.get_base_pose(3, rows, 0, side=OP.NEAR_SIDE, pose_type=OP.ENTER_POSE)
    from oracle import geometry as geo
    c = geo.rect_corners(end[None, :], car.body_ext)[0]
    want = np.array([[-1.858, 9.49], [-1.858, 8.01], [1.542, 8.01], [1.542, 9.49]])
    safe_end_x = -1.30805046
    c_safe = geo.rect_corners(np.array([[safe_end_x, 8.75, 0.0]]), car.body_ext)[0]
    np.testing.assert_allclose(c_safe, want, atol=6e-4)
    mower = geo.rect_corners(np.array([[safe_end_x, 8.75, 0.0]]), geo.aux_extent([[-1.84, 0.5], 1.0, 1.1]))[0]
    np.testing.assert_allclose(sorted(map(tuple, np.round(mower, 3))),
                               sorted([(-3.148, 9.25), (-2.048, 9.25), (-2.048, 8.25), (-3.148, 8.25)]), atol=6e-4)
    # cell 11: of the three fish-tail words only an R-L-R is feasible (boundary_check=False)
    feas = []
    for p in rs_port.calc_all_paths(safe_end_x, 3.75, math.pi, safe_end_x, 8.75, 0.0, car.curvature, 0.1):
        traj = np.array([p.x, p.y, p.yaw]).T
        if env.check_path_feasibility(car, traj, boundary_check=False):
            feas.append("".join(p.ctypes))
    assert feas == ["RLR"]


def test_notebook_obca_goldens():
    """test/obca.ipynb cells 3, 5, 7, 9, 17: curvature radius, start / end poses, the Y-park
    sweep's first feasible candidate (1.70, 2.00, 0.00, 0.50) and ``counter of nodes: 1``."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(__file__)))
    from headland_trajectory_planning_b200 import scenarios as SC
    rows, env = _notebook_orchard(0.0)
    car = OP.CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48)
    assert 1 / car.curvature == 3.098978705155902                                            # cell 5 stdout
    start = OP.get_base_pose(1, rows, -1.0, side=OP.NEAR_SIDE, pose_type=OP.LEAVE_POSE)
    end = OP.get_base_pose(3, rows, 3.66, side=OP.NEAR_SIDE, pose_type=OP.ENTER_POSE)
    np.testing.assert_allclose(start[:2], [1.66122618, 3.75], atol=5e-9)                      # cell 17 init state
    np.testing.assert_allclose(end, [-2.11713892, 8.75, 0.0], atol=5e-9)                      # cell 17 end state
    # Y-park sweep of search_y_type_parking_path (headland_path_planning.py:382-451) with cell 9's arguments
    bdir = np.sign(math.cos(start[2])) if end[1] - start[1] > 0 else np.sign(-math.cos(start[2]))
    steer_b = list(np.arange(0.0, 0.15 + 0.1, 0.1))
    if np.max(steer_b) < 0.15:
        steer_b.append(0.15)
    steer_f = list(np.arange(0.5, 0.55 + 0.1, 0.1))
    if np.max(steer_f) < 0.55:
        steer_f.append(0.55)
    found = None
    for bl in np.arange(3.0, 1.0, -0.1):
        for fl in np.arange(2.0, 1.0, -0.1):
            for sb in steer_b:
                for sf in steer_f:
                    path = SC.y_park_path(end, bl, sb * bdir, fl, -sf * bdir, car.WHEEL_BASE, 0.2)
                    if env.check_path_feasibility(car, path):
                        found = (bl, fl, sb, sf, path)
                        break
                if found:
                    break
            if found:
                break
        if found:
            break
    assert found is not None
    assert ["%.2f" % v for v in found[:4]] == ["1.70", "2.00", "0.00", "0.50"]               # cell 9 stdout
    goal = found[4][0]
    np.random.seed(0)
    way = env.get_topology_waypoints(start, goal, drive_row_offset=4.5)
    heur = OP.ReferenceLineHeuristic(way, goal, car)
    s = OP.HybridAStarSearch(start, goal, env, car, heur, motion_type="King", plan_resolution=0.2)
    x, y, yaw, dirs, ks, counter = s.hybrid_a_star_search(max_nodes=400)
    assert counter == 1 and len(x) > 10                                                      # "counter of nodes: 1"


def test_king_primitive_table():
    """14 steers 0.55 .. -0.58446, directions alternate."""
    car = OP.CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48)
    rows, env = _notebook_orchard(0.0)
    s = OP.HybridAStarSearch([0, 0, 0], [1, 1, 0], env, car, None, motion_type="King")
    want = [0.55, 0.46273, 0.37547, 0.2882, 0.20093, 0.11367, 0.0264, -0.06087, -0.14813, -0.2354, -0.32266,
            -0.40993, -0.4972, -0.58446]
    np.testing.assert_allclose(s.motion_steers[:, 0], want, atol=6e-6)
    assert list(s.motion_steers[:, 1]) == [1, -1] * 7
    assert round(1.5 / 0.2) == 8 and round(1.5 / 0.1) == 15


def test_oracle_search_matches_reference_run_golden():
    """tests/golden/astar_ref_golden.npz = the REFERENCE's own ``hybrid_a_star_search`` (imported unmodified,
    ``oracle/gen_golden.py astar_ref``) on config-5 scenarios 0..191 with the oracle's geometry objects.  The
    oracle-generated fixture the GPU tests use (astar_golden.npz) must agree with it scenario by scenario:
    counter and every path value bit for bit -- this pins the oracle's search loop, costs, rollout and
    analytic shot on the reference itself (GEOS predicates and the heapdict port stay restated)."""
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "astar_ref_golden.npz"))
    ora = np.load(os.path.join(os.path.dirname(__file__), "golden", "astar_golden.npz"))
    n = len(ref["index"])
    assert n >= 192 and np.array_equal(ref["index"], ora["index"][:n])
    assert np.array_equal(ref["counter"], ora["counter"][:n])
    assert np.array_equal(ref["path_len"], ora["path_len"][:n])
    total = int(ref["path_len"].sum())
    a, b = ref["path"][:total], ora["path"][:total]
    assert np.array_equal(a[:, [0, 1, 3, 4]], b[:, [0, 1, 3, 4]])
    assert np.array_equal(a[:, 2], b[:, 2])
    assert (ref["counter"] > 200).sum() >= 12         # long searches are covered, not only counter == 1
