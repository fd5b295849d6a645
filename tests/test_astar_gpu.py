"""K4 parity: expanded-node key sequence, counter, status and the selected warm-start
path of hl_hybrid_astar_batch vs the CPU oracle (north_star tier 1: bit-exact indices
and node sequence; tier 2: states within 1e-5)."""
import math

import numpy as np
import pytest

import hl_helpers as H
from oracle import planner as OP

pytestmark = pytest.mark.gpu


def _run_pair(rows, start, goal, way, res, max_nodes, headland_width=6.0, axle_to_front=3.0, obstacles=()):
    from headland_trajectory_planning_b200.hybrid_a_star_search import HybridAStarSearch
    (o_env, o_car, o_h), (g_env, g_car, g_h) = H.make_pair(rows, obstacles=obstacles, headland_width=headland_width,
                                                            axle_to_front=axle_to_front, waypoints=way, goal=goal)
    o = OP.HybridAStarSearch(start, goal, o_env, o_car, o_h, motion_type="King", plan_resolution=res)
    want = o.hybrid_a_star_search(max_nodes=max_nodes)
    g = HybridAStarSearch(start, goal, g_env, g_car, g_h, motion_type="King", plan_resolution=res)
    got = g.hybrid_a_star_search(max_nodes=max_nodes)
    return o, want, g, got


def _check(o, want, g, got):
    assert g.status == o.status, (g.status, o.status)
    assert got[5] == want[5], ("counter", got[5], want[5])
    assert g.expanded == [tuple(int(v) for v in k) for k in o.expanded]
    assert len(got[0]) == len(want[0])
    if len(want[0]):
        np.testing.assert_allclose(got[0], want[0], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got[1], want[1], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got[2], want[2], rtol=1e-5, atol=1e-6)
        assert list(got[3]) == [int(d) for d in want[3]]
        np.testing.assert_allclose(got[4], np.asarray(want[4], dtype=np.float64), rtol=1e-12)


def test_notebook_case_counter_1(built_library):
    """test/obca.ipynb cell 9: the first Reeds-Shepp shot from the start node succeeds
    (``counter of nodes: 1``)."""
    rows = H.canonical_rows(l_std=0.0)
    start = OP.get_base_pose(1, rows, -1.0, side=OP.NEAR_SIDE, pose_type=OP.LEAVE_POSE)
    goal = np.array([-2.11713892, 8.75, 0.0]) + np.array([0.0, 0.0, 0.0])
    way = np.vstack((start[:2], [[rows[2, 0, 0] - 4.5, 5.0], [rows[3, 0, 0] - 4.5, 7.5]], goal[:2]))
    o, want, g, got = _run_pair(rows, start, goal, way, 0.2, 400)
    _check(o, want, g, got)


@pytest.mark.parametrize("res,max_nodes,hw,atf", [(0.2, 60, 6.0, 3.0), (0.1, 40, 5.0, 3.6), (0.2, 120, 4.6, 4.2)])
def test_multi_expansion_parity(built_library, res, max_nodes, hw, atf):
    rows = H.canonical_rows(l_std=0.0)
    start = OP.get_base_pose(1, rows, 0.0, side=OP.NEAR_SIDE, pose_type=OP.LEAVE_POSE)
    goal = OP.get_base_pose(4, rows, 2.0, side=OP.NEAR_SIDE, pose_type=OP.ENTER_POSE)
    ends = rows[:, 0, :]
    way = np.vstack((start[:2], [[ends[i, 0] - 4.5, ends[i, 1]] for i in (2, 3, 4)], goal[:2]))
    o, want, g, got = _run_pair(rows, start, goal, way, res, max_nodes, headland_width=hw, axle_to_front=atf)
    _check(o, want, g, got)
    assert want[5] > 1


def test_start_blocked(built_library):
    rows = H.canonical_rows(l_std=0.0)
    start = np.array([5.0, 2.5, 0.0])           # on a tree row
    goal = np.array([-3.0, 8.75, 0.0])
    way = np.vstack((start[:2], goal[:2]))
    o, want, g, got = _run_pair(rows, start, goal, way, 0.2, 50)
    assert want == ([], [], [], [], [], 0)
    assert got == ([], [], [], [], [], 0)
