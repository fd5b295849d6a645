"""K1 parity: per-pose collision booleans from hl_collision_check (through the C
ABI) vs the CPU oracle -- bit-exact (north_star tier 1)."""
import math

import numpy as np
import pytest

import hl_helpers as H

pytestmark = pytest.mark.gpu


def _cmp(gpu_flags, oracle_flags, what):
    g = gpu_flags.cpu().numpy().astype(bool)
    bad = np.nonzero(g != oracle_flags)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} of {len(g)} booleans differ, first at {bad[:5]}"


@pytest.mark.parametrize("l_std,slope", [(0.0, 10.0), (1.0, 10.0), (0.5, 0.0)])
def test_body_obstacles_and_boundary(built_library, l_std, slope):
    rows = H.canonical_rows(l_std=l_std, slope_deg=slope)
    (o_env, o_car, _), (g_env, g_car, _) = H.make_pair(rows, obstacles=[(-3.0, 6.0), (24.0, 11.0)])
    rng = np.random.default_rng(7)
    poses = np.concatenate([H.random_poses(rng, 60000), H.headland_poses(rng, 60000, rows)])
    for bc in (False, True):
        want = o_env.pose_flags(o_car, poses, boundary_check=bc)
        got = g_env.pose_flags(g_car, poses, boundary_check=bc)
        _cmp(got, want, f"boundary_check={bc}")
        assert 0.02 < want.mean() < 0.98


@pytest.mark.parametrize("aux", [H.MOWER_AUX, H.PRUNER_AUX, H.SPRAYER_AUX])
def test_aux_rectangles_stride2(built_library, aux):
    rows = H.canonical_rows()
    (o_env, o_car, _), (g_env, g_car, _) = H.make_pair(rows, axle_to_front=2.85, aux=aux)
    rng = np.random.default_rng(11)
    poses = H.headland_poses(rng, 40001, rows)
    want = o_env.pose_flags(o_car, poses, boundary_check=True, aux_check=True)
    got = g_env.pose_flags(g_car, poses, boundary_check=True, aux_check=True)
    _cmp(got, want, "aux")
    # aux only counts at even pose indices (car_model.py:58)
    body_only = o_env.pose_flags(o_car, poses, boundary_check=True, aux_check=False)
    assert (want[1::2] == body_only[1::2]).all() and (want[0::2] != body_only[0::2]).any()


def test_lane_containment(built_library):
    rows = H.canonical_rows()
    start = np.array([0.0, 3.75, math.pi])
    goal = np.array([-1.5, 11.0, 0.3])
    way = np.array([[0.0, 3.75], [-4.5, 5.0], [-4.06, 7.5], [-3.62, 10.0], [-1.5, 11.0]])
    (_, o_car, o_h), (_, g_car, g_h) = H.make_pair(rows, waypoints=way, goal=goal)
    rng = np.random.default_rng(3)
    n = 80000
    ang = rng.uniform(0, 2 * math.pi, n)
    rad = rng.uniform(0.0, 9.0, n)
    c = way[rng.integers(0, len(way), n)]
    poses = np.stack([c[:, 0] + rad * np.cos(ang), c[:, 1] + rad * np.sin(ang),
                      rng.uniform(-math.pi, math.pi, n)], axis=1)
    want = o_h.pose_flags(o_car, poses)
    from headland_trajectory_planning_b200 import ops
    got, n_exact = ops.collision_check(g_h._env_batch(g_car), poses, flags=ops.CHECK_LANE, count_exact=True)
    _cmp(got, want, "lane")
    assert 0.05 < want.mean() < 0.95
    # the float64 path is an escalation, not the norm
    assert int(n_exact.item()) < 0.25 * n


def test_touching_counts_as_collision(built_library):
    """GEOS ``intersects`` is closed: a footprint that exactly touches a tree row collides."""
    rows = np.array([[[0.0, 2.0 * i], [20.0, 2.0 * i]] for i in range(8)])
    (o_env, o_car, _), (g_env, g_car, _) = H.make_pair(rows, tree_width=0.5, headland_width=8.0)
    # rows occupy y in [2i-0.25, 2i+0.25]; body half width 0.74 -> touching at y = 2i + 0.25 + 0.74
    ys = np.array([0.25 + 0.74, 0.25 + 0.74 + 1e-9, 0.25 + 0.74 - 1e-9, 0.9])
    poses = np.stack([np.full(4, 5.0), ys, np.zeros(4)], axis=1)
    want = o_env.pose_flags(o_car, poses, boundary_check=False)
    got = g_env.pose_flags(g_car, poses, boundary_check=False)
    _cmp(got, want, "touching")
    assert want.tolist() == [True, want[1], True, True]


def test_empty_and_single(built_library):
    rows = H.canonical_rows()
    (o_env, o_car, _), (g_env, g_car, _) = H.make_pair(rows)
    assert g_env.check_path_feasibility(g_car, np.zeros((0, 3))) is True
    p = np.array([[-3.0, 3.75, math.pi]])
    assert g_env.check_path_feasibility(g_car, p) == o_env.check_path_feasibility(o_car, p)
    far = np.array([[1e4, -2e4, 0.3], [float("nan"), 0.0, 0.0]])
    got = g_env.pose_flags(g_car, far).cpu().numpy().astype(bool)
    assert got[0]


def test_non_finite_poses_are_infeasible(built_library):
    """A pose with NaN / infinite coordinates can never be certified free: K1 reports it infeasible (this is also
    how hl_ypark_paths poisons a candidate whose pose count disagrees with the host's)."""
    rows = H.canonical_rows()
    (o_env, o_car, _), (g_env, g_car, _) = H.make_pair(rows)
    poses = np.array([[np.nan, 1.0, 0.0], [1.0, np.nan, 0.0], [-3.0, 8.0, np.nan], [np.inf, 0.0, 0.0],
                      [-3.0, 8.0, np.inf], [-4.0, 9.0, 0.3]])
    got = g_env.pose_flags(g_car, poses).cpu().numpy().astype(bool)
    assert got[:5].all()
    assert got[5] == o_env.pose_flags(o_car, poses[5:6])[0]


def test_candidate_generators_empty(built_library):
    from headland_trajectory_planning_b200 import ops
    p, o = ops.ypark_paths(np.zeros((0, 8)), 0.1)
    assert p.shape == (0, 3) and list(o) == [0]
    p, o = ops.arc_paths(np.zeros((0, 8)), 0.1)
    assert p.shape == (0, 3) and list(o) == [0]


@pytest.mark.parametrize("aux", [None, H.SPRAYER_AUX])
def test_staged_pipeline_equals_per_pose_filter(built_library, aux):
    """K1 has two paths: the staged / compacted pipeline on a TMA-fed slice of one shared-memory environment, and the
    monolithic per-pose filter on global memory for slices that mix environments.  Two copies of the same
    environment with alternating env_id force the second path; both must give the same booleans on 1.5 M poses
    (every flag combination, ragged tail, a pose pointer that is not 16-byte aligned = no bulk copy)."""
    import torch
    from headland_trajectory_planning_b200 import ops
    from headland_trajectory_planning_b200.env_batch import EnvBatch, make_record
    rows = H.canonical_rows(l_std=0.5)
    start = np.array([0.0, 3.75, math.pi])
    goal = np.array([-1.5, 11.0, 0.3])
    way = np.array([[0.0, 3.75], [-4.5, 5.0], [-4.06, 7.5], [-3.62, 10.0], [-1.5, 11.0]])
    (_, _, _), (g_env, g_car, g_h) = H.make_pair(rows, axle_to_front=2.85, aux=aux, waypoints=way, goal=goal)
    rec = make_record(g_env, g_car, g_h)
    envs = EnvBatch([rec, rec])
    rng = np.random.default_rng(5)
    n = 1_500_000 + 77
    poses = H.headland_poses(rng, n, rows)
    poses[::1000, 0] = np.nan                      # undecidable poses go to the float64 path in both
    poses[7::1000, 0] = 5e4                        # beyond the environment's reach
    d_poses = torch.from_numpy(poses).cuda()
    idx = torch.arange(n, dtype=torch.int32, device="cuda")
    alt = (idx & 1).to(torch.int32)
    for flags in (1, 3, 8, 11, 7 if aux else 2, 15 if aux else 9):
        fast = ops.collision_check(envs, d_poses, pose_idx=idx, flags=flags)
        slow = ops.collision_check(envs, d_poses, env_id=alt, pose_idx=idx, flags=flags)
        assert torch.equal(fast, slow), f"flags={flags}: {(fast != slow).sum().item()} booleans differ"
        assert 0.01 < fast.float().mean().item() < 0.995
        # unaligned source: the same poses shifted by one row (24 bytes)
        un = ops.collision_check(envs, d_poses[1:], pose_idx=idx[1:].contiguous(), flags=flags)
        assert torch.equal(un, fast[1:]), f"flags={flags}: unaligned source differs"


def test_float64_queue_overflow_and_culls(built_library):
    """K1 queues the poses that need the float64 predicates (48 per CTA) and resolves them after the CTA's last tile;
    when the queue is full they are resolved in place.  A block of 6000 consecutive undecidable poses (NaN position /
    absurd heading) overflows the queue of every CTA that meets it; the booleans must still equal the per-pose filter on
    global memory (second K1 path) and the oracle on a sample.  The bounding-box culls are exercised by poses far
    outside the field polygon's box and far from every obstacle."""
    import torch
    from headland_trajectory_planning_b200 import ops
    from headland_trajectory_planning_b200.env_batch import EnvBatch, make_record
    rows = H.canonical_rows(l_std=0.3)
    way = np.array([[0.0, 3.75], [-4.5, 5.0], [-4.06, 7.5], [-3.62, 10.0], [-1.5, 11.0]])
    (o_env, o_car, o_h), (g_env, g_car, g_h) = H.make_pair(rows, aux=H.SPRAYER_AUX, waypoints=way, goal=np.array([-1.5, 11.0, 0.3]))
    rec = make_record(g_env, g_car, g_h)
    envs = EnvBatch([rec, rec])
    rng = np.random.default_rng(11)
    n = 300_000
    poses = H.headland_poses(rng, n, rows)
    poses[50_000:53_000, 0] = np.nan
    poses[53_000:56_000, 2] = 3e7                   # |yaw| >= 1e6: the float32 heading reduction is not trusted
    poses[100_000:110_000, 0] -= 40.0               # far outside the box of the field polygon
    poses[120_000:130_000, 1] += 60.0
    d_poses = torch.from_numpy(poses).cuda()
    idx = torch.arange(n, dtype=torch.int32, device="cuda")
    alt = (idx & 1).to(torch.int32)
    for flags in (3, 7, 15):
        fast = ops.collision_check(envs, d_poses, pose_idx=idx, flags=flags)
        slow = ops.collision_check(envs, d_poses, env_id=alt, pose_idx=idx, flags=flags)
        assert torch.equal(fast, slow), f"flags={flags}: {(fast != slow).sum().item()} booleans differ"
        got = fast.cpu().numpy().astype(bool)
        assert got[50_000:53_000].all()             # NaN position: infeasible
        assert got[100_000:110_000].all() and got[120_000:130_000].all()
    sel = np.concatenate([np.arange(0, 3000), np.arange(53_000, 53_200), np.arange(100_000, 100_200)])
    want = o_env.pose_flags(o_car, poses[sel])
    got = ops.collision_check(envs, d_poses[torch.from_numpy(sel).cuda()], flags=ops.CHECK_OBSTACLES | ops.CHECK_BOUNDARY).cpu().numpy().astype(bool)
    assert (want == got).all()
