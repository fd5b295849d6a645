"""K2/K3 parity: Reeds-Shepp words (set + order + sample counts exact, lengths and
states <= 1e-5 relative; north_star tier 2) and per-word collision flags (exact)
from hl_rs_all_paths / hl_rs_sample vs the pinned oracle port."""
import math
import os

import numpy as np
import pytest

import hl_helpers as H
from oracle import rs_port
from oracle import planner as OP

pytestmark = pytest.mark.gpu
MAXC = math.tan(0.55) / 1.9
RTOL = 1e-5          # north_star: Reeds-Shepp lengths / states within 1e-5 relative
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _pairs(n, seed=0):
    rng = np.random.default_rng(seed)
    sg = np.empty((n, 6))
    sg[:, [0, 1, 3, 4]] = rng.uniform(-10, 10, (n, 4))
    sg[:, [2, 5]] = rng.uniform(-math.pi, math.pi, (n, 2))
    return sg


@pytest.mark.parametrize("step", [0.1, 0.2])
def test_words_match_oracle(built_library, step):
    from headland_trajectory_planning_b200 import ops
    sg = _pairs(3000, seed=int(step * 10))
    words, count, order = ops.rs_all_paths(sg, MAXC, step)
    w = ops.rs_words_to_host(words)
    count = count.cpu().numpy()
    for i in range(len(sg)):
        ref = rs_port.calc_all_paths(*sg[i], MAXC, step)
        assert count[i] == len(ref), (i, count[i], len(ref))
        for k, p in enumerate(ref):
            assert w["cand"][i, k] == p.cand
            assert w["npts"][i, k] == len(p.x), (i, k)
            np.testing.assert_allclose(w["len"][i, k, :len(p.lengths)], p.lengths, rtol=RTOL, atol=1e-9)
            np.testing.assert_allclose(w["L"][i, k], p.L, rtol=RTOL)


def test_sampled_states_match_oracle(built_library):
    from headland_trajectory_planning_b200.utils import reeds_shepp as rs_gpu
    cases = [(0, 0, 0, 3, 4, 1.0, 0.1), (-1.30805046, 3.75, math.pi, -1.30805046, 8.75, 0, 0.1),
             (1, 2, -2, -4, 1.5, 2.5, 0.2)]
    cases += [tuple(q) + (0.1,) for q in _pairs(40, seed=5)]
    for c in cases:
        ref = rs_port.calc_all_paths(*c[:6], MAXC, c[6])
        got = rs_gpu.calc_all_paths(*c[:6], MAXC, c[6])
        assert [p.ctypes for p in got] == [p.ctypes for p in ref]
        for a, b in zip(got, ref):
            assert len(a.x) == len(b.x)
            assert a.directions == b.directions
            np.testing.assert_allclose(a.x, b.x, rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(a.y, b.y, rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(a.yaw, b.yaw, rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(a.cs, b.cs, rtol=1e-12)
            np.testing.assert_allclose(a.lengths, b.lengths, rtol=RTOL, atol=1e-9)


def test_pop_order_and_collision_flags(built_library):
    """Heapdict pop order of the candidates (ties are the norm) and the per-word
    collision booleans of the shot loop (hybrid_a_star_search.py:265-276)."""
    from headland_trajectory_planning_b200 import ops
    from headland_trajectory_planning_b200.env_batch import EnvBatch, make_record
    rows = H.canonical_rows()
    (o_env, o_car, _), (g_env, g_car, _) = H.make_pair(rows)
    rng = np.random.default_rng(2)
    n = 1500
    sg = np.empty((n, 6))
    sg[:, [0, 3]] = rng.uniform(-7.0, 1.0, (n, 2))
    sg[:, [1, 4]] = rng.uniform(-1.0, 19.0, (n, 2))
    sg[:, [2, 5]] = rng.uniform(-math.pi, math.pi, (n, 2))
    envs = EnvBatch([make_record(g_env, g_car)])
    words, count, order = ops.rs_all_paths(sg, MAXC, 0.2, envs=envs, flags=ops.CHECK_OBSTACLES | ops.CHECK_BOUNDARY)
    w = ops.rs_words_to_host(words)
    order = order.cpu().numpy()
    count = count.cpu().numpy()
    search = OP.HybridAStarSearch([0, 0, 0], [1, 1, 0], o_env, o_car, None, motion_type="King", plan_resolution=0.2)
    n_free = 0
    for i in range(n):
        node = OP.Node((0, 0, 0), [list(sg[i, :3])], [0], 0, [1], (0, 0, 0))
        search.goal_node = OP.Node((0, 0, 0), [list(sg[i, 3:])], [0], 0, [1], (0, 0, 0))
        ref = search.rs_candidates_in_pop_order(node)
        assert count[i] == len(ref)
        got_order = [int(w["cand"][i, k]) for k in order[i, :count[i]]]
        assert got_order == [p.cand for p, _ in ref], i
        for (p, cost), k in zip(ref, order[i, :count[i]]):
            assert w["cost"][i, k] == cost
            traj = np.array([p.x, p.y, p.yaw]).T
            want = not o_env.check_path_feasibility(o_car, traj)
            assert bool(w["collide"][i, k]) == want, (i, k)
            n_free += (not want)
    assert n_free > 20


def test_words_match_reference_golden(built_library):
    """tests/golden/rs_golden.npz = the REFERENCE's own ``reeds_shepp.calc_all_paths`` (403 pairs incl. the three
    known-answer cases of SURVEY 8c): word set + order + sample counts exact, lengths <= 1e-5 relative."""
    from headland_trajectory_planning_b200 import ops
    g = np.load(os.path.join(GOLD, "rs_golden.npz"))
    for step in np.unique(g["steps"]):
        idx = np.nonzero(g["steps"] == step)[0]
        words, count, _ = ops.rs_all_paths(g["sg"][idx], MAXC, float(step), want_order=False)
        w = ops.rs_words_to_host(words)
        count = count.cpu().numpy()
        assert np.array_equal(count, g["count"][idx])
        for a, i in enumerate(idx):
            for k in range(count[a]):
                nseg = int(w["n_seg"][a, k])
                assert np.array_equal(np.nonzero(g["letters"][i, k] >= 0)[0], np.arange(nseg)) or nseg == (g["letters"][i, k] != -1).sum()
                assert w["npts"][a, k] == g["npts"][i, k], (i, k)
                np.testing.assert_allclose(w["len"][a, k, :nseg], g["lens"][i, k, :nseg], rtol=RTOL, atol=1e-9)
                np.testing.assert_allclose(w["L"][a, k], g["L"][i, k], rtol=RTOL)


def test_rs_sweep_config3_full_size(built_library):
    """1 048 576 pose pairs (default_rng(0), U([-10,10]^2 x [-pi,pi))), step 0.1, every word
    sampled and collision-checked against the canonical 8-row orchard (body only, boundary_check=False).
    Full size: size-independent properties + determinism; a 1024-pair subsample against the oracle
    (word set / order / npts exact, lengths 1e-5, collision flags exact)."""
    import torch
    from headland_trajectory_planning_b200 import ops
    from headland_trajectory_planning_b200.env_batch import EnvBatch, make_record
    n = 1 << 20
    rng = np.random.default_rng(0)
    sg = np.empty((n, 6))
    sg[:, [0, 1, 3, 4]] = rng.uniform(-10, 10, (n, 4))
    sg[:, [2, 5]] = rng.uniform(-math.pi, math.pi, (n, 2))
    rows = H.canonical_rows()
    (o_env, o_car, _), (g_env, g_car, _) = H.make_pair(rows)
    envs = EnvBatch([make_record(g_env, g_car)])
    d_sg = torch.from_numpy(sg).cuda()
    ops.rs_all_paths(d_sg[:4096], MAXC, 0.1, envs=envs, flags=ops.CHECK_OBSTACLES, want_order=False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    words, count, _ = ops.rs_all_paths(d_sg, MAXC, 0.1, envs=envs, flags=ops.CHECK_OBSTACLES, want_order=False)
    b.record()
    torch.cuda.synchronize()
    print(f"config 3: {n} pairs in {a.elapsed_time(b):.1f} ms = {n / a.elapsed_time(b) / 1e3:.2f} M pairs/s")
    wi = words.view(torch.int32).reshape(n, 46, 28)
    wf = words.view(torch.float64).reshape(n, 46, 14)
    valid = torch.arange(46, device="cuda")[None, :] < count[:, None]
    assert int(count.min()) >= 1 and int(count.max()) <= 46
    nseg, npts, coll = wi[..., 1], wi[..., 2], wi[..., 3]
    L, lens = wf[..., 2], wf[..., 4:9]
    assert bool(((nseg >= 3) & (nseg <= 5))[valid].all())
    assert bool(((coll == 0) | (coll == 1))[valid].all())
    seg_ok = torch.arange(5, device="cuda")[None, None, :] < nseg[..., None]
    Lsum = (lens.abs() * seg_ok).sum(-1)
    assert bool(((Lsum - L).abs() <= 1e-9 * (1 + L))[valid].all())            # L = sum |l_i| (metric)
    diff = (npts.double() - L / 0.1)[valid]                                     # one sample per step + a few per segment end
    print("npts - L/step: min %.2f max %.2f" % (float(diff.min()), float(diff.max())))
    assert float(diff.min()) > -6.0 and float(diff.max()) < 13.0            # |npts - L/step| <= a sample or two per segment
    assert bool((L[valid] < 1000.0 / MAXC).all()) and bool((L[valid] >= 0.01 / MAXC - 1e-12).all())   # MAX_LENGTH, assert L >= 0.01
    words2, count2, _ = ops.rs_all_paths(d_sg, MAXC, 0.1, envs=envs, flags=ops.CHECK_OBSTACLES, want_order=False)
    assert torch.equal(count, count2) and torch.equal(words, words2)                                   # deterministic
    sub = np.arange(0, n, n // 1024)[:1024]
    w = ops.rs_words_to_host(words[torch.from_numpy(sub).cuda()])
    cnt = count.cpu().numpy()[sub]
    for a_i, i in enumerate(sub):
        ref = rs_port.calc_all_paths(*sg[i], MAXC, 0.1)
        assert cnt[a_i] == len(ref)
        for k, p in enumerate(ref):
            assert w["cand"][a_i, k] == p.cand and w["npts"][a_i, k] == len(p.x)
            np.testing.assert_allclose(w["len"][a_i, k, :len(p.lengths)], p.lengths, rtol=RTOL, atol=1e-9)
            want = not o_env.check_path_feasibility(o_car, np.array([p.x, p.y, p.yaw]).T, boundary_check=False)
            assert bool(w["collide"][a_i, k]) == want, (i, k)


def test_config1_fishtail_notebook_case(built_library):
    """BASELINE config 1 = test/classic_planner.ipynb cells 3-4, 6, 10-11 through the CUDA path: the notebook orchard
    (np.random.seed(1), l_std = 1.0), the fish-tail Reeds-Shepp query between the safe start / end poses, every word
    checked with boundary_check=False -> exactly one feasible word, an R-L-R (cell 11 stdout).  Both ways a caller
    can do it: hl_rs_all_paths with the environment (per-word collision flag), and calc_all_paths +
    check_path_feasibility like the notebook."""
    import math
    import numpy as np
    import hl_helpers as H
    from headland_trajectory_planning_b200 import ops
    from headland_trajectory_planning_b200.utils import reeds_shepp as rs_curves
    rows = H.canonical_rows(l_std=1.0)
    (o_env, o_car, _), (g_env, g_car, _) = H.make_pair(rows, axle_to_front=2.85)
    sx = -1.30805046
    sg = np.array([[sx, 3.75, math.pi, sx, 8.75, 0.0]])
    envs = g_env._env_batch(g_car)
    words, count, _ = ops.rs_all_paths(sg, g_car.curvature, 0.1, envs=envs, flags=ops.CHECK_OBSTACLES)
    w = ops.rs_words_to_host(words)[0][:int(count[0].item())]
    letters = ["".join(rs_curves.row_letters(int(c))) for c in w["cand"]]
    assert letters == ["LRL", "RLR", "RLR"]
    assert [l for l, c in zip(letters, w["collide"]) if not c] == ["RLR"]
    feas = []
    for p in rs_curves.calc_all_paths(sx, 3.75, math.pi, sx, 8.75, 0.0, g_car.curvature, 0.1):
        traj = np.array([p.x, p.y, p.yaw]).T
        ok = g_env.check_path_feasibility(g_car, traj, boundary_check=False)
        assert ok == o_env.check_path_feasibility(o_car, traj, boundary_check=False)
        if ok:
            feas.append("".join(p.ctypes))
    assert feas == ["RLR"]
