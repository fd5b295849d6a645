"""K6 / K7: distance field bit-exact vs the pinned oracle port (float64 fixed point), and
the occupancy-grid footprint check vs a brute-force cell-by-cell SAT."""
import math

import numpy as np
import pytest

from oracle import distance_field as DF
from oracle import geometry as geo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,motion", [(64, "King"), (200, "King"), (200, "Pawn"), (333, "King")])
def test_distance_field_bit_exact(built_library, n, motion):
    from headland_trajectory_planning_b200.utils import a_star_utils
    occ, goal = DF.synthetic_grid(n, seed=n)
    want = DF.holonomic_costs_with_obstacles(goal, occ, motion)
    got = a_star_utils.holonomic_costs_with_obstacles(goal, occ, motion)
    assert np.array_equal(np.isinf(want), np.isinf(got))
    assert np.array_equal(want, got)          # float64 least fixed point: bit-exact
    assert np.isfinite(want).sum() > n


def test_wrap_around_quirk_matches_reference(built_library):
    """Grids with free border cells: the reference's index wrap-around (a_star_utils.py:49-64) is reproduced -- the
    oracle port is pinned bit for bit on the reference's own module, known answers included: a free 8 x 8 grid with
    goal (1, 1) gives 11.3137... AT THE GOAL CELL in King mode and an all-inf column 0 in Pawn mode."""
    from headland_trajectory_planning_b200.utils import a_star_utils
    free = np.zeros((8, 8), dtype=bool)
    king = a_star_utils.holonomic_costs_with_obstacles((1, 1), free, "King")
    assert np.array_equal(king, DF.holonomic_costs_with_obstacles((1, 1), free, "King"))
    assert abs(king[1, 1] - 8 * math.sqrt(2.0)) < 1e-12
    pawn = a_star_utils.holonomic_costs_with_obstacles((1, 1), free, "Pawn")
    assert np.array_equal(pawn, DF.holonomic_costs_with_obstacles((1, 1), free, "Pawn"))
    assert np.isinf(pawn[:, 0]).all()
    rng = np.random.default_rng(0)
    for _ in range(60):
        n, m = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        occ = rng.random((n, m)) < rng.choice([0.0, 0.1, 0.25])
        cells = np.argwhere(~occ) if rng.random() < 0.8 else np.argwhere(np.ones_like(occ))    # sometimes an occupied goal
        g = tuple(int(v) for v in cells[rng.integers(len(cells))])
        for motion in ("King", "Pawn"):
            want = DF.holonomic_costs_with_obstacles(g, occ, motion)
            got = a_star_utils.holonomic_costs_with_obstacles(g, occ, motion)
            assert np.array_equal(want, got), (n, m, g, motion)
    # a half-open map (two occupied sides) and a closed border with the goal ON the border
    occ, goal = DF.synthetic_grid(96, seed=3)
    occ[0, :] = False
    occ[:, -1] = False
    assert np.array_equal(DF.holonomic_costs_with_obstacles(goal, occ, "King"),
                          a_star_utils.holonomic_costs_with_obstacles(goal, occ, "King"))
    occ, _ = DF.synthetic_grid(64, seed=5)
    assert np.array_equal(DF.holonomic_costs_with_obstacles((0, 10), occ, "King"),
                          a_star_utils.holonomic_costs_with_obstacles((0, 10), occ, "King"))


def test_distance_field_properties_full_size(built_library):
    """BASELINE config 4 size: 4096 x 4096.  The oracle would need ~45 min, so check
    size-independent properties: goal is 0, every finite cell satisfies the fixed-point
    equation with its best neighbour, occupied cells are inf."""
    import torch
    from headland_trajectory_planning_b200 import ops
    n = 4096
    occ, goal = DF.synthetic_grid(n, seed=1)
    out, sweeps = ops.distance_field(occ, goal, "King")
    D = out.cpu().numpy()
    assert D[goal] == 0.0
    assert np.isinf(D[occ & ~(np.arange(n)[:, None] == goal[0]) | occ & ~(np.arange(n)[None, :] == goal[1])]).all()
    best = np.full_like(D, np.inf)
    r2 = math.hypot(1, 1)
    for di, dj, w in [(-1, 0, 1.0), (1, 0, 1.0), (0, -1, 1.0), (0, 1, 1.0), (-1, -1, r2), (-1, 1, r2), (1, -1, r2), (1, 1, r2)]:
        sh = np.full_like(D, np.inf)
        a0, a1 = max(0, di), n + min(0, di)
        b0, b1 = max(0, dj), n + min(0, dj)
        sh[a0:a1, b0:b1] = D[a0 - di:a1 - di, b0 - dj:b1 - dj] + w
        best = np.minimum(best, sh)
    free = ~occ
    free[goal] = False
    assert np.array_equal(D[free], best[free])
    assert sweeps > 10


def test_grid_footprint_vs_bruteforce(built_library):
    from headland_trajectory_planning_b200 import ops
    rng = np.random.default_rng(4)
    W, H, res = 160, 120, 0.1
    occ = rng.random((W, H)) < 0.02
    ext = (-0.55, 3.0, -0.74, 0.74)
    n = 3000
    poses = np.stack([rng.uniform(-1, W * res + 1, n), rng.uniform(-1, H * res + 1, n), rng.uniform(-math.pi, math.pi, n)], 1)
    bits = ops.grid_pack(occ)
    got = ops.grid_footprint_check(bits, occ.shape, res, poses, ext).cpu().numpy().astype(bool)
    want = np.zeros(n, dtype=bool)
    ii, jj = np.nonzero(occ)
    corners = geo.rect_corners(poses, ext)
    for i, j in zip(ii, jj):
        sq = geo.ccw(np.array([[i * res, j * res], [(i + 1) * res, j * res], [(i + 1) * res, (j + 1) * res], [i * res, (j + 1) * res]]))
        want |= geo.rects_hit_convex(poses, ext, sq, corners)
    assert np.array_equal(got, want)
    assert 0.05 < want.mean() < 0.95
