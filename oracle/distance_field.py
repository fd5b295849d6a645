"""Restatement of ``holonomic_costs_with_obstacles``
(``path_planner/utils/a_star_utils.py:75-142``): Dijkstra from the goal cell with a
binary heap of ``(cost, (i, j))`` tuples, 8 ("King", :8-21) or 5 ("Pawn", :24-34) moves,
edge cost ``hypot(di, dj)`` (:68-70), no decrease-key (:131), costs copied out of the
closed set into a float64 matrix initialised to ``inf`` (:138-140).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PINNED: ``oracle/gen_golden.py`` checks
it bit-for-bit against the reference's own module (importable here), including the
index wrap-around quirk (validity is ``abs(index) < dim`` and Python negative indexing
reads the aliased cell, :49-64): node coordinates range over -(W-1)..W-1 and aliases
write the same matrix cell, last closed alias wins.  State lives in offset arrays instead
of per-node objects; the heap entries and tie order are the reference's.
"""
import heapq
import math

import numpy as np

KING = ((-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1))
PAWN = ((-1, 0), (0, 1), (-1, 1), (1, 1), (1, 0))


def holonomic_costs_with_obstacles(goal_index, obstacles, motion_type="King"):
    moves = KING if motion_type == "King" else PAWN
    step_cost = [math.hypot(m[0], m[1]) for m in moves]
    W, H = obstacles.shape
    occ = np.asarray(obstacles).astype(bool)
    # node (i, j) with -W < i < W, -H < j < H lives at [i + W - 1, j + H - 1]
    OPEN, CLOSED = 1, 2
    state = np.zeros((2 * W - 1, 2 * H - 1), dtype=np.int8)
    cost = np.zeros((2 * W - 1, 2 * H - 1), dtype=np.float64)
    gi, gj = goal_index
    state[gi + W - 1, gj + H - 1] = OPEN
    n_open = 1
    heap = [(0, (gi, gj))]
    out = np.full((W, H), np.inf, dtype=np.float64)
    order = []
    while n_open:
        _, (ci, cj) = heapq.heappop(heap)
        state[ci + W - 1, cj + H - 1] = CLOSED
        n_open -= 1
        ccost = cost[ci + W - 1, cj + H - 1]
        order.append((ci, cj))
        for (di, dj), w in zip(moves, step_cost):
            ni, nj = ci + di, cj + dj
            if abs(ni) >= W or abs(nj) >= H:
                continue
            if occ[ni, nj]:                      # negative indices alias, as in the reference
                continue
            s = state[ni + W - 1, nj + H - 1]
            if s == CLOSED:
                continue
            c = ccost + w
            if s == OPEN:
                if c < cost[ni + W - 1, nj + H - 1]:
                    cost[ni + W - 1, nj + H - 1] = c
            else:
                state[ni + W - 1, nj + H - 1] = OPEN
                cost[ni + W - 1, nj + H - 1] = c
                n_open += 1
                heapq.heappush(heap, (c, (ni, nj)))
    for (i, j) in order:                        # dict insertion order == closing order
        out[i, j] = cost[i + W - 1, j + H - 1]
    return out


def synthetic_grid(n, seed=1, n_rows=None, n_blocks=None):
    """Config-4 style occupancy grid, n x n: occupied 1-cell border, tree
    rows as 8-cell-wide bars with headland gaps, random 6x6 blocks.  Returns (bool grid, goal)."""
    rng = np.random.default_rng(seed)
    occ = np.zeros((n, n), dtype=bool)
    occ[0, :] = occ[-1, :] = occ[:, 0] = occ[:, -1] = True
    n_rows = n_rows if n_rows is not None else max(2, n // 64)
    gap = max(8, n // 10)
    pitch = max(16, (n - 2 * gap) // n_rows)
    for r in range(n_rows):
        j = gap + r * pitch
        occ[gap:n - gap, j:j + min(8, max(1, pitch // 3))] = True
    n_blocks = n_blocks if n_blocks is not None else max(4, (n * n) // 8192)
    for _ in range(n_blocks):
        a, b = rng.integers(1, n - 7, 2)
        occ[a:a + 6, b:b + 6] = True
    c = n // 2
    free = np.argwhere(~occ)
    goal = tuple(int(v) for v in free[np.argmin(np.abs(free[:, 0] - c) + np.abs(free[:, 1] - c))])
    return occ, goal
