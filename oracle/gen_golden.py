"""Generate the committed golden fixtures under ``tests/golden/`` (run in the BUILD
container, where ``/root/reference`` exists):

    python -m oracle.gen_golden rs df        # from the reference's OWN modules (pins the ports)
    python -m oracle.gen_golden astar N      # oracle Hybrid A* on config-5 scenarios 0..N-1

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

* ``rs_golden.npz``  -- ``reeds_shepp.calc_all_paths`` of the REFERENCE on 400 seeded pose
  pairs (+ the 3 KATs of SURVEY.md 8c): word letters, lengths, L, sample counts, and the
  sampled states of the first 40 pairs.  The port must reproduce them bit-for-bit.
* ``df_golden.npz``  -- ``a_star_utils.holonomic_costs_with_obstacles`` of the REFERENCE on
  seeded grids (open and closed borders, King and Pawn).
* ``astar_golden.npz`` -- the ORACLE's search results (status, counter, expanded keys, path)
  on config-5 scenarios: lets the GPU tests check node-sequence parity at a scale the oracle
  cannot run on the GPU box in test time.  (parity unpinned at the GEOS/heapdict boundary.)
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
MAXC = math.tan(0.55) / 1.9
LET = {"S": 0, "L": 1, "R": 2}


def rs_cases():
    rng = np.random.default_rng(20261018)
    sg = np.empty((400, 6))
    sg[:, [0, 1, 3, 4]] = rng.uniform(-10, 10, (400, 4))
    sg[:, [2, 5]] = rng.uniform(-math.pi, math.pi, (400, 2))
    kats = np.array([[0, 0, 0, 3, 4, 1.0], [-1.30805046, 3.75, math.pi, -1.30805046, 8.75, 0],
                     [1, 2, -2, -4, 1.5, 2.5]], dtype=np.float64)
    sg = np.vstack([kats, sg])
    steps = np.where(np.arange(len(sg)) % 2 == 0, 0.1, 0.2)
    steps[2] = 0.2
    steps[:2] = 0.1
    return sg, steps


def pack_paths(paths, keep_states):
    letters = np.full((46, 5), -1, dtype=np.int8)
    lens = np.zeros((46, 5))
    L = np.zeros(46)
    npts = np.zeros(46, dtype=np.int32)
    states = []
    for k, p in enumerate(paths):
        for s, c in enumerate(p.ctypes):
            letters[k, s] = LET[c]
        lens[k, :len(p.lengths)] = p.lengths
        L[k] = p.L
        npts[k] = len(p.x)
        if keep_states:
            states.append(np.stack([p.x, p.y, p.yaw, np.asarray(p.cs, dtype=np.float64),
                                    np.asarray(p.directions, dtype=np.float64)], axis=1))
    return len(paths), letters, lens, L, npts, states


def gen_rs():
    from . import ref_loader, rs_port
    ref = ref_loader.load("reeds_shepp")
    sg, steps = rs_cases()
    out = dict(sg=sg, steps=steps, count=[], letters=[], lens=[], L=[], npts=[])
    states_all = []
    for i, (q, st) in enumerate(zip(sg, steps)):
        paths = ref.calc_all_paths(*q, MAXC, st)
        mine = rs_port.calc_all_paths(*q, MAXC, st)
        assert len(paths) == len(mine)
        for a, b in zip(paths, mine):
            assert a.ctypes == b.ctypes and a.lengths == b.lengths and a.L == b.L and a.x == b.x and a.y == b.y \
                and a.yaw == b.yaw and a.cs == b.cs and a.directions == b.directions, i
        n, letters, lens, L, npts, states = pack_paths(paths, i < 43)
        out["count"].append(n); out["letters"].append(letters); out["lens"].append(lens)
        out["L"].append(L); out["npts"].append(npts)
        states_all += states
    np.savez_compressed(os.path.join(GOLD, "rs_golden.npz"), sg=sg, steps=steps, count=np.array(out["count"]),
                        letters=np.array(out["letters"]), lens=np.array(out["lens"]), L=np.array(out["L"]),
                        npts=np.array(out["npts"]), states=np.concatenate(states_all),
                        states_len=np.array([len(s) for s in states_all]))
    print("rs_golden.npz:", len(sg), "pairs,", int(np.sum(out["count"])), "words; port == reference bit-for-bit")


def gen_df():
    from . import ref_loader, distance_field as DF
    ref = ref_loader.load("a_star_utils")
    rng = np.random.default_rng(7)
    grids, goals, motions, outs = [], [], [], []
    for trial in range(10):
        n = int(rng.integers(10, 40))
        occ = rng.random((n, n)) < 0.25
        if trial % 3 != 2:
            occ[0, :] = occ[-1, :] = occ[:, 0] = occ[:, -1] = True
        free = np.argwhere(~occ)
        g = tuple(int(v) for v in free[rng.integers(len(free))])
        for mt in ("King", "Pawn"):
            a = ref.holonomic_costs_with_obstacles(g, occ, mt)
            b = DF.holonomic_costs_with_obstacles(g, occ, mt)
            assert np.array_equal(a, b), (trial, mt)
            pad = np.full((40, 40), np.nan)
            pad[:n, :n] = a
            po = np.zeros((40, 40), dtype=bool)
            po[:n, :n] = occ
            grids.append(po); goals.append((n,) + g); motions.append(0 if mt == "King" else 1); outs.append(pad)
    occ, g = DF.synthetic_grid(96, seed=3)
    a = ref.holonomic_costs_with_obstacles(g, occ, "King")
    assert np.array_equal(a, DF.holonomic_costs_with_obstacles(g, occ, "King"))
    np.savez_compressed(os.path.join(GOLD, "df_golden.npz"), grids=np.array(grids), goals=np.array(goals),
                        motions=np.array(motions), outs=np.array(outs), big_occ=occ, big_goal=np.array(g), big_out=a)
    print("df_golden.npz:", len(grids), "small grids + 96x96; port == reference bit-for-bit")


def _astar_one(i):
    from . import baseline as OB
    sys.path.insert(0, os.path.dirname(HERE))
    from headland_trajectory_planning_b200 import scenarios as SC
    import contextlib
    import io
    sp = SC.scenario_spec(i)
    feas = OB.candidate_feasibility(sp)
    scn = SC.finalize(sp, feas)
    with contextlib.redirect_stdout(io.StringIO()):
        r = OB.run_scenario(scn)
    r["feas"] = np.array(feas, dtype=bool)
    r["goal"] = scn["goal"]
    return r


def gen_astar(n):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(_astar_one, range(n), chunksize=1)
    status_code = {"ok": 0, "start_goal_blocked": 1, "open_empty": 2, "max_nodes": 3}
    exp = np.concatenate([r["expanded"] for r in res]) if res else np.zeros((0, 3), np.int32)
    path = np.concatenate([np.stack([r["x"], r["y"], r["yaw"], r["ks"], r["dirs"].astype(np.float64)], axis=1)
                           .reshape(-1, 5) for r in res])
    np.savez_compressed(
        os.path.join(GOLD, "astar_golden.npz"),
        index=np.array([r["index"] for r in res]), status=np.array([status_code[r["status"]] for r in res]),
        counter=np.array([r["counter"] for r in res]), n_expanded=np.array([len(r["expanded"]) for r in res]),
        expanded=exp.astype(np.int32), path_len=np.array([len(r["x"]) for r in res]), path=path,
        feas=np.array([r["feas"] for r in res]), goal=np.array([r["goal"] for r in res]),
        seconds=np.array([r["seconds"] for r in res]),
        rs_poses=np.array([r["stats"]["rs_poses"] for r in res]),
        primitive_poses=np.array([r["stats"]["primitive_poses"] for r in res]))
    print("astar_golden.npz:", n, "scenarios; counters", [r["counter"] for r in res][:20], "...")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    args = sys.argv[1:]
    if "rs" in args:
        gen_rs()
    if "df" in args:
        gen_df()
    if "astar" in args:
        gen_astar(int(args[args.index("astar") + 1]))
