"""Generate the committed golden fixtures under ``tests/golden/`` (run in the BUILD
container, where ``/root/reference`` exists):

    python -m oracle.gen_golden rs df        # from the reference's OWN modules (pins the ports)
    python -m oracle.gen_golden astar N      # oracle Hybrid A* on config-5 scenarios 0..N-1
    python -m oracle.gen_golden astar_ref N  # the REFERENCE's own search loop on scenarios 0..N-1
    python -m oracle.gen_golden ypark N      # the REFERENCE's own Y-park sweep on N scenarios
    python -m oracle.gen_golden offset N     # the REFERENCE's own get_offset_pose on N scenarios
    python -m oracle.gen_golden refpath      # the REFERENCE's own get_init_ref_path on the golden planner paths

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

* ``rs_golden.npz``  -- ``reeds_shepp.calc_all_paths`` of the REFERENCE on 400 seeded pose
  pairs (+ the 3 KATs of SURVEY.md 8c): word letters, lengths, L, sample counts, and the
  sampled states of the first 40 pairs.  The port must reproduce them bit-for-bit.
* ``df_golden.npz``  -- ``a_star_utils.holonomic_costs_with_obstacles`` of the REFERENCE on
  seeded grids (open and closed borders, King and Pawn).
* ``astar_golden.npz`` -- the ORACLE's search results (status, counter, expanded keys, path)
  on config-5 scenarios: lets the GPU tests check node-sequence parity at a scale the oracle
  cannot run on the GPU box in test time.  (parity unpinned at the GEOS/heapdict boundary.)
* ``astar_ref_golden.npz`` -- ``HybridAStarSearch.hybrid_a_star_search`` of the REFERENCE
  (``path_planner/hybrid_a_star_search.py`` imported unmodified through
  ``ref_loader.load_planner``: shapely/dubins stubbed, ``heapdict`` -> ``oracle.heapdict_port``)
  fed the oracle's environment / car / heuristic objects: pins the oracle's search loop, costs,
  rollout and Reeds-Shepp shot against the reference itself; only the GEOS predicates and the
  heapdict port stay restated.
* ``ypark_golden.npz`` -- ``search_y_type_parking_path`` of the REFERENCE
  (``headland_path_planning.py:382-451``, same loader) on config-5 environments with the
  notebook's and the function's default parameter sets.
* ``offset_golden.npz`` -- ``safety_forward_path_plan.get_offset_pose`` of the REFERENCE (:248-283) with the
  reference's ``CarModel.calculate_motion_path`` (car_model.py:202-234) grafted onto the oracle car.
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


def gen_astar_full(n):
    """Compact golden for ALL scenarios of the benchmarked sweep (bench.py: config 5, 0..n-1): the discrete outcome of
    every search -- status, counter, number of expanded nodes, CRC-32 of the expanded-key sequence (int32 [k, 3]
    bytes), path length, the oracle's pose-check tally -- plus two float64 checksums of the path (sum of x, sum of
    y).  tests/test_astar_gpu.py::test_full_sweep_golden compares the GPU sweep with it scenario by scenario."""
    import multiprocessing as mp
    import zlib
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(_astar_one, range(n), chunksize=4)
    status_code = {"ok": 0, "start_goal_blocked": 1, "open_empty": 2, "max_nodes": 3}
    np.savez_compressed(
        os.path.join(GOLD, "astar_full_golden.npz"),
        status=np.array([status_code[r["status"]] for r in res], dtype=np.int8),
        counter=np.array([r["counter"] for r in res], dtype=np.int32),
        n_expanded=np.array([len(r["expanded"]) for r in res], dtype=np.int32),
        keys_crc=np.array([zlib.crc32(np.ascontiguousarray(r["expanded"], dtype=np.int32).tobytes()) for r in res], dtype=np.uint32),
        path_len=np.array([len(r["x"]) for r in res], dtype=np.int32),
        path_sum=np.array([[np.sum(r["x"]), np.sum(r["y"])] for r in res], dtype=np.float64),
        feas_crc=np.array([zlib.crc32(np.packbits(r["feas"]).tobytes()) for r in res], dtype=np.uint32),
        goal=np.array([r["goal"] for r in res]),
        pose_checks_ref=np.array([r["stats"]["rs_poses"] + r["stats"]["primitive_poses"] for r in res], dtype=np.int64))
    print("astar_full_golden.npz:", n, "scenarios; status histogram", np.bincount([status_code[r["status"]] for r in res]))


SAMPLING_CASES = [
    # (l_std, slope_deg, row_width, headland_width, aux, axle_to_front, start_row, end_row, dubins accuracy, rs accuracy)
    (0.0, 10.0, 2.5, 9.0, "mower", 3.0, 1, 4, 0.5, 1.0),
    (0.5, 10.0, 2.5, 8.0, "mower", 2.85, 2, 5, 0.6, 1.2),
    (0.0, 0.0, 3.0, 10.0, "sprayer", 2.85, 1, 3, 0.7, 1.5),
    (1.0, 5.0, 2.8, 9.0, "pruner", 3.0, 0, 2, 0.6, 1.2),
    (0.0, 12.0, 2.5, 7.0, "none", 3.5, 3, 4, 0.5, 1.0),
]
_AUX = {"mower": [[[-1.84, 0.5], 1.0, 1.1]], "pruner": [[[3.259, -0.175], 1.325, 0.3]],
        "sprayer": [[[-2.1, 1.0 / 2], 1.0, 1.22], [[-1.0, 4.3 / 2], 0.4, 0.5], [[-1.0, -4.3 / 2 + 0.4], 0.4, 0.5]], "none": []}


def _sampling_objects(case):
    from . import planner as OP
    l_std, slope, rw, hw, aux, atf = case[:6]
    np.random.seed(1)
    rows = OP.create_tree_rows(8, rw, 20, slope_angle=math.radians(slope), l_std=l_std)
    env = OP.OrchardGeometryEnvironment(rows, [], tree_width=0.3, headland_width=hw)
    car = OP.CarModel(max_steer=0.55, axle_to_front=atf, axle_to_back=0.55, width=1.48, aux_poly_features=_AUX[aux],
                      with_aux=bool(_AUX[aux]))
    return rows, env, car


def _sampling_one(k):
    """The REFERENCE's own sampling / turn functions (path_planner/safety_forward_path_plan.py, loaded unmodified with
    ``dubins`` := oracle.dubins_port and the oracle's duck-typed environment / car) on one case."""
    from . import ref_loader
    case = SAMPLING_CASES[k]
    rows, env, car = _sampling_objects(case)
    ref = ref_loader.load_planner("safety_forward_path_plan", dubins_port=True)
    s_row, e_row, acc_d, acc_r = case[6:]
    out = {"case": k}

    def pack(r):
        return np.full(8, np.nan) if r is None else np.concatenate([np.asarray(r[0], float), np.asarray(r[1], float), [r[2], r[3]]])
    out["dubins"] = pack(_quiet(ref.sample_start_end_pose_for_dubins, rows, s_row, e_row, car, env, accuracy=acc_d))
    out["rs"] = pack(_quiet(ref.sample_start_end_pose_for_reeds_shepp, rows, s_row, e_row, car, env, accuracy=acc_r))
    out["circle"] = pack(_quiet(ref.sample_start_end_pose_for_circle_back, rows, s_row, s_row + 1, car, env))
    start, end = ref.get_start_end_pose(rows, s_row, s_row + 1)
    out["circle_path"] = np.asarray(_quiet(ref.get_circle_back_path_full, np.asarray(start, float), end, 1.0 / car.curvature, car))
    sb = ref.get_base_pose(s_row, rows, 0.3, pose_type=ref.LEAVE_POSE)
    eb = ref.get_base_pose(e_row, rows, 0.5, pose_type=ref.ENTER_POSE)
    sb[0] -= 1.5
    eb[0] -= 1.0
    dp = np.asarray(_quiet(ref.get_dubins_path_full, sb, eb, 1.0 / car.curvature))
    out["dubins_pair"] = np.concatenate([sb, eb])
    out["dubins_path"] = dp
    out["dubins_min_dist"] = env.get_min_distance_to_boundary(car, dp[:, :3], with_aux=True)
    out["dubins_feasible"] = env.check_path_feasibility(car, dp[:, :3], boundary_check=False, aux_check=True)
    return out


def gen_sampling():
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(len(SAMPLING_CASES), os.cpu_count())) as pool:
        res = pool.map(_sampling_one, range(len(SAMPLING_CASES)), chunksize=1)
    np.savez_compressed(
        os.path.join(GOLD, "sampling_golden.npz"),
        dubins=np.array([r["dubins"] for r in res]), rs=np.array([r["rs"] for r in res]),
        circle=np.array([r["circle"] for r in res]),
        circle_path=np.concatenate([r["circle_path"] for r in res]),
        circle_path_len=np.array([len(r["circle_path"]) for r in res]),
        dubins_pair=np.array([r["dubins_pair"] for r in res]),
        dubins_path=np.concatenate([r["dubins_path"] for r in res]),
        dubins_path_len=np.array([len(r["dubins_path"]) for r in res]),
        dubins_min_dist=np.array([r["dubins_min_dist"] for r in res]),
        dubins_feasible=np.array([r["dubins_feasible"] for r in res]))
    print("sampling_golden.npz:", len(res), "cases;", [tuple(np.round(r["dubins"][6:], 2)) for r in res],
          [tuple(np.round(r["rs"][6:], 2)) for r in res], [tuple(np.round(r["circle"][6:], 2)) for r in res])


YPARK_PARAM_SETS = [
    # (max_steer_backward, max_steer_forward, max_backward_distance, max_forward_distance,
    #  min_forward_distance, min_backward_distance, min_steer_backward, min_steer_forward, step)
    (0.15, 0.55, 3.0, 2.0, 1.0, 1.0, 0.0, 0.5, 0.2),        # test/obca.ipynb cell 9
    (0.35, 0.55, 2.0, 2.5, 1.4, 0.7, 0.22, 0.50, 0.1),      # headland_planner_y_type_park defaults (:131-139)
    (0.4, 0.45, 3.5, 2.0, 1.4, 0.7, 0.3, 0.3, 0.1),         # search_y_type_parking_path defaults (:388-396)
]


def _quiet(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _ypark_one(i):
    from . import planner as OP
    from . import ref_loader
    sys.path.insert(0, os.path.dirname(HERE))
    from headland_trajectory_planning_b200 import scenarios as SC
    H = ref_loader.load_planner("headland_path_planning")
    sp = SC.scenario_spec(i)
    env = OP.OrchardGeometryEnvironment(sp["rows"], [], tree_width=sp["tree_width"], headland_width=sp["headland_width"])
    car = OP.CarModel(**sp["car"])
    out = []
    # the row-enter pose pulled back into the headland by 0 / 1.5 / 3 m so that the sweeps end in a mix of
    # early hits, late hits and failures
    shift = (0.0, 1.5, 3.0)[i % 3]
    end = np.array(sp["end"], dtype=np.float64)
    end[0] -= shift * math.cos(end[2])
    end[1] -= shift * math.sin(end[2])
    for k, ps in enumerate(YPARK_PARAM_SETS):
        bdir = H.get_backward_steer_dir_for_y_type_parking(sp["start"], end)
        path, par = _quiet(H.search_y_type_parking_path, car, env, end, bdir, -bdir, *ps[:8], step_size=ps[8], debug=True)
        out.append(dict(index=i, pset=k, bdir=float(bdir), found=len(par) > 0, end=end,
                        par=np.array(par if len(par) else [np.nan] * 4, dtype=np.float64),
                        path=np.asarray(path, dtype=np.float64).reshape(-1, 5)))
    return out


def gen_ypark(n):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = [r for rs in pool.map(_ypark_one, range(n), chunksize=1) for r in rs]
    np.savez_compressed(
        os.path.join(GOLD, "ypark_golden.npz"), param_sets=np.array(YPARK_PARAM_SETS),
        index=np.array([r["index"] for r in res]), pset=np.array([r["pset"] for r in res]),
        bdir=np.array([r["bdir"] for r in res]), found=np.array([r["found"] for r in res]),
        end=np.array([r["end"] for r in res]),
        par=np.array([r["par"] for r in res]), path_len=np.array([len(r["path"]) for r in res]),
        path=np.concatenate([r["path"] for r in res]))
    print("ypark_golden.npz:", len(res), "sweeps;", int(sum(r["found"] for r in res)), "found a path")


def _astar_ref_one(i):
    from . import baseline as OB
    from . import planner as OP
    from . import ref_loader
    sys.path.insert(0, os.path.dirname(HERE))
    from headland_trajectory_planning_b200 import scenarios as SC
    R = ref_loader.load_planner("hybrid_a_star_search")
    sp = SC.scenario_spec(i)
    scn = SC.finalize(sp, OB.candidate_feasibility(sp))
    env = OP.OrchardGeometryEnvironment(scn["rows"], [], tree_width=scn["tree_width"], headland_width=scn["headland_width"])
    car = OP.CarModel(**scn["car"])
    heur = OP.ReferenceLineHeuristic(scn["waypoints"], scn["goal"], car)
    s = _quiet(R.HybridAStarSearch, scn["start"], scn["goal"], env, car, heur, motion_type="King",
               plan_resolution=scn["step_size"])
    x, y, yaw, dirs, ks, counter = _quiet(s.hybrid_a_star_search, max_nodes=400)
    path = np.stack([np.asarray(x, float), np.asarray(y, float), np.asarray(yaw, float), np.asarray(ks, float),
                     np.asarray(dirs, float)], axis=1).reshape(-1, 5)
    return dict(index=i, counter=int(counter), path=path)


def gen_astar_ref(n):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(_astar_ref_one, range(n), chunksize=1)
    np.savez_compressed(os.path.join(GOLD, "astar_ref_golden.npz"), index=np.array([r["index"] for r in res]),
                        counter=np.array([r["counter"] for r in res]), path_len=np.array([len(r["path"]) for r in res]),
                        path=np.concatenate([r["path"] for r in res]))
    print("astar_ref_golden.npz:", n, "scenarios run by the reference's own search loop; counters",
          [r["counter"] for r in res][:24])


PAWN_MAX_NODES = 120


def _pawn_one(i):
    """Pawn mode (forward primitives + Dubins goal extension) on config-5 scenario i, step 0.2 m: the REFERENCE's own
    hybrid_a_star_search.py (dubins := oracle.dubins_port, heapdict := the port, oracle geometry) next to the oracle's
    restatement; both must agree before the golden is written.  The expanded-key sequence comes from the oracle."""
    from . import baseline as OB
    from . import planner as OP
    from . import ref_loader
    sys.path.insert(0, os.path.dirname(HERE))
    from headland_trajectory_planning_b200 import scenarios as SC
    R = ref_loader.load_planner("hybrid_a_star_search", dubins_port=True)
    sp = SC.scenario_spec(i)
    scn = SC.finalize(sp, OB.candidate_feasibility(sp))
    env = OP.OrchardGeometryEnvironment(scn["rows"], [], tree_width=scn["tree_width"], headland_width=scn["headland_width"])
    car = OP.CarModel(**scn["car"])
    heur = OP.ReferenceLineHeuristic(scn["waypoints"], scn["goal"], car)
    s = _quiet(R.HybridAStarSearch, scn["start"], scn["goal"], env, car, heur, motion_type="Pawn", plan_resolution=scn["step_size"])
    x, y, yaw, dirs, ks, counter = _quiet(s.hybrid_a_star_search, max_nodes=PAWN_MAX_NODES)
    o = OP.HybridAStarSearch(scn["start"], scn["goal"], env, car, heur, motion_type="Pawn", plan_resolution=scn["step_size"])
    ox, oy, oyaw, odirs, oks, ocounter = _quiet(o.hybrid_a_star_search, max_nodes=PAWN_MAX_NODES)
    assert counter == ocounter and len(x) == len(ox), (i, counter, ocounter)
    if len(x):
        assert np.array_equal(np.asarray(x, float), np.asarray(ox, float)) and np.array_equal(np.asarray(ks, float), np.asarray(oks, float))
    path = np.stack([np.asarray(x, float), np.asarray(y, float), np.asarray(yaw, float), np.asarray(ks, float),
                     np.asarray(dirs, float)], axis=1).reshape(-1, 5)
    status_code = {"ok": 0, "start_goal_blocked": 1, "open_empty": 2, "max_nodes": 3}
    return dict(index=i, counter=int(counter), path=path, status=status_code[o.status],
                expanded=np.array(o.expanded, dtype=np.int32).reshape(-1, 3), feas=np.array(OB.candidate_feasibility(sp), dtype=bool))


def gen_pawn(n):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(_pawn_one, range(n), chunksize=1)
    np.savez_compressed(os.path.join(GOLD, "pawn_golden.npz"), index=np.array([r["index"] for r in res]),
                        status=np.array([r["status"] for r in res]), counter=np.array([r["counter"] for r in res]),
                        n_expanded=np.array([len(r["expanded"]) for r in res]),
                        expanded=np.concatenate([r["expanded"] for r in res]),
                        path_len=np.array([len(r["path"]) for r in res]), path=np.concatenate([r["path"] for r in res]),
                        feas=np.array([r["feas"] for r in res]), max_nodes=np.array(PAWN_MAX_NODES))
    print("pawn_golden.npz:", n, "scenarios (reference loop == oracle); counters", [r["counter"] for r in res][:32],
          "status", np.bincount([r["status"] for r in res]))


def _offset_one(i):
    from . import planner as OP
    from . import ref_loader
    sys.path.insert(0, os.path.dirname(HERE))
    from headland_trajectory_planning_b200 import scenarios as SC
    S = ref_loader.load_planner("safety_forward_path_plan")
    C = ref_loader.load_planner("car_model")

    class Car(OP.CarModel):
        calculate_motion_path = C.CarModel.calculate_motion_path

    sp = SC.scenario_spec(i)
    env = OP.OrchardGeometryEnvironment(sp["rows"], [], tree_width=sp["tree_width"], headland_width=sp["headland_width"])
    car = Car(**sp["car"])
    turn = S.get_steer_dir_for_enter_calculation(sp["start"], sp["end"])
    out = []
    for steer in (0.55, 0.5):
        for pose, typ in ((sp["start"], S.LEAVE_POSE), (sp["end"], S.ENTER_POSE)):
            d, p, path = _quiet(S.get_offset_pose, pose, typ, turn, car, env, steer_angle=steer)
            out.append(dict(index=i, steer=steer, pose_type=typ, turn=float(turn), init=np.array(pose, float),
                            dist=float(d), pose=np.array(p, float), path=np.asarray(path, float).reshape(-1, 5)))
    # get_start_end_pose_for_reeds_shepp (:300-364) for the rows of this scenario, reference code end to end
    rng = np.random.default_rng(99 + i)
    r0 = int(rng.integers(0, 5)); r1 = r0 + int(rng.integers(1, 3))
    side = int(sp["side"])
    res = _quiet(S.get_start_end_pose_for_reeds_shepp, sp["rows"], r0, r1, car, env, side=side)
    out.append(dict(index=i, steer=-1.0, pose_type=100 + side, turn=float(r0 * 10 + r1), init=np.array(res[0], float),
                    dist=float(res[2]), pose=np.array(res[1], float), path=np.array([[res[3], 0, 0, 0, 0]], float)))
    return out


def gen_offset(n):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = [r for rs in pool.map(_offset_one, range(n), chunksize=1) for r in rs]
    np.savez_compressed(
        os.path.join(GOLD, "offset_golden.npz"), index=np.array([r["index"] for r in res]),
        steer=np.array([r["steer"] for r in res]), pose_type=np.array([r["pose_type"] for r in res]),
        turn=np.array([r["turn"] for r in res]), init=np.array([r["init"] for r in res]),
        dist=np.array([r["dist"] for r in res]), pose=np.array([r["pose"] for r in res]),
        path_len=np.array([len(r["path"]) for r in res]), path=np.concatenate([r["path"] for r in res]))
    print("offset_golden.npz:", len(res), "sweeps; distances", sorted(set(np.round([r["dist"] for r in res], 1)))[:30])


def oge_cases():
    """Orchards for the OBCA obstacle extraction (``OGE_OBCA.py``): slopes, jittered row ends, an explicit field contour;
    turns on both sides over one and over several rows."""
    from . import planner as OP
    cases = []
    for k, (n_rows, width, slope, l_std, headland, tree_w) in enumerate([
            (8, 2.5, 10.0, 0.0, 6.0, 0.3), (8, 2.5, 0.0, 0.0, 7.0, 0.5), (10, 3.0, -8.0, 0.5, 6.5, 0.4),
            (8, 2.2, 15.0, 1.0, 5.5, 0.3), (12, 2.8, 4.0, 0.2, 8.0, 0.5), (8, 2.5, 10.0, 0.3, 6.0, 0.3)]):
        np.random.seed(100 + k)
        rows = OP.create_tree_rows(n_rows, width, 20, slope_angle=math.radians(slope), l_std=l_std)
        contour = []
        if k == 5:                      # explicit field contour: a wavy 14-gon around the rows
            lo, hi = rows[:, :, 1].min() - 3.0, rows[:, :, 1].max() + 3.0
            ys = np.linspace(lo, hi, 7)
            near = np.stack([rows[:, 0, 0].min() - 6.0 + 0.4 * np.sin(ys), ys], axis=1)
            far = np.stack([rows[:, 1, 0].max() + 6.5 + 0.5 * np.cos(ys[::-1]), ys[::-1]], axis=1)
            contour = np.vstack([near, far])
        turns = []
        for (a, b) in ((1, 4), (5, 2), (2, 3), (0, n_rows - 2)):
            for side in (1, -1):
                col = 0 if side == 1 else 1
                ya = 0.5 * (rows[a, col, 1] + rows[a + 1, col, 1])
                yb = 0.5 * (rows[b, col, 1] + rows[b + 1, col, 1])
                x = rows[a, col, 0] - 1.0 * side
                turns.append((np.array([x, ya, math.pi if side == 1 else 0.0]), np.array([x, yb, 0.0 if side == 1 else math.pi]), side))
        cases.append(dict(rows=rows, contour=contour, headland=headland, tree_width=tree_w, turns=turns, seed=500 + k))
    return cases


def oge_outputs(cls, case):
    """The calls of ``test/obca.ipynb`` cell 12 (+ ``get_tree_row_obstacles``) on one orchard; the jitter of the fit line
    (``np.random.uniform`` inside ``create_headland_countour_lines``) is seeded the same way for every implementation."""
    env = cls(case["rows"], [], contour_points=case["contour"], tree_width=case["tree_width"], headland_width=case["headland"])
    np.random.seed(case["seed"])
    boundary = env.create_boundary_polygons()
    out = {"near": list(boundary[0]), "far": list(boundary[1]), "low": list(boundary[2]), "up": list(boundary[3])}
    for t, (start, end, side) in enumerate(case["turns"]):
        rows_polys = env.get_obstacle_tree_rows(start, end)
        out[f"rows{t}"] = list(rows_polys)
        out[f"block{t}"] = list(env.get_tree_row_obstacles(start, end))
        out[f"obca{t}"] = list(env.get_obstacles_for_OBCA(boundary, rows_polys, start, end, side=side))
    return out


def pack_polys(polys):
    return (np.vstack([np.asarray(p, dtype=np.float64).reshape(-1, 2) for p in polys]) if len(polys) else np.zeros((0, 2)),
            np.array([len(p) for p in polys], dtype=np.int64))


def unpack_polys(v, n):
    off = np.concatenate([[0], np.cumsum(n)])
    return [v[off[i]:off[i + 1]] for i in range(len(n))]


def gen_oge():
    """tests/golden/oge_golden.npz: the reference's own ``OGE_OBCA.orchard_environment_OBCA`` (``ref_loader.load_oge_obca``)
    on ``oge_cases()``."""
    from . import ref_loader as RL
    ref = RL.load_oge_obca()
    store = {"n_cases": np.int64(len(oge_cases()))}
    total = 0
    for c, case in enumerate(oge_cases()):
        for name, polys in oge_outputs(ref.orchard_environment_OBCA, case).items():
            v, n = pack_polys(polys)
            store[f"c{c}_{name}_v"] = v
            store[f"c{c}_{name}_n"] = n
            total += len(n)
    np.savez_compressed(os.path.join(GOLD, "oge_golden.npz"), **store)
    print(f"oge: {len(oge_cases())} orchards, {total} polygons")


def gen_refpath():
    """obca_py/util.get_init_ref_path of the REFERENCE on the paths of astar_golden.npz (+ synthetic paths with
    several direction changes, repeated poses and 2- / 3-pose pieces)."""
    from . import ref_loader
    U = ref_loader.load_obca_util()

    class Car:
        WHEEL_BASE = 1.9

    g = np.load(os.path.join(GOLD, "astar_golden.npz"))
    po = np.concatenate([[0], np.cumsum(g["path_len"])])
    paths = [g["path"][po[i]:po[i + 1]] for i in range(96) if g["path_len"][i] >= 4]
    rng = np.random.default_rng(7)
    for k in range(24):                                   # synthetic: arcs glued with direction flips
        n_seg = int(rng.integers(1, 5))
        pts, d = [], 1.0
        x, y, yaw = rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-3, 3)
        for sgi in range(n_seg):
            m = int(rng.choice([2, 3, 4, 9, 30]))
            kap = rng.uniform(-0.3, 0.3)
            for j in range(m):
                pts.append([x, y, yaw, kap, d])
                if rng.random() < 0.1:
                    pts.append([x, y, yaw, kap, d])        # repeated pose (singular point)
                x += d * 0.2 * math.cos(yaw); y += d * 0.2 * math.sin(yaw); yaw += d * 0.2 * kap
            d = -d
        paths.append(np.array(pts))
    keep, trajs = [], []
    for p in paths:
        try:
            t = U.get_init_ref_path(Car(), p[:, 0], p[:, 1], p[:, 2], p[:, 3], p[:, 4])
        except ValueError:
            t = np.zeros((0, 5))                           # a piece with fewer than 2 distinct poses: the reference raises
        keep.append(p); trajs.append(np.asarray(t, dtype=np.float64).reshape(-1, 5))
    np.savez_compressed(os.path.join(GOLD, "refpath_golden.npz"), path_len=np.array([len(p) for p in keep]),
                        path=np.concatenate(keep), traj_len=np.array([len(t) for t in trajs]), traj=np.concatenate(trajs))
    print("refpath_golden.npz:", len(keep), "paths;", int(sum(len(t) == 0 for t in trajs)), "raise in the reference")


def dubins_ref_cases(n=2000, seed=77):
    """Seeded (q0, q1, rho) cases of the Dubins cross-check (shared by the generator and the live test)."""
    rng = np.random.default_rng(seed)
    q = np.empty((n, 7))
    q[:, [0, 1, 3, 4]] = rng.uniform(-10, 10, (n, 4))
    q[:, [2, 5]] = rng.uniform(-math.pi, math.pi, (n, 2))
    q[:, 6] = rng.uniform(1.0, 4.0, n)
    # the shapes the planner asks for: opposite headings one row spacing apart (fish-tail / Omega turns)
    q[: n // 10, 2] = math.pi / 2
    q[: n // 10, 5] = -math.pi / 2
    q[: n // 10, 4] = q[: n // 10, 1]
    return q


def dubins_ref_outputs(mod, cases):
    """Word and length (metres) the reference's own pure-Python Dubins planner (path_planner/utils/dubins_path.py)
    gives: ``planning_from_origin`` tries LSL, RSR, LSR, RSL, RLR, LRL and keeps the first minimum."""
    names = ("LSL", "LSR", "RSL", "RSR", "RLR", "LRL")             # pydubins / dubins.c word numbering
    word = np.empty(len(cases), dtype=np.int32)
    length = np.empty(len(cases))
    for k, c in enumerate(cases):
        r = mod.calc_dubins_path(c[0], c[1], c[2], c[3], c[4], c[5], 1.0 / c[6], step_size=0.5)[0]
        word[k] = names.index("".join(r.mode))
        length[k] = r.L * c[6]
    return word, length


def gen_dubins_ref():
    """tests/golden/dubins_ref_golden.npz: the reference repository carries a second, pure-Python Dubins planner
    (path_planner/utils/dubins_path.py) next to the un-vendored pydubins its planner imports; it runs here
    unmodified (matplotlib and its sibling ``draw`` stubbed: plotting only).  Pins the WORD and LENGTH of
    oracle/dubins_port.py and of the K9 kernels on code of the reference itself."""
    import types
    from . import ref_loader
    sys.modules.setdefault("draw", types.ModuleType("draw"))
    mod = ref_loader.load("dubins_path")
    cases = dubins_ref_cases()
    word, length = dubins_ref_outputs(mod, cases)
    np.savez_compressed(os.path.join(GOLD, "dubins_ref_golden.npz"), cases=cases, word=word, length=length)
    print("dubins_ref:", len(cases), "cases, words", np.bincount(word, minlength=6).tolist())


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    args = sys.argv[1:]
    if "rs" in args:
        gen_rs()
    if "df" in args:
        gen_df()
    if "astar" in args:
        gen_astar(int(args[args.index("astar") + 1]))
    if "pawn" in args:
        gen_pawn(int(args[args.index("pawn") + 1]))
    if "sampling" in args:
        gen_sampling()
    if "astar_full" in args:
        gen_astar_full(int(args[args.index("astar_full") + 1]))
    if "astar_ref" in args:
        gen_astar_ref(int(args[args.index("astar_ref") + 1]))
    if "refpath" in args:
        gen_refpath()
    if "oge" in args:
        gen_oge()
    if "offset" in args:
        gen_offset(int(args[args.index("offset") + 1]))
    if "dubins_ref" in args:
        gen_dubins_ref()
    if "ypark" in args:
        gen_ypark(int(args[args.index("ypark") + 1]))
