"""CPU baseline runner: the oracle's Hybrid A* on plain-data scenarios, one scenario per
worker process.  TEST INFRASTRUCTURE / bench ``cpu_baseline`` + ``--impl reference`` only
(see ``oracle/__init__.py``); it stands in for the reference's Python path because
shapely / heapdict cannot be installed (kind = "port")."""
import multiprocessing as mp
import os
import time

import numpy as np

from . import planner as OP


def candidate_feasibility(spec):
    env = OP.OrchardGeometryEnvironment(spec["rows"], [], tree_width=spec["tree_width"],
                                        headland_width=spec["headland_width"])
    car = OP.CarModel(**spec["car"])
    return [env.check_path_feasibility(car, p) for p in spec["ypark_candidates"]]


def run_scenario(scn, max_nodes=400):
    env = OP.OrchardGeometryEnvironment(scn["rows"], [], tree_width=scn["tree_width"],
                                        headland_width=scn["headland_width"])
    car = OP.CarModel(**scn["car"])
    heur = OP.ReferenceLineHeuristic(scn["waypoints"], scn["goal"], car)
    s = OP.HybridAStarSearch(scn["start"], scn["goal"], env, car, heur, motion_type="King",
                             plan_resolution=scn["step_size"])
    t0 = time.perf_counter()
    x, y, yaw, dirs, ks, counter = s.hybrid_a_star_search(max_nodes=max_nodes)
    dt = time.perf_counter() - t0
    return dict(index=scn["index"], status=s.status, counter=counter, expanded=np.array(s.expanded, dtype=np.int32).reshape(-1, 3),
                x=np.array(x), y=np.array(y), yaw=np.array(yaw), dirs=np.array(dirs, dtype=np.int8),
                ks=np.array(ks, dtype=np.float64), seconds=dt, stats=dict(s.stats))


def _quiet_run(scn):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return run_scenario(scn)


def run_pool(scns, cores=None):
    """Search time only (object construction excluded, like the GPU arm excludes host
    geometry construction): returns (results, wall seconds, cores)."""
    cores = cores or os.cpu_count() or 1
    cores = max(1, min(cores, len(scns)))
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    t0 = time.perf_counter()
    if cores == 1:
        res = [_quiet_run(s) for s in scns]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_quiet_run, scns, chunksize=1)
    return res, time.perf_counter() - t0, cores
