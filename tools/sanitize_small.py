"""Small K1 / K2 / K4 workload for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Uses the committed golden scenarios (tests/golden/astar_golden.npz), so no oracle run is needed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep
from headland_trajectory_planning_b200.env_batch import EnvBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "astar_golden.npz"))
specs = [SC.scenario_spec(int(i)) for i in g["index"][:n]]
scns = [SC.finalize(sp, f) for sp, f in zip(specs, g["feas"][:n])]
recs, scen, car = sweep.build_records(scns)
envs = EnvBatch(recs)
params = sweep.search_params(car, max_nodes=int(sys.argv[2]) if len(sys.argv) > 2 else 40)
out = ops.hybrid_astar_batch(envs, scen, params, path_capacity=2048 * n)
print("K4 status histogram", np.unique(out["results"]["status"], return_counts=True), "pops", int(out["results"]["counter"].sum()))
rng = np.random.default_rng(0)
poses = np.stack([rng.uniform(-10, 4, 20000), rng.uniform(-2, 20, 20000), rng.uniform(-3.2, 3.2, 20000)], axis=1)
poses[::500, 0] = np.nan
bad = ops.collision_check(envs, poses, env_id=np.zeros(20000, np.int32), flags=ops.CHECK_OBSTACLES | ops.CHECK_BOUNDARY | ops.CHECK_LANE)
print("K1 infeasible", float(bad.float().mean().item()))
sg = np.concatenate([poses[:256], poses[256:512]], axis=1)
sg[np.isnan(sg)] = 0.0
words, count, _ = ops.rs_all_paths(sg, car.curvature, 0.1, envs=envs, flags=ops.CHECK_OBSTACLES, want_order=True)
torch.cuda.synchronize()
print("K2 words", int(count.sum().item()))
