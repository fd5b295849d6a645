import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep
from headland_trajectory_planning_b200.env_batch import EnvBatch
g = np.load("tests/golden/astar_golden.npz")
n = len(g["index"])
specs = [SC.scenario_spec(int(i)) for i in g["index"]]
feas = SC.gpu_candidate_feasibility(specs)
scns = [SC.finalize(sp, f) for sp, f in zip(specs, feas)]
recs, scen, car = sweep.build_records(scns)
out = ops.hybrid_astar_batch(EnvBatch(recs), scen, sweep.search_params(car), path_capacity=2048 * n)
res = out["results"]
want = g["rs_poses"] + g["primitive_poses"]
bad = np.nonzero(res["n_pose_checks_ref"] != want)[0]
print("mismatch", len(bad), "of", n)
for i in bad[:12]:
    print(i, "status", res["status"][i], "arrival", res["arrival"][i], "counter", res["counter"][i], "got", res["n_pose_checks_ref"][i],
          "want", want[i], "rs", g["rs_poses"][i], "prim", g["primitive_poses"][i])
