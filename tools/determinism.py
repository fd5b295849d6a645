import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep
from headland_trajectory_planning_b200.env_batch import EnvBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scns = SC.make_scenarios_gpu(list(range(n)))
recs, scen, car = sweep.build_records(scns)
envs = EnvBatch(recs)
params = sweep.search_params(car)
outs = []
for rep in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    o = ops.hybrid_astar_batch(envs, scen, params, path_capacity=1024 * n)
    b.record(); torch.cuda.synchronize()
    r = o["results"]
    print("rep", rep, "ms", round(a.elapsed_time(b), 1), "expansions", int(r["n_expanded"].sum()), "checks", int(r["n_pose_checks"].sum()),
          "exact", int(r["n_exact"].sum()), "status", np.unique(r["status"], return_counts=True)[1].tolist(),
          "max Mcycles", r["cycles"].max() / 1e6, "argmax", int(r["cycles"].argmax()), "counter there", int(r["counter"][r["cycles"].argmax()]))
    outs.append(o)
for rep in (1, 2):
    d = np.nonzero((outs[0]["results"]["counter"] != outs[rep]["results"]["counter"]) | (outs[0]["results"]["status"] != outs[rep]["results"]["status"]))[0]
    print("rep", rep, "differs from rep 0 in", len(d), "scenarios", d[:10].tolist())
    for i in d[:3]:
        print("   ", i, outs[0]["results"][i], outs[rep]["results"][i])
top = np.argsort(-outs[0]["results"]["cycles"])[:8]
print("slowest:", [(int(i), int(outs[0]["results"]["counter"][i]), round(outs[0]["results"]["cycles"][i] / 1e6, 1), int(outs[0]["results"]["n_pose_checks"][i]), int(outs[0]["results"]["n_exact"][i])) for i in top])
