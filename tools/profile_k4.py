import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep
from headland_trajectory_planning_b200.env_batch import EnvBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
import pickle
cache = f"/tmp/hl_k4_scen_{n}.pkl"            # scenario construction takes ~10 s: reuse it across variant builds
if os.path.exists(cache):
    recs, scen, car = pickle.load(open(cache, "rb"))
else:
    scns = SC.make_scenarios_gpu(list(range(n)))
    recs, scen, car = sweep.build_records(scns)
    try:
        pickle.dump((recs, scen, car), open(cache, "wb"))
    except Exception:
        pass
envs = EnvBatch(recs)
params = sweep.search_params(car)
d_scen = torch.from_numpy(scen.view(np.uint8).reshape(-1)).cuda()
times = []
for rep in range(int(os.environ.get('REPS', '6'))):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    o = ops.hybrid_astar_batch(envs, d_scen, params, path_capacity=1024 * n, to_host=False)
    b.record(); torch.cuda.synchronize()
    times.append(round(a.elapsed_time(b), 1))
print("ms", times, "min", min(times), "median", sorted(times)[len(times) // 2])
ph = ops.astar_phase_cycles()
tot = sum(ph.values()) or 1
print({k: round(100 * v / tot, 1) for k, v in ph.items()})
