import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep
from headland_trajectory_planning_b200.env_batch import EnvBatch, pack_structs
n = 4096
scns = SC.make_scenarios_gpu(list(range(n)))
recs, scen, car = sweep.build_records(scns)
params = sweep.search_params(car)
structs = pack_structs(recs)
host_scen = torch.from_numpy(scen.view(np.uint8).reshape(-1).copy()).pin_memory()
dev = torch.device("cuda", 0)
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    envs = EnvBatch(recs, structs=structs); torch.cuda.synchronize(); t1 = time.perf_counter()
    d_s = host_scen.to(dev, non_blocking=True)
    o = ops.hybrid_astar_batch(envs, d_s, params, path_capacity=1024 * n, to_host=False); torch.cuda.synchronize(); t2 = time.perf_counter()
    used = int(o["cursor"].item()); kused = int(o["kcursor"].item())
    r = o["results"].cpu(); e = o["expanded"][:kused].cpu(); xs = [o[k][:used].cpu() for k in ("x", "y", "yaw", "k", "dir")]
    torch.cuda.synchronize(); t3 = time.perf_counter()
    envs.close(); t4 = time.perf_counter()
    print(f"upload {1e3*(t1-t0):.1f} ms  search {1e3*(t2-t1):.1f} ms  d2h {1e3*(t3-t2):.1f} ms  free {1e3*(t4-t3):.1f} ms  total {1e3*(t4-t0):.1f}")
