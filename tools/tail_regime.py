"""How does the sweep time depend on the number of LONG scenarios in flight?  Times sweeps made only of
scenarios that end with max_nodes (401 pops): 148 (one per SM), 296, 592, 888 (all slots), 546 (the bench mix)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep
from headland_trajectory_planning_b200.env_batch import EnvBatch
from headland_trajectory_planning_b200 import _lib

import pickle
n = 4096
cache = f"/tmp/hl_k4_scen_{n}.pkl"
if os.path.exists(cache):
    recs, scen, car = pickle.load(open(cache, "rb"))
else:
    scns = SC.make_scenarios_gpu(list(range(n)))
    recs, scen, car = sweep.build_records(scns)
    pickle.dump((recs, scen, car), open(cache, "wb"))
envs = EnvBatch(recs)
params = sweep.search_params(car)

def run(sc, reps=4):
    d = torch.from_numpy(sc.view(np.uint8).reshape(-1)).cuda()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        o = ops.hybrid_astar_batch(envs, d, params, path_capacity=1024 * len(sc), to_host=False)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), o

t, o = run(scen, 4)
res = o["results"].cpu().numpy().view(_lib.RESULT_DTYPE)
print("full sweep ms", round(t, 2))
cyc = res["cycles"].astype(np.float64)
cnt = res["counter"]
print("sum cycles / (888 slots) in ms at 1.965 GHz:", round(cyc.sum() / 888 / 1.965e6, 2))
longm = res["status"] == 3
print("long scenarios", int(longm.sum()), "mean ms of a long one", round(cyc[longm].mean() / 1.965e6, 2), "max", round(cyc[longm].max() / 1.965e6, 2))
print("counter histogram", np.histogram(cnt, bins=[0, 2, 5, 10, 20, 50, 100, 200, 400, 402])[0])
ms = cyc / 1.965e6
slots = 740
print("sum of per-scenario ms / 740 slots:", round(ms.sum() / slots, 2), " longest scenario ms:", round(ms.max(), 2))
import heapq
def makespan(order):
    h = [0.0] * slots
    heapq.heapify(h)
    end = 0.0
    for i in order:
        t0 = heapq.heappop(h)
        t1 = t0 + ms[i]
        end = max(end, t1)
        heapq.heappush(h, t1)
    return end
print("list-scheduling makespan with the measured durations: index order", round(makespan(range(n)), 2),
      " longest first", round(makespan(np.argsort(-ms)), 2), " non-trivial (counter > 1) first",
      round(makespan(np.argsort(cnt <= 1, kind="stable")), 2))
# the same sweep with the hard scenarios first in the queue
order = np.argsort(-cnt, kind="stable")
t2, _ = run(scen[order].copy(), 4)
print("sweep with scenarios sorted by (oracle) counter, longest first:", round(t2, 2), "ms")
order = np.argsort(cnt <= 1, kind="stable")
t3, _ = run(scen[order].copy(), 4)
print("sweep with the 1-pop scenarios last:", round(t3, 2), "ms")
idx = np.nonzero(longm)[0]
for k in (1, 37, 148, 296, 546):
    sub = scen[idx[:k]].copy()
    t, o = run(sub, 4)
    r = o["results"].cpu().numpy().view(_lib.RESULT_DTYPE)
    print(f"{k:4d} long scenarios: {t:7.2f} ms   mean per-scenario ms {r['cycles'].mean() / 1.965e6:6.2f}")
short = np.nonzero(~longm)[0]
t, o = run(scen[short].copy(), 4)
print(f"{len(short)} short scenarios only: {t:7.2f} ms")
