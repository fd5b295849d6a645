"""What slows a long (401-pop) search down when it shares its SM with four others: the role alignment waiting for the
slowest of five DIFFERENT iterations, or contention for the SM?  Times k copies of the SAME long scenario (identical
iterations in every slot: nothing to wait for) against k different long scenarios."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pickle
from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep, _lib
from headland_trajectory_planning_b200.env_batch import EnvBatch
n = 4096
cache = f"/tmp/hl_k4_scen_{n}.pkl"
if os.path.exists(cache):
    recs, scen, car = pickle.load(open(cache, "rb"))
else:
    recs, scen, car = sweep.build_records(SC.make_scenarios_gpu(list(range(n))))
    pickle.dump((recs, scen, car), open(cache, "wb"))
envs = EnvBatch(recs)
params = sweep.search_params(car)

def run(sc, reps=3):
    d = torch.from_numpy(sc.view(np.uint8).reshape(-1)).cuda()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        o = ops.hybrid_astar_batch(envs, d, params, path_capacity=1024 * len(sc), to_host=False)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    r = o["results"].cpu().numpy().view(_lib.RESULT_DTYPE)
    return min(ts), r["cycles"].mean() / 1.965e6

t, _ = run(scen)
d = torch.from_numpy(scen.view(np.uint8).reshape(-1)).cuda()
res = ops.hybrid_astar_batch(envs, d, params, path_capacity=1024 * n, to_host=False)["results"].cpu().numpy().view(_lib.RESULT_DTYPE)
idx = np.nonzero(res["status"] == 3)[0]
print("full sweep", round(t, 2), "ms;", len(idx), "long scenarios")
for k in (1, 5, 148, 740):
    row = []
    for base in idx[:3]:
        tt, per = run(np.repeat(scen[base:base + 1], k).copy())
        row.append(f"{tt:6.2f} (mean {per:5.2f})")
    print(f"{k:4d} copies of ONE long scenario [3 different ones]: " + "  ".join(row))
for k in (5, 148, 546):
    tt, per = run(scen[idx[:k]].copy())
    print(f"{k:4d} DIFFERENT long scenarios: {tt:6.2f} ms (mean per scenario {per:5.2f})")
tt, per = run(np.concatenate([scen[idx], scen[idx[:194]]]).copy())
print(f" 740 long scenarios (546 different + 194 repeats): {tt:6.2f} ms (mean per scenario {per:5.2f})")

# where does the time of a long search go, alone and with four identical neighbours on its SM?  (phase timers, lane-0
# cycles of the lead expander and of the shooter, in microseconds per pop at 1.965 GHz)
def phases(sc):
    d = torch.from_numpy(sc.view(np.uint8).reshape(-1)).cuda()
    ops.astar_phase_cycles()
    ops.hybrid_astar_batch(envs, d, params, path_capacity=1024 * len(sc), to_host=False)
    torch.cuda.synchronize()
    ph = ops.astar_phase_cycles()
    return {k: v / len(sc) / 401 / 1965.0 for k, v in ph.items()}
if os.environ.get("PHASES", "1") == "1":
    base = idx[1]
    a = phases(scen[base:base + 1].copy()); b = phases(np.repeat(scen[base:base + 1], 5).copy()); c = phases(scen[idx[:5]].copy())
    print(f"{'phase':16s} {'alone':>8s} {'5 copies':>8s} {'5 different':>11s}   (us per pop)")
    for k in a:
        print(f"{k:16s} {a[k]:8.2f} {b[k]:8.2f} {c[k]:11.2f}")
    print(f"{'sum':16s} {sum(a.values()):8.2f} {sum(b.values()):8.2f} {sum(c.values()):11.2f}")
