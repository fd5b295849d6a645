import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep
from headland_trajectory_planning_b200.env_batch import EnvBatch
g = np.load("tests/golden/astar_golden.npz")
ids = [int(v) for v in sys.argv[1:]]
specs = [SC.scenario_spec(i) for i in ids]
scns = [SC.finalize(sp, list(g["feas"][i])) for sp, i in zip(specs, ids)]
recs, scen, car = sweep.build_records(scns)
out = ops.hybrid_astar_batch(EnvBatch(recs), scen, sweep.search_params(car), path_capacity=4096 * len(ids))
eo = np.concatenate([[0], np.cumsum(g["n_expanded"])])
for k, i in enumerate(ids):
    r = out["results"][k]
    want = g["expanded"][eo[i]:eo[i + 1]]
    got = ops.expanded_of(out, k)
    n = min(len(want), len(got))
    d = np.nonzero((want[:n] != got[:n]).any(axis=1))[0]
    print(i, "status", r["status"], g["status"][i], "counter", r["counter"], g["counter"][i], "n_exp", len(got), len(want),
          "first diff", (int(d[0]), want[d[0]].tolist(), got[d[0]].tolist()) if len(d) else None,
          "path_len", r["path_len"], g["path_len"][i], "rs_word", r["rs_word"], "exact", r["n_exact"])
    if len(d):
        j = int(d[0])
        # is the oracle's next key present later in ours (pure order swap)?
        later = np.nonzero((got[j:] == want[j]).all(axis=1))[0]
        print("   oracle key appears in ours at +", later[:3].tolist(), " ours appears in oracle at +",
              np.nonzero((want[j:] == got[j]).all(axis=1))[0][:3].tolist())
po = np.concatenate([[0], np.cumsum(g["path_len"])])
from headland_trajectory_planning_b200.hybrid_a_star_search import unpack_path
for k, i in enumerate(ids):
    x, y, yaw, dirs, ks = unpack_path(out, k)
    want = g["path"][po[i]:po[i + 1]]
    got = np.stack([x, y, yaw, ks, np.asarray(dirs, float)], axis=1)
    bad = np.nonzero(~np.isclose(got, want, rtol=1e-5, atol=1e-6).all(axis=1))[0]
    print(i, "bad rows", bad[:10].tolist(), "of", len(want))
    for b in bad[:4]:
        print("   want", want[b].tolist(), "\n   got ", got[b].tolist())
