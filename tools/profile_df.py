"""Times hl_distance_field on BASELINE config 4 (4096 x 4096, King): REPS fields with CUDA events (the call includes
its host looks between launch batches, so a busy host shows up here), min / median.
Under `ncu --metrics gpu__time_duration.sum -k regex:k_df_` the summed kernel time is the host-independent figure
(tools/variants_prebuilt.sh with NCU_SUM=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from headland_trajectory_planning_b200 import ops
from headland_trajectory_planning_b200.utils.occupancy_grid_utils import synthetic_grid
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(os.environ.get("REPS", "8"))
occ, goal = synthetic_grid(n, seed=1)
d_occ = torch.from_numpy(occ.astype(np.uint8)).cuda()
ops.distance_field(d_occ, goal, "King")
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out, sweeps = ops.distance_field(d_occ, goal, "King")
    b.record(); torch.cuda.synchronize()
    ts.append(round(a.elapsed_time(b), 2))
print("ms", ts, "min", min(ts), "median", sorted(ts)[len(ts) // 2], "launches", int(sweeps),
      "checksum", float(out[torch.isfinite(out)].sum().item()))
