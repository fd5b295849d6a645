"""K1 micro-benchmark, both pose orders (bench.collision_microbench): one line each with checks/s and the roofline fraction."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from headland_trajectory_planning_b200 import ops
torch.cuda.set_device(0)
peak = ops.measure_fp32_peak(0)
for n, mode in ((1 << 24, "random"), (1 << 23, "paths")):
    class A:
        collision_poses = n
        collision_mode = mode
    r = bench.collision_microbench(A, torch.device("cuda", 0), peak)
    print(mode, round(r["value"] / 1e9, 3), "G checks/s", round(r["ms_per_launch"], 4), "ms  frac", round(r["roofline"]["frac"], 4), "infeasible", r["infeasible_frac"])
