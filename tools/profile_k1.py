import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from headland_trajectory_planning_b200 import ops
class A:
    collision_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
    collision_mode = sys.argv[2] if len(sys.argv) > 2 else 'random'
torch.cuda.set_device(0)
print(bench.collision_microbench(A, torch.device("cuda", 0), ops.measure_fp32_peak(0)))
