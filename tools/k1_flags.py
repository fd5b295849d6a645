import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from headland_trajectory_planning_b200 import ops
from headland_trajectory_planning_b200.car_model import CarModel
from headland_trajectory_planning_b200.env_batch import EnvBatch, make_record
from headland_trajectory_planning_b200.orchard_geometry_environment import OrchardGeometryEnvironment
from headland_trajectory_planning_b200.reference_line_heuristic import ReferenceLineHeuristic
from headland_trajectory_planning_b200.utils import map_utils
dev = torch.device("cuda", 0)
np.random.seed(1)
rows = map_utils.create_tree_rows(8, 2.5, 20, slope_angle=math.radians(10), l_std=0.0)
env = OrchardGeometryEnvironment(rows, [], tree_width=0.3, headland_width=6.0)
car = CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48)
ends = rows[:, 0, :]
way = np.vstack(([0.88, 3.75], [[ends[i, 0] - 4.5, ends[i, 1]] for i in (2, 3, 4)], [-1.2, 11.25]))
heur = ReferenceLineHeuristic(way, [-1.2, 11.25, 0.0], car)
envs = EnvBatch([make_record(env, car, heur)])
n = 1 << 24
g = torch.Generator(device=dev).manual_seed(0)
poses = torch.empty((n, 3), dtype=torch.float64, device=dev)
poses[:, 0] = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 14.0 - 10.0
poses[:, 1] = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 22.0 - 2.0
poses[:, 2] = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 2.0 - 1.0) * math.pi
for name, flags in [("obs", 1), ("obs+bnd", 3), ("lane", 8), ("all", 11), ("bnd", 2)]:
    for _ in range(2):
        out, nex = ops.collision_check(envs, poses, flags=flags, count_exact=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = ops.collision_check(envs, poses, flags=flags); b.record(); torch.cuda.synchronize()
    print(f"{name:8s} {a.elapsed_time(b):7.3f} ms  exact_frac {nex.item()/n:.2e}  bad_frac {out.float().mean().item():.3f}")
