"""Times k_rs_all_paths on BASELINE config 3 (random pose pairs, every word sampled at 0.1 m and collision-checked
against the canonical orchard): python tools/profile_rs.py [n_pairs = 1 Mi]; REPS fields, CUDA events."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from headland_trajectory_planning_b200 import ops
from headland_trajectory_planning_b200.car_model import CarModel
from headland_trajectory_planning_b200.env_batch import EnvBatch, make_record
from headland_trajectory_planning_b200.orchard_geometry_environment import OrchardGeometryEnvironment
from headland_trajectory_planning_b200.utils import map_utils
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rng = np.random.default_rng(0)
sg = np.empty((n, 6))
sg[:, [0, 1, 3, 4]] = rng.uniform(-10, 10, (n, 4))
sg[:, [2, 5]] = rng.uniform(-math.pi, math.pi, (n, 2))
np.random.seed(1)
rows = map_utils.create_tree_rows(8, 2.5, 20, slope_angle=math.radians(10), l_std=0.0)
env = OrchardGeometryEnvironment(rows, [], tree_width=0.3, headland_width=6.0)
car = CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48)
envs = EnvBatch([make_record(env, car)])
d_sg = torch.from_numpy(sg).cuda()
words, count, _ = ops.rs_all_paths(d_sg, car.curvature, 0.1, envs=envs, flags=ops.CHECK_OBSTACLES, want_order=False)
del words
torch.cuda.synchronize()
ts = []
for _ in range(int(os.environ.get("REPS", "4"))):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    words, count, _ = ops.rs_all_paths(d_sg, car.curvature, 0.1, envs=envs, flags=ops.CHECK_OBSTACLES, want_order=False)
    b.record(); torch.cuda.synchronize()
    ts.append(round(a.elapsed_time(b), 2))
    chk = (int(count.sum().item()), int(words.view(torch.uint8)[: 1 << 24].to(torch.int64).sum().item()))
    del words
print("ms", ts, "min", min(ts), "M pairs/s", round(n / min(ts) / 1e3, 2), "words", chk)
