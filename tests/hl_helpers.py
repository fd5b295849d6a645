"""Shared scenario builders for the tests: the same plain inputs are handed to the
CPU oracle classes and to the CUDA-backed mirror classes."""
import math

import numpy as np

from oracle import planner as OP

MOWER_AUX = [[[-1.84, 0.5], 1.0, 1.1]]                     # test/obca.ipynb cell 5
PRUNER_AUX = [[[3.259, -0.175], 1.325, 0.3]]
SPRAYER_AUX = [[[-2.1, 1.0 / 2], 1.0, 1.22], [[-1.0, 4.3 / 2], 0.4, 0.5], [[-1.0, -4.3 / 2 + 0.4], 0.4, 0.5]]


def canonical_rows(l_std=0.0, seed=1, row_width=2.5, slope_deg=10.0):
    np.random.seed(seed)
    return OP.create_tree_rows(8, row_width, 20, slope_angle=math.radians(slope_deg), l_std=l_std)


def make_pair(rows, obstacles=(), tree_width=0.3, headland_width=6.0, axle_to_front=3.0, aux=None,
              waypoints=None, goal=None):
    """Returns (oracle_env, oracle_car, oracle_heur), (gpu_env, gpu_car, gpu_heur)."""
    from headland_trajectory_planning_b200.car_model import CarModel
    from headland_trajectory_planning_b200.orchard_geometry_environment import OrchardGeometryEnvironment
    from headland_trajectory_planning_b200.reference_line_heuristic import ReferenceLineHeuristic
    obstacles = list(obstacles)
    kw = dict(max_steer=0.55, axle_to_front=axle_to_front, axle_to_back=0.55, width=1.48,
              aux_poly_features=aux or [], with_aux=bool(aux))
    o_env = OP.OrchardGeometryEnvironment(rows, obstacles, tree_width=tree_width, headland_width=headland_width)
    o_car = OP.CarModel(**kw)
    g_env = OrchardGeometryEnvironment(rows, obstacles, tree_width=tree_width, headland_width=headland_width)
    g_car = CarModel(**kw)
    o_h = g_h = None
    if waypoints is not None:
        o_h = OP.ReferenceLineHeuristic(np.asarray(waypoints), goal, o_car)
        g_h = ReferenceLineHeuristic(np.asarray(waypoints), goal, g_car)
    return (o_env, o_car, o_h), (g_env, g_car, g_h)


def random_poses(rng, n, box=(-12.0, 35.0, -6.0, 24.0)):
    x = rng.uniform(box[0], box[1], n)
    y = rng.uniform(box[2], box[3], n)
    yaw = rng.uniform(-math.pi, math.pi, n)
    return np.stack([x, y, yaw], axis=1)


def headland_poses(rng, n, rows, side_x=None):
    """Poses concentrated in the near-side headland where the planner works."""
    x0 = rows[:, 0, 0].min()
    x = rng.uniform(x0 - 8.0, x0 + 3.0, n)
    y = rng.uniform(rows[:, 0, 1].min() - 4.0, rows[:, 0, 1].max() + 4.0, n)
    yaw = rng.uniform(-math.pi, math.pi, n)
    return np.stack([x, y, yaw], axis=1)
