"""Mirror of ``path_planner/hybrid_a_star_search.py``: same class, constructor,
class constants and ``hybrid_a_star_search(plt=None, max_nodes=2000) ->
(x, y, yaw, dirs, ks, counter)`` contract; the search itself runs in
``hl_hybrid_astar_batch`` (one CTA per scenario).  A single search is a batch of one;
sweeps build one ``EnvBatch`` for all scenarios and call ``ops.hybrid_astar_batch``."""
import math

import numpy as np

from . import _lib, ops
from .env_batch import EnvBatch, make_record


def make_search_params(car_model, motion_type="King", yaw_resolution=math.radians(10), plan_resolution=0.1,
                       max_nodes=2000, max_path_poses=16384, costs=None):
    """HlSearchParams with the primitive table of ``_get_motion_steers_reeds_shepp`` ("King",
    hybrid_a_star_search.py:343-354) or ``_get_motion_steers_dubins`` ("Pawn", :331-341) and the per-primitive constants
    of ``kinematic_simulation_node`` / ``simulated_path_cost`` (:370-375, :322, :402), all evaluated on the host with the
    reference's own numpy / math calls.  "Pawn" selects the Dubins goal extension (:184-230) in the kernel; the Dubins
    solver restates the un-vendored pydubins (parity unpinned)."""
    if motion_type not in ("King", "Pawn"):
        raise ValueError(f"unknown motion_type {motion_type!r}")
    c = dict(STEER_COST=1, DELTA_STEER_COST=5, DIRECTION_CHANGE_COST=1000, REVERSE_COST=5000, HYBRID_COST=50,
             MIN_LENGTH_TO_GOAL=1000)
    c.update(costs or {})
    if motion_type == "King":
        steers = np.arange(car_model.MAX_STEER, -(car_model.MAX_STEER + yaw_resolution / 2.0), -yaw_resolution / 2.0)
        dirs = np.ones_like(steers)
        dirs[1:len(dirs):2] = -1
    else:
        steers = np.arange(car_model.MAX_STEER, -(car_model.MAX_STEER + yaw_resolution), -yaw_resolution)
        dirs = np.ones_like(steers)
    if len(steers) > _lib.HL_MAX_PRIMS:
        raise _lib.HeadlandError(f"{len(steers)} primitives exceed HL_MAX_PRIMS={_lib.HL_MAX_PRIMS}")
    p = _lib.HlSearchParams()
    p.plan_resolution = plan_resolution
    p.yaw_resolution = yaw_resolution
    p.maxc = car_model.curvature
    p.max_steer = car_model.MAX_STEER
    p.wheel_base = car_model.WHEEL_BASE
    p.n_prims = len(steers)
    for i, (st, d) in enumerate(zip(steers, dirs)):
        p.prim_steer[i] = st
        p.prim_dir[i] = d
        p.prim_yaw_step[i] = d * plan_resolution / car_model.WHEEL_BASE * math.tan(st)
        curv = np.tan(st) / car_model.WHEEL_BASE
        p.prim_curv[i] = curv
        p.prim_steer_eff[i] = math.atan(curv * car_model.WHEEL_BASE)
    p.steps_default = round(1.5 / plan_resolution)
    p.steps_large = round(1.0 / plan_resolution)
    p.steer_cost = c["STEER_COST"]
    p.delta_steer_cost = c["DELTA_STEER_COST"]
    p.direction_change_cost = c["DIRECTION_CHANGE_COST"]
    p.reverse_cost = c["REVERSE_COST"]
    p.hybrid_cost = c["HYBRID_COST"]
    p.min_length_to_goal = c["MIN_LENGTH_TO_GOAL"]
    p.max_nodes = int(max_nodes)
    p.max_path_poses = int(max_path_poses)
    p.motion_type = 0 if motion_type == "King" else 1
    p.dubins_capacity = 0
    return p, np.vstack((steers, dirs)).T


def scenario_array(env_ids, starts, goals):
    n = len(env_ids)
    a = np.zeros(n, dtype=_lib.SCENARIO_DTYPE)
    a["env_id"] = env_ids
    a["start"] = np.asarray(starts, dtype=np.float64).reshape(n, 3)
    a["goal"] = np.asarray(goals, dtype=np.float64).reshape(n, 3)
    return a


def unpack_path(out, i):
    """(x, y, yaw, dirs, ks) lists of scenario ``i`` from a host result dict."""
    r = out["results"][i]
    a, b = int(r["path_offset"]), int(r["path_offset"]) + int(r["path_len"])
    return (out["x"][a:b].tolist(), out["y"][a:b].tolist(), out["yaw"][a:b].tolist(),
            [int(d) for d in out["dir"][a:b]], out["k"][a:b].tolist())


class Node:
    def __init__(self, grid_index, traj, curvature, cost, direction, parent_index):
        self.grid_index = grid_index
        self.traj = traj
        self.curvature = curvature
        self.cost = cost
        self.parent_index = parent_index
        self.direction = direction

    def get_hybrid_index(self):
        return tuple([self.grid_index[0], self.grid_index[1], self.grid_index[2]])


class HybridAStarSearch(object):
    STEER_COST = 1
    DELTA_STEER_COST = 5
    DEVIATION_COST = 1
    DISTANCE_COST = 1
    DIRECTION_CHANGE_COST = 1000
    REVERSE_COST = 5000
    HYBRID_COST = 50
    MIN_LENGTH_TO_GOAL = 1000

    def __init__(self, start_pose, goal_pose, config_environment, car_model, search_heuristic,
                 motion_type="Pawn", yaw_resolution=math.radians(10), plan_resolution=0.1):
        self.plan_resolution = plan_resolution
        self.yaw_resolution = yaw_resolution
        self.config_env = config_environment
        self.car_model = car_model
        self.search_heuristic = search_heuristic
        self.motion_type = motion_type
        self.start_pose = [float(start_pose[0]), float(start_pose[1]), float(start_pose[2])]
        self.goal_pose = [float(goal_pose[0]), float(goal_pose[1]), float(goal_pose[2])]
        _, self.motion_steers = make_search_params(car_model, motion_type, yaw_resolution, plan_resolution)
        self.start_node = self.init_node(self.start_pose)
        self.goal_node = self.init_node(self.goal_pose)
        self.last_result = None
        self.expanded = []

    def calculate_node_index(self, x, y, yaw):
        return (round(x / self.plan_resolution), round(y / self.plan_resolution), round(yaw / self.yaw_resolution))

    def init_node(self, pose):
        idx = self.calculate_node_index(pose[0], pose[1], pose[2])
        return Node(idx, [[pose[0], pose[1], pose[2]]], [0], 0, [1], idx)

    def _costs(self):
        return {k: getattr(self, k) for k in ("STEER_COST", "DELTA_STEER_COST", "DIRECTION_CHANGE_COST",
                                               "REVERSE_COST", "HYBRID_COST", "MIN_LENGTH_TO_GOAL")}

    def hybrid_a_star_search(self, plt=None, max_nodes=2000):
        params, _ = make_search_params(self.car_model, self.motion_type, self.yaw_resolution,
                                       self.plan_resolution, max_nodes, costs=self._costs())
        envs = EnvBatch([make_record(self.config_env, self.car_model, self.search_heuristic)])
        try:
            out = ops.hybrid_astar_batch(envs, scenario_array([0], [self.start_pose], [self.goal_pose]), params,
                                         path_capacity=params.max_path_poses)
        finally:
            envs.close()
        r = out["results"][0]
        self.last_result = r
        self.status = _lib.STATUS_NAMES[int(r["status"])]
        self.expanded = [tuple(int(v) for v in k) for k in ops.expanded_of(out, 0)]
        if r["status"] in (4, 5):
            raise _lib.HeadlandError(f"hybrid_a_star_search: device status {self.status}")
        if r["status"] == 1:
            print("start or goal position is interfere with obstacles!!")
            return [], [], [], [], [], 0
        if r["status"] == 3:
            print("drop the planner")
        if r["status"] == 2:
            print("No solution is available")
        print("counter of nodes: ", int(r["counter"]))
        x, y, yaw, dirs, ks = unpack_path(out, 0)
        if plt is not None and len(x):
            plt.plot(x, y, linewidth=0.3, color="g")
        return (x, y, yaw, dirs, ks, int(r["counter"]))
