"""Synthetic headland scenarios of BASELINE.json config 5: varied
row spacing, headland width and vehicle length, one Hybrid A* warm-start problem each.

A scenario is PLAIN DATA (dict of numpy arrays / floats) so that the CUDA path and the
CPU oracle are fed identical inputs.  The goal pose is the first pose of a Y-type
parking manoeuvre into the target row (the pose the reference's planner searches to,
``headland_path_planning.py:191-219``) picked from a short candidate list by a
feasibility callback -- the GPU collision kernel in ``bench.py`` / the sweep, the oracle
in CPU tests; both give the same booleans, so the same scenarios.
"""
import math

import numpy as np

from .utils import map_utils
from .utils.path_utils import angle_wrap

ROW_NUM = 8
ROW_LENGTH = 20.0
TREE_WIDTH = 0.3
DRIVE_ROW_OFFSET = 4.5          # headland_planner_y_type_park default, headland_path_planning.py:130
# (backward length, forward length, backward steer, forward steer): a thinned version of the
# 4-deep sweep of search_y_type_parking_path (headland_path_planning.py:382-451)
YPARK_CANDIDATES = [(bl, fl, bs, 0.5) for bl in (6.0, 5.0, 4.0, 3.0) for fl in (2.8, 2.0) for bs in (0.0, 0.15)]


def _motion_path(init_pose, steer, direction, length, wheel_base, step):
    """headland_path_planning.calculate_motion_path (:455-485)."""
    num_steps = round(length / step)
    yaw_step = direction * step / wheel_base * math.tan(steer)
    init_yaw = angle_wrap(init_pose[-1] + yaw_step)
    yaws = angle_wrap(np.linspace(init_yaw, init_yaw + yaw_step * num_steps, num_steps + 1))
    xs = init_pose[0] + np.cumsum(step * np.cos(yaws[:-1]) * direction)
    ys = init_pose[1] + np.cumsum(step * np.sin(yaws[:-1]) * direction)
    return np.vstack([np.asarray(init_pose, dtype=np.float64)[None, :3], np.vstack([xs, ys, yaws[1:]]).T])


def y_park_path(end_pose, backward_length, backward_steer, forward_length, forward_steer, wheel_base, step):
    """Poses [P,3] of the Y-type parking path in the odom frame, planned backwards from the
    row-enter pose (get_y_type_parking_path + get_path_in_odom, :488-528)."""
    back = _motion_path([0.0, 0.0, 0.0], backward_steer, -1, backward_length, wheel_base, step)
    fwd = _motion_path(back[-1], forward_steer, 1, forward_length, wheel_base, step)
    local = np.vstack([fwd[::-1], back[::-1]])
    c, s = math.cos(end_pose[2]), math.sin(end_pose[2])
    out = np.empty_like(local)
    out[:, 0] = c * local[:, 0] - s * local[:, 1] + end_pose[0]
    out[:, 1] = s * local[:, 0] + c * local[:, 1] + end_pose[1]
    out[:, 2] = local[:, 2] + end_pose[2]
    return out


def scenario_spec(i, step_size=0.2):
    """Deterministic plain-data description of scenario ``i`` (before the goal is chosen)."""
    rng = np.random.default_rng(1234 + i)
    row_width = rng.uniform(2.2, 4.0)
    slope = math.radians(rng.uniform(0.0, 15.0))
    l_std = 0.5 if rng.integers(0, 2) else 0.0
    headland_width = rng.uniform(5.0, 9.0)
    axle_to_front = rng.uniform(2.85, 4.5)
    r_start = int(rng.integers(0, 5))
    r_end = r_start + int(rng.integers(1, 3))
    side = map_utils.NEAR_SIDE if rng.integers(0, 2) else map_utils.FAR_SIDE
    np.random.seed(1234 + i)
    rows = map_utils.create_tree_rows(ROW_NUM, row_width, ROW_LENGTH, slope_angle=slope, l_std=l_std)
    start = map_utils.get_base_pose(r_start, rows, 0.0, side=side, pose_type=map_utils.LEAVE_POSE)
    end = map_utils.get_base_pose(r_end, rows, 0.0, side=side, pose_type=map_utils.ENTER_POSE)
    car = dict(max_steer=0.55, wheel_base=1.9, axle_to_front=axle_to_front, axle_to_back=0.55, width=1.48)
    # steer direction of the backward leg (get_backward_steer_dir_for_y_type_parking, :367-376)
    bdir = np.sign(math.cos(start[2])) if end[1] - start[1] > 0 else np.sign(-math.cos(start[2]))
    cands = [y_park_path(end, bl, bs * bdir, fl, -fs * bdir, car["wheel_base"], step_size)
             for (bl, fl, bs, fs) in YPARK_CANDIDATES]
    return dict(index=i, rows=rows, tree_width=TREE_WIDTH, headland_width=headland_width, car=car, side=side,
                start=start, end=end, ypark_candidates=cands, step_size=step_size, seed=1234 + i)


def finalize(spec, candidate_feasible):
    """Choose the goal (first feasible Y-park candidate's first pose, else the row-enter pose)
    and derive the guide waypoints (get_topology_waypoints, orchard_geometry_environment.py:199-248)."""
    from .orchard_geometry_environment import OrchardGeometryEnvironment
    goal = spec["end"].copy()
    pick = -1
    for k, ok in enumerate(candidate_feasible):
        if ok:
            goal = spec["ypark_candidates"][k][0].copy()
            pick = k
            break
    goal[2] = float(angle_wrap(goal[2]))
    env = OrchardGeometryEnvironment(spec["rows"], [], tree_width=spec["tree_width"],
                                     headland_width=spec["headland_width"])
    np.random.seed(spec["seed"])          # check_side_of_a_point draws jitter (Appendix A, 9)
    way = env.get_topology_waypoints(spec["start"], goal, drive_row_offset=DRIVE_ROW_OFFSET)
    out = dict(spec)
    out.update(goal=goal, ypark_pick=pick, waypoints=way)
    return out


def build_host_objects(scn):
    """(env, car, heuristic) mirror objects of a finalized scenario (host geometry only)."""
    from .car_model import CarModel
    from .orchard_geometry_environment import OrchardGeometryEnvironment
    from .reference_line_heuristic import ReferenceLineHeuristic
    car = CarModel(**scn["car"])
    env = OrchardGeometryEnvironment(scn["rows"], [], tree_width=scn["tree_width"], headland_width=scn["headland_width"])
    heur = ReferenceLineHeuristic(scn["waypoints"], scn["goal"], car)
    return env, car, heur


def gpu_candidate_feasibility(specs):
    """Feasibility of every Y-park candidate of every spec in ONE collision launch
    (check_path_feasibility with boundary_check=True, body only)."""
    import torch
    from . import ops
    from .car_model import CarModel
    from .env_batch import EnvBatch, make_record
    from .orchard_geometry_environment import OrchardGeometryEnvironment
    recs, poses, env_id, starts = [], [], [], [0]
    for e, sp in enumerate(specs):
        env = OrchardGeometryEnvironment(sp["rows"], [], tree_width=sp["tree_width"], headland_width=sp["headland_width"])
        recs.append(make_record(env, CarModel(**sp["car"])))
        for path in sp["ypark_candidates"]:
            poses.append(path)
            env_id.append(np.full(len(path), e, dtype=np.int32))
            starts.append(starts[-1] + len(path))
    envs = EnvBatch(recs)
    dev = ops._device()
    bad = ops.collision_check(envs, np.concatenate(poses), env_id=np.concatenate(env_id),
                              flags=ops.CHECK_OBSTACLES | ops.CHECK_BOUNDARY)
    path_bad = ops.path_reduce(envs, bad, torch.from_numpy(np.asarray(starts, dtype=np.int64)).to(dev)).cpu().numpy()
    envs.close()
    ncand = len(YPARK_CANDIDATES)
    return [[not bool(b) for b in path_bad[e * ncand:(e + 1) * ncand]] for e in range(len(specs))]


def make_scenarios_gpu(indices, step_size=0.2):
    specs = [scenario_spec(i, step_size) for i in indices]
    feas = gpu_candidate_feasibility(specs)
    return [finalize(sp, f) for sp, f in zip(specs, feas)]
