"""Mirror of the Y-type parking search of ``path_planner/headland_path_planning.py``.

The reference walks a 4-deep loop over (backward length, forward length, backward steer, forward steer),
builds each candidate path on the CPU and asks ``check_path_feasibility`` for it; the first feasible
candidate wins (``headland_path_planning.py:382-451``).  Here ALL candidates are generated on the GPU by
``hl_ypark_paths`` (K5), checked by ONE ``hl_collision_check`` launch and reduced per candidate by
``hl_path_reduce``; picking the first feasible candidate in loop order gives the reference's answer.
``search_y_type_parking_path_batch`` does the same for many (environment, end pose) problems at once --
the stage that feeds ``hl_hybrid_astar_batch`` its goal poses in a sweep.

Names, arguments, defaults and return conventions follow the reference.
"""
import math

import numpy as np

from . import ops
from .utils.path_utils import angle_wrap


def get_backward_steer_dir_for_y_type_parking(start_pose, end_pose):
    """headland_path_planning.py:360-368."""
    if end_pose[1] - start_pose[1] > 0:
        return np.sign(1 * math.cos(start_pose[2]))
    return np.sign(-1 * math.cos(start_pose[2]))


def calculate_motion_path(init_pose, motion_command, search_length, wheel_base, step):
    """headland_path_planning.py:455-485 (host utility, numpy): rows (x, y, yaw, curvature, direction)."""
    steer_angle, speed_direction = motion_command[0], motion_command[1]
    num_steps = round(search_length / step)
    yaw_step = speed_direction * step / wheel_base * math.tan(steer_angle)
    init_yaw = angle_wrap(init_pose[-1] + yaw_step)
    yaws = angle_wrap(np.linspace(init_yaw, init_yaw + yaw_step * num_steps, num_steps + 1))
    xs = init_pose[0] + np.cumsum(step * np.cos(yaws[:-1]) * speed_direction)
    ys = init_pose[1] + np.cumsum(step * np.sin(yaws[:-1]) * speed_direction)
    path = np.vstack([init_pose, np.vstack([xs, ys, yaws[1:]]).T])
    curvature = math.tan(steer_angle) / wheel_base if abs(steer_angle) > 0.00001 else 0
    return np.hstack((path, np.ones((len(path), 1)) * curvature, np.ones((len(path), 1)) * speed_direction))


def get_y_type_parking_path(car_model, backward_length, backward_steer, forward_length, forward_steer, step):
    """headland_path_planning.py:488-516 (host utility)."""
    back_path = calculate_motion_path([0, 0, 0], [backward_steer, -1], backward_length, car_model.WHEEL_BASE, step)
    forward_path = calculate_motion_path(back_path[-1, :3], [forward_steer, 1], forward_length,
                                         car_model.WHEEL_BASE, step)
    back_path[:, -1] = 1
    back_path = back_path[::-1]
    forward_path[:, -1] = -1
    forward_path = forward_path[::-1]
    return np.vstack([forward_path, back_path])


def y_park_candidates(max_steer_backward=0.4, max_steer_forward=0.45, max_backward_distance=3.5,
                      max_forward_distance=2.0, min_forward_distance=1.4, min_backward_distance=0.7,
                      min_steer_backward=0.3, min_steer_forward=0.3):
    """Rows (backward_length, forward_length, steer_backward, steer_forward) of the reference's 4-deep loop
    (:405-420), in loop order, steers unsigned."""
    steer_backwards = list(np.arange(min_steer_backward, max_steer_backward + 0.1, 0.1))
    if np.max(steer_backwards) < max_steer_backward:
        steer_backwards.append(max_steer_backward)
    steer_forwards = list(np.arange(min_steer_forward, max_steer_forward + 0.1, 0.1))
    if np.max(steer_forwards) < max_steer_forward:
        steer_forwards.append(max_steer_forward)
    bl = np.arange(max_backward_distance, min_backward_distance, -0.1)
    fl = np.arange(max_forward_distance, min_forward_distance, -0.1)
    grid = np.array(np.meshgrid(bl, fl, np.array(steer_backwards), np.array(steer_forwards), indexing="ij"))
    return np.ascontiguousarray(grid.reshape(4, -1).T)


def _device_rows(cands, end_pose, backward_steer_dir, forward_steer_dir, wheel_base):
    rows = np.empty((len(cands), 8), dtype=np.float64)
    rows[:, 0] = cands[:, 0]
    rows[:, 1] = cands[:, 1]
    rows[:, 2] = cands[:, 2] * backward_steer_dir
    rows[:, 3] = cands[:, 3] * forward_steer_dir
    rows[:, 4:7] = np.asarray(end_pose, dtype=np.float64)[:3]
    rows[:, 7] = wheel_base
    return rows


def _with_ks_dirs(poses, row, step):
    """[P,3] poses of one candidate -> the reference's [P,5] rows (x, y, yaw, curvature, direction):
    forward arc first (direction -1), then the backward arc (direction 1) (:506-513)."""
    nf = int(round(row[1] / step))
    k_b = math.tan(row[2]) / row[7] if abs(row[2]) > 0.00001 else 0
    k_f = math.tan(row[3]) / row[7] if abs(row[3]) > 0.00001 else 0
    out = np.empty((len(poses), 5), dtype=np.float64)
    out[:, :3] = poses
    out[: nf + 1, 3] = k_f
    out[: nf + 1, 4] = -1
    out[nf + 1:, 3] = k_b
    out[nf + 1:, 4] = 1
    return out


def search_y_type_parking_path(car_model, config_env, end_pose, backward_steer_dir, forward_steer_dir,
                               max_steer_backward=0.4, max_steer_forward=0.45, max_backward_distance=3.5,
                               max_forward_distance=2.0, min_forward_distance=1.4, min_backward_distance=0.7,
                               min_steer_backward=0.3, min_steer_forward=0.3, step_size=0.1, debug=False):
    """headland_path_planning.py:382-451 -- same arguments and returns; one K5 + K1 + reduce launch."""
    if not config_env.check_path_feasibility(car_model, np.array([end_pose], dtype=np.float64)):
        print(" [Y-type Planner] The end pose is interfered with the environment!")
        return [], []                      # a TUPLE whatever `debug` says, like the reference (:400-402)
    cands = y_park_candidates(max_steer_backward, max_steer_forward, max_backward_distance, max_forward_distance,
                              min_forward_distance, min_backward_distance, min_steer_backward, min_steer_forward)
    if len(cands) == 0:
        return ([], []) if debug else []
    rows = _device_rows(cands, end_pose, backward_steer_dir, forward_steer_dir, car_model.WHEEL_BASE)
    envs = config_env._env_batch(car_model)
    poses, offsets = ops.ypark_paths(rows, step_size)
    bad = ops.collision_check(envs, poses, flags=ops.CHECK_OBSTACLES | ops.CHECK_BOUNDARY)
    import torch
    path_bad = ops.path_reduce(envs, bad, torch.from_numpy(offsets).to(bad.device))
    free = torch.nonzero(path_bad == 0)
    if free.numel() == 0:
        return ([], []) if debug else []
    first = int(free[0].item())
    a, b = int(offsets[first]), int(offsets[first + 1])
    path = _with_ks_dirs(poses[a:b].cpu().numpy(), rows[first], step_size)
    bl, fl, sb, sf = cands[first]
    print("backward distance:%.2f, forward distance:%.2f, backward steer:%.2f, forward steer:%.2f,  " % (bl, fl, sb, sf))
    return (path, [bl, fl, sb, sf]) if debug else path


def search_y_type_parking_path_batch(envs, env_ids, end_poses, backward_steer_dirs, wheel_bases, cands, step_size=0.1):
    """Many sweeps at once.  ``envs``: an ``EnvBatch``; problem ``i`` uses environment ``env_ids[i]``, end pose
    ``end_poses[i]``, backward steer direction ``backward_steer_dirs[i]`` (forward = -backward, :146) and wheel
    base ``wheel_bases[i]``; ``cands`` [C,4] are the shared unsigned candidate rows in loop order
    (``y_park_candidates``).  Returns (first [n] int64: index of the first feasible candidate or -1,
    feasible [n,C] bool, goal [n,3]: first pose of the chosen path, NaN where none)."""
    import torch
    env_ids = np.asarray(env_ids, dtype=np.int32)
    end_poses = np.asarray(end_poses, dtype=np.float64).reshape(-1, 3)
    n, c = len(env_ids), len(cands)
    cands = np.asarray(cands, dtype=np.float64).reshape(-1, 4)
    dev = torch.device("cuda", envs.device)
    # the [n, C, 8] candidate table is built ON THE DEVICE from its factors (a few KB up instead of n * C * 64 bytes);
    # the signed steers are products with +-1, exact on either side
    d_cands = torch.from_numpy(cands).to(dev)
    d_bdir = torch.from_numpy(np.asarray(backward_steer_dirs, dtype=np.float64)).to(dev)[:, None]
    d_end = torch.from_numpy(end_poses).to(dev)
    d_wb = torch.from_numpy(np.asarray(wheel_bases, dtype=np.float64)).to(dev)[:, None]
    rows = torch.empty((n, c, 8), dtype=torch.float64, device=dev)
    rows[:, :, 0] = d_cands[None, :, 0]
    rows[:, :, 1] = d_cands[None, :, 1]
    rows[:, :, 2] = d_cands[None, :, 2] * d_bdir
    rows[:, :, 3] = d_cands[None, :, 3] * -d_bdir
    rows[:, :, 4:7] = d_end[:, None, :]
    rows[:, :, 7] = d_wb
    poses, offsets = ops.ypark_paths(rows.reshape(-1, 8), step_size)
    # environment id of every pose, expanded on the device (plumbing only)
    counts = offsets[1:] - offsets[:-1]
    cand_env = torch.from_numpy(env_ids).to(dev).repeat_interleave(c)
    pose_env = torch.repeat_interleave(cand_env, counts).to(torch.int32)
    bad = ops.collision_check(envs, poses, env_id=pose_env, flags=ops.CHECK_OBSTACLES | ops.CHECK_BOUNDARY)
    path_bad = ops.path_reduce(envs, bad, offsets)
    feasible = (path_bad.reshape(n, c) == 0)
    any_free = feasible.any(dim=1)
    first = torch.where(any_free, feasible.to(torch.int8).argmax(dim=1), torch.full_like(any_free, -1, dtype=torch.int64))
    sel = torch.clamp(first, min=0) + torch.arange(n, device=dev) * c
    goal = poses[offsets[sel]].cpu().numpy()
    first_h = first.cpu().numpy().astype(np.int64)
    goal[first_h < 0] = np.nan
    return first_h, feasible.cpu().numpy(), goal
