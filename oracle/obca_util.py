"""Warm-start -> OBCA hand-off formatting: ``obca_py/util.py:7-113`` +
``path_planner/utils/cubic_spline.py:19-112``.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference splits the planner's path by driving direction, fits scipy's not-a-knot ``CubicSpline`` x(s), y(s)
over the chord length of every piece, resamples at ``ds`` and emits rows (x, y, v, yaw, steer).  scipy IS
installed, so this restatement calls ``CubicSpline`` exactly like the reference; it is pinned on the reference's
own ``get_init_ref_path`` (``tests/golden/refpath_golden.npz``, ``oracle/gen_golden.py refpath``)."""
import math

import numpy as np
from scipy.interpolate import CubicSpline


def wrap_angle(angle):
    """obca_py/util.py:7-13."""
    return (angle + math.pi) % (2 * math.pi) - math.pi


def convert_angle_to_monotonic(raw_angles):
    """obca_py/util.py:29-44."""
    if len(raw_angles) <= 1:
        return np.copy(raw_angles)
    out = np.zeros(len(raw_angles))
    out[0] = raw_angles[0]
    for i in range(1, len(raw_angles)):
        out[i] = out[i - 1] + wrap_angle(raw_angles[i] - raw_angles[i - 1])
    return out


def process_angle(raw_angles):
    """obca_py/util.py:16-26."""
    adjusted = np.zeros_like(raw_angles)
    for i in range(len(adjusted)):
        adjusted[i] = wrap_angle(raw_angles[i])
    return convert_angle_to_monotonic(adjusted)


def calc_spline_course(x, y, ds=0.1):
    """cubic_spline.py:92-112 with Spline2D (:19-90)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    singular = np.where((np.diff(x) == 0) & (np.diff(y) == 0))
    x, y = np.delete(x, singular, axis=0), np.delete(y, singular, axis=0)
    s = [0]
    s.extend(np.cumsum(np.hypot(np.diff(x), np.diff(y))))
    sx, sy = CubicSpline(s, x), CubicSpline(s, y)
    ss = list(np.arange(0, s[-1] + ds, ds))
    rx, ry, ryaw, rk = [], [], [], []
    for i_s in ss:
        rx.append(sx(i_s)); ry.append(sy(i_s))
        dx, dy = np.asarray(sx(i_s, 1)), np.asarray(sy(i_s, 1))
        ddx, ddy = np.asarray(sx(i_s, 2)), np.asarray(sy(i_s, 2))
        ryaw.append(np.arctan2(dy, dx))
        rk.append((ddy * dx - ddx * dy) / ((dx ** 2 + dy ** 2) ** (3.0 / 2.0)))
    return rx, ry, ryaw, rk, ss


def get_init_ref_path(wheel_base, path_xs, path_ys, path_yaws, path_ks, dirs, desired_v=0.5, ds=0.1):
    """obca_py/util.py:62-113 -> rows (x, y, v, yaw, steer)."""
    ref_path = np.vstack([path_xs, path_ys, path_yaws, path_ks, dirs]).T
    dividers = np.where(np.diff(ref_path[:, -1]) != 0)[0]
    pieces = []
    if len(dividers) > 0:
        for i, idx in enumerate(dividers):
            pieces.append(ref_path[: idx + 1] if i == 0 else ref_path[dividers[i - 1] + 1: idx + 1])
        pieces.append(ref_path[idx + 1:])
    else:
        pieces.append(ref_path)
    ref_traj = np.array([])
    for path in pieces:
        xs, ys, yaws, ks, _ = calc_spline_course(path[:, 0], path[:, 1], ds=ds)
        if path[-1, -1] < 0:
            yaws = wrap_angle(np.array(yaws) + np.pi)
            steer_dir = -1
        else:
            steer_dir = 1
        steers = np.arctan(wheel_base * np.array(ks)) * steer_dir
        vs = np.ones_like(xs) * path[0, -1] * desired_v
        vs[0] = 0
        steers[0] = 0
        traj = np.vstack([xs, ys, vs, yaws, steers]).T
        ref_traj = traj if len(ref_traj) == 0 else np.vstack([ref_traj, traj])
    ref_traj[:, 3] = process_angle(ref_traj[:, 3])
    ref_traj[0, 2] = 0
    ref_traj[-1, 2] = 0
    return ref_traj
