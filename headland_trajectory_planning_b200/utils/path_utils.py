"""Mirror of the reference's ``path_planner/utils/path_utils.py`` helpers used by
callers on the host (tiny scalar utilities; the per-node versions run inside the
search kernel)."""
import math

import numpy as np


def calculate_path_length(xs, ys):
    """path_utils.py:5-12."""
    return np.cumsum(np.hypot(np.diff(xs), np.diff(ys)))[-1]


def get_projection_point(x_m, y_m, yaw_m, k_m, x, y):
    """path_utils.py:15-23."""
    d = np.array([x - x_m, y - y_m])
    tau = np.array([math.cos(yaw_m), math.sin(yaw_m)])
    return np.array([x_m, y_m]) + d.dot(tau) * tau, yaw_m + k_m * d.dot(tau)


def angle_wrap(angles):
    """path_utils.py:26-29: ``(a + pi) % (2 pi) - pi`` -> [-pi, pi)."""
    return (angles + math.pi) % (2 * math.pi) - math.pi
