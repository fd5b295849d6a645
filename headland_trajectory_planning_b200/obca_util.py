"""Mirror of the warm-start -> OBCA hand-off of ``obca_py/util.py``.

``get_init_ref_path`` (:62-113) turns the planner's (x, y, yaw, k, dir) path into the initial guess the OBCA solve
starts from: rows (x, y, v, yaw, steer) resampled at ``ds`` along not-a-knot cubic splines, one spline per
driving-direction piece.  Here it is K8 (``hl_ref_path_count`` / ``hl_ref_path_fill``); the batch form takes the pooled
output of ``hl_hybrid_astar_batch`` directly, so a sweep's paths never leave the GPU before they are OBCA-ready."""
import numpy as np

from . import ops


def get_init_ref_path(car, path_xs, path_ys, path_yaws, path_ks, dirs, desired_v=0.5, ds=0.1):
    """obca_py/util.py:62-113 -> [T,5] rows (x, y, v, yaw, steer).  (``path_yaws`` / ``path_ks`` are not used by the
    reference either: yaw and steer come from the splines.)"""
    n = len(path_xs)
    traj, off, status = ops.ref_path_batch(np.asarray(path_xs, dtype=np.float64), np.asarray(path_ys, dtype=np.float64),
                                           np.asarray(dirs, dtype=np.float64).astype(np.int8), [0, n], car.WHEEL_BASE,
                                           desired_v, ds)
    if status[0]:
        raise ValueError("`x` must contain at least 2 elements.")          # what scipy raises inside the reference
    return traj.cpu().numpy()


def get_init_ref_path_batch(out, wheel_base, desired_v=0.5, ds=0.1):
    """Initial guesses of every path of a ``ops.hybrid_astar_batch(..., to_host=False)`` result.
    Returns (traj [T,5] CUDA tensor, offsets [n+1], status [n]); scenarios without a path get no rows."""
    import torch
    res = out["results"]
    if torch.is_tensor(res):
        from . import _lib
        res = res.cpu().numpy().view(_lib.RESULT_DTYPE)
    n = len(res)
    # per-scenario slices of the pooled path buffers (bump-allocated, so not in scenario order)
    lens = res["path_len"].astype(np.int64)
    starts = res["path_offset"].astype(np.int64)
    # gather into scenario order on the device
    idx = np.concatenate([np.arange(s, s + l) for s, l in zip(starts, lens)]) if lens.sum() else np.zeros(0, np.int64)
    gi = torch.from_numpy(idx).to(out["x"].device) if torch.is_tensor(out["x"]) else idx
    x, y, d = out["x"][gi], out["y"][gi], out["dir"][gi]
    in_off = np.concatenate([[0], np.cumsum(lens)])
    return ops.ref_path_batch(x, y, d, in_off, wheel_base, desired_v, ds)
