"""Thin Python wrappers over the C ABI (device tensors in, device tensors out)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CHECK_AUX, CHECK_BOUNDARY, CHECK_LANE, CHECK_OBSTACLES  # noqa: F401


def _torch():
    import torch
    return torch


def collision_check(envs, poses, env_id=None, pose_idx=None, flags=CHECK_OBSTACLES | CHECK_BOUNDARY,
                    count_exact=False):
    """Per-pose infeasibility flags (uint8 CUDA tensor).  ``poses``: [N,3] float64
    (CUDA tensor, or host array which is copied).  ``pose_idx``: index of each pose
    within its path (implement rectangles are tested at even indices only)."""
    torch = _torch()
    lib = _lib.load_library()
    dev = torch.device("cuda", torch.cuda.current_device())
    if not torch.is_tensor(poses):
        poses = torch.from_numpy(np.ascontiguousarray(np.asarray(poses, dtype=np.float64)[:, :3])).to(dev)
    poses = poses.contiguous()
    n = poses.shape[0]
    if env_id is not None and not torch.is_tensor(env_id):
        env_id = torch.from_numpy(np.ascontiguousarray(env_id, dtype=np.int32)).to(dev)
    if pose_idx is not None and not torch.is_tensor(pose_idx):
        pose_idx = torch.from_numpy(np.ascontiguousarray(pose_idx, dtype=np.int32)).to(dev)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    n_exact = torch.zeros(1, dtype=torch.int64, device=dev) if count_exact else None
    _lib.check(lib.hl_collision_check(envs.ctx, envs.handle, _lib.ptr(env_id), _lib.ptr(poses),
                                      _lib.ptr(pose_idx), n, flags, _lib.ptr(out), _lib.ptr(n_exact),
                                      _lib.stream_ptr()), "hl_collision_check")
    if count_exact:
        return out, n_exact
    return out


def path_reduce(envs, pose_bad, path_start):
    torch = _torch()
    lib = _lib.load_library()
    n_paths = path_start.shape[0] - 1
    out = torch.empty(n_paths, dtype=torch.uint8, device=pose_bad.device)
    _lib.check(lib.hl_path_reduce(envs.ctx, _lib.ptr(pose_bad), _lib.ptr(path_start), n_paths,
                                  _lib.ptr(out), _lib.stream_ptr()), "hl_path_reduce")
    return out


def measure_fp32_peak(device=None):
    lib = _lib.load_library()
    v = C.c_double()
    _lib.check(lib.hl_measure_fp32_peak(_lib.get_ctx(device), C.byref(v)), "hl_measure_fp32_peak")
    return v.value
