"""Scenario sweeps: shard independent headland scenarios across the GPUs of one box
(one process per GPU, scenario i -> rank i mod world_size), run the batched search on
each shard, and gather fixed-stride result records on rank 0 (the only collective).
"""
import numpy as np

from . import _lib, ops, scenarios as SC
from .env_batch import make_record
from .hybrid_a_star_search import make_search_params, scenario_array


def shard_indices(n_total, rank, world_size):
    """Interleaved shard: evens out the 1..401-expansion imbalance between ranks."""
    return list(range(rank, n_total, world_size))


def build_records(scns):
    """Host geometry of finalized scenarios -> (EnvRecord list, scenario array, car)."""
    recs, starts, goals = [], [], []
    car0 = None
    for scn in scns:
        env, car, heur = SC.build_host_objects(scn)
        car0 = car0 or car
        recs.append(make_record(env, car, heur))
        starts.append(scn["start"])
        goals.append(scn["goal"])
    scen = scenario_array(np.arange(len(scns), dtype=np.int32), starts, goals)
    return recs, scen, car0


def search_params(car, step_size=0.2, max_nodes=400, max_path_poses=16384, motion_type="King"):
    p, _ = make_search_params(car, motion_type, plan_resolution=step_size, max_nodes=max_nodes,
                              max_path_poses=max_path_poses)
    return p


def algorithmic_flops(recs, results, field="n_pose_checks_ref"):
    """FP32 flop of the footprint checks of a sweep, SURVEY.md 8(d): F_check = 32 + 128*K + 80*E_f + 48*S per
    pose rectangle, times the ALGORITHMIC number of pose checks (``n_pose_checks_ref``: what the reference
    tests for the same searches; ``field="n_pose_checks"`` gives the executed count instead)."""
    total = 0.0
    for rec, r in zip(recs, results):
        f = 32 + 128 * len(rec.obs) + 80 * len(rec.field) + 48 * len(rec.seg_xy)
        total += f * float(r[field])
    return total


_PATH_FIELDS = (("x", 8), ("y", 8), ("yaw", 8), ("k", 8), ("dir", 1))


def _pack_layout(n_rows, kmax, pmax):
    """Byte offsets of one rank's packed sweep output: result records | expanded keys | x | y | yaw | k | dir,
    every section 16-byte aligned."""
    al = lambda v: (v + 15) & ~15
    off, o = {}, 0
    off["results"] = o; o += al(n_rows * _lib.RESULT_DTYPE.itemsize)
    off["expanded"] = o; o += al(kmax * 12)
    for name, w in _PATH_FIELDS:
        off[name] = o; o += al(pmax * w)
    return off, o


def gather_sweep(out, world_size, rank):
    """The one data-path collective of a sweep: every rank's search output -- ``ops.hybrid_astar_batch(...,
    to_host=False)``, still on the device -- goes to rank 0 over NCCL / NVLink and is copied to the host ONCE, there.
    Two collectives: an all_gather of three int64 per rank (key rows, path poses, records: what each rank produced)
    and one gather of a packed byte buffer padded to the largest rank.  Ranks other than 0 never copy anything to
    the host.  With CPU tensors in ``out`` the same code runs over gloo (tests/test_sweep_gloo.py).

    Returns on rank 0 a list (one entry per rank) of host dicts with ``results`` (structured), ``expanded``
    [rows, 3] int32 and the pooled path arrays ``x, y, yaw, k, dir``; ``None`` on the other ranks."""
    import torch
    import torch.distributed as dist
    res = out["results"]
    dev = res.device
    n = int(out["n"])
    keys_cap, path_cap = out["expanded"].shape[0], out["x"].shape[0]
    mine = torch.cat([out["kcursor"].reshape(1).to(torch.int64), out["cursor"].reshape(1).to(torch.int64),
                      torch.tensor([n], dtype=torch.int64, device=dev)])
    sizes = torch.empty(3 * world_size, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, mine)
    sizes = sizes.cpu().numpy().reshape(world_size, 3)               # the step's one host sync on every rank
    kused = np.minimum(sizes[:, 0], keys_cap)
    pused = np.minimum(sizes[:, 1], path_cap)
    kmax, pmax, n_rows = int(kused.max()), int(pused.max()), int(sizes[:, 2].max())
    off, total = _pack_layout(n_rows, kmax, pmax)
    buf = torch.empty(total, dtype=torch.uint8, device=dev)
    nb = n * _lib.RESULT_DTYPE.itemsize
    buf[off["results"]:off["results"] + nb].copy_(res.reshape(-1)[:nb])
    k_me, p_me = int(kused[rank]), int(pused[rank])
    if k_me:
        buf[off["expanded"]:off["expanded"] + 12 * k_me].copy_(out["expanded"][:k_me].reshape(-1).view(torch.uint8))
    for name, w in _PATH_FIELDS:
        if p_me:
            buf[off[name]:off[name] + w * p_me].copy_(out[name][:p_me].view(torch.uint8))
    if rank != 0:
        dist.gather(buf, None, dst=0)
        return None
    big = torch.empty((world_size, total), dtype=torch.uint8, device=dev)
    dist.gather(buf, list(big.unbind(0)), dst=0)
    if big.is_cuda:
        host = _pinned(world_size * total).view(world_size, total)
        host.copy_(big, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        host = host.numpy()
    else:
        host = big.numpy()
    shards = []
    for r in range(world_size):
        row = host[r]
        nr, kr, pr = int(sizes[r, 2]), int(kused[r]), int(pused[r])
        d = {"results": row[off["results"]:off["results"] + nr * _lib.RESULT_DTYPE.itemsize].view(_lib.RESULT_DTYPE),
             "expanded": row[off["expanded"]:off["expanded"] + 12 * kr].view(np.int32).reshape(-1, 3),
             "keys_used": int(sizes[r, 0]), "used": int(sizes[r, 1]), "n": nr}
        for name, w in _PATH_FIELDS:
            d[name] = row[off[name]:off[name] + w * pr].view(np.float64 if w == 8 else np.int8)
        shards.append(d)
    return shards


_pinned_cache = {}


def _pinned(nbytes):
    """Grow-only pinned host buffer for the gathered sweep (cudaHostAlloc per step would cost more than the copy)."""
    import torch
    t = _pinned_cache.get("buf")
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8).pin_memory()
        _pinned_cache["buf"] = t
    return t[:nbytes]


def merge_shards(shards, n_total, world_size):
    """Undo the interleaved sharding (record i of rank r is scenario r + i*world_size).  The per-rank key and path
    pools are concatenated rank-major and the records' ``keys_offset`` / ``path_offset`` rebased into them, so
    nothing is re-ordered pose by pose.  Returns a host dict shaped like ``ops.hybrid_astar_batch(to_host=True)``
    (``ops.expanded_of`` / ``hybrid_a_star_search.unpack_path`` work on it)."""
    res = np.zeros(n_total, dtype=_lib.RESULT_DTYPE)
    kbase = pbase = 0
    for r in range(world_size):
        sh = shards[r]
        idx = np.arange(r, n_total, world_size)
        rr = np.array(sh["results"][:len(idx)])
        rr["keys_offset"] += kbase
        rr["path_offset"] += pbase
        res[idx] = rr
        kbase += len(sh["expanded"])
        pbase += len(sh["x"])
    merged = {"results": res, "n": n_total, "keys_used": kbase, "used": pbase,
              "expanded": np.concatenate([sh["expanded"] for sh in shards]) if shards else np.zeros((0, 3), np.int32)}
    for name, _ in _PATH_FIELDS:
        merged[name] = np.concatenate([sh[name] for sh in shards])
    return merged


class UploadPrefetcher:
    """Double buffering for back-to-back sweeps: ``hl_env_upload`` of the NEXT batch (host packing + H2D copy on the
    context's own copy stream) runs in a worker thread while the current batch is being searched, so the
    upload disappears from the critical path.  ctypes releases the GIL inside the C calls.

        pf = UploadPrefetcher()
        pf.submit(records, structs)                 # batch 0
        for k in range(n):
            envs = pf.result()
            if k + 1 < n: pf.submit(records_next, structs_next)
            out = ops.hybrid_astar_batch(envs, scen, params)
            envs.close()
    """

    def __init__(self, device=None):
        import torch
        from concurrent.futures import ThreadPoolExecutor
        # The CUDA current device is PER THREAD and a new thread starts on device 0: capture the device of the
        # constructing thread (the rank's own GPU) and hand it to the worker explicitly.
        self.device = int(torch.cuda.current_device() if device is None else device)
        self._ex = ThreadPoolExecutor(max_workers=1, initializer=self._init_worker, initargs=(self.device,))
        self._fut = None

    @staticmethod
    def _init_worker(device):
        import torch
        torch.cuda.set_device(device)

    def submit(self, records, structs=None):
        from .env_batch import EnvBatch
        self._fut = self._ex.submit(EnvBatch, records, device=self.device, structs=structs)

    def result(self):
        fut, self._fut = self._fut, None
        return fut.result()

    def close(self):
        self._ex.shutdown(wait=True)
