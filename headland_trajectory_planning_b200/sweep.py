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
_SECTIONS = (("results", None), ("expanded", 12)) + _PATH_FIELDS


def _gather_layout(sizes, keys_cap, path_cap):
    """Byte layout of rank 0's gathered buffer: one SECTION per output array (result records | expanded keys | x | y |
    yaw | k | dir), and inside a section the ranks' pieces back to back in rank order -- i.e. the rank-major pools
    ``merge_shards`` hands out, so that nothing is copied again after the gather.  ``sizes`` is [world, 3] = (key rows,
    path poses, records) per rank.  Returns (per-section byte offsets [world + 1], per-rank counts, total bytes)."""
    kused = np.minimum(sizes[:, 0], keys_cap).astype(np.int64)
    pused = np.minimum(sizes[:, 1], path_cap).astype(np.int64)
    nrec = sizes[:, 2].astype(np.int64)
    off, o = {}, 0
    for name, w in _SECTIONS:
        if name == "results":
            nbytes = nrec * _lib.RESULT_DTYPE.itemsize
        elif name == "expanded":
            nbytes = kused * w
        else:
            nbytes = pused * w
        off[name] = o + np.concatenate([[0], np.cumsum(nbytes)])
        o = (int(off[name][-1]) + 15) & ~15
    return off, (kused, pused, nrec), max(o, 16)


def _piece(out, name, count):
    """uint8 view of the used part of one output array of this rank."""
    import torch
    if name == "results":
        return out["results"].reshape(-1)[:count * _lib.RESULT_DTYPE.itemsize]
    if name == "expanded":
        return out["expanded"][:count].reshape(-1).view(torch.uint8)
    return out[name][:count].view(torch.uint8)


def gather_sweep(out, world_size, rank):
    """The one data-path collective of a sweep: every rank's search output -- ``ops.hybrid_astar_batch(...,
    to_host=False)``, still on the device -- goes to rank 0 over NCCL / NVLink and is copied to the host ONCE, there.
    Two collectives: an all_gather of three int64 per rank (key rows, path poses, records: what each rank produced)
    and one grouped send / recv (a gatherv: every rank sends exactly the used part of each output array, straight from
    the search's own buffers, and rank 0 receives each piece at its FINAL place in the rank-major pools -- no padding,
    no packing copy, no merge copy).  Ranks other than 0 never copy anything to the host.  With CPU tensors in ``out``
    the same code runs over gloo (tests/test_sweep_gloo.py).

    Returns on rank 0 a dict with the rank-major host pools ``expanded`` [rows, 3] int32, ``x, y, yaw, k, dir``, the
    structured ``results`` of all ranks rank-major, and ``counts`` = per-rank (key rows, path poses, records) for
    ``merge_shards``; ``None`` on the other ranks.  The host arrays are views of one pinned buffer that a later
    gather_sweep call overwrites (copy what must outlive it).  ``SweepDownloader`` is the pipelined form."""
    return gather_sweep_begin(out, world_size, rank).finish()


class _Pending:
    """A sweep output on its way to the host (``gather_sweep_begin`` / ``SweepDownloader.begin``)."""

    def __init__(self, keep, event, build):
        self._keep, self._event, self._build = keep, event, build

    def finish(self):
        """Wait for the device -> host copy and return the host dict (``None`` on ranks other than 0)."""
        if self._event is not None:
            self._event.synchronize()
        out = self._build() if self._build is not None else None
        self._keep = self._build = None                     # device buffers go back to the allocator
        return out


def gather_sweep_begin(out, world_size, rank, stream=None, release=None, slot=0):
    """First half of ``gather_sweep``: the size exchange (the step's one host sync: it returns when this rank's search
    is finished), the gatherv, and on rank 0 the device -> host copy QUEUED on ``stream`` (a side stream: the copy
    engine then moves the gathered sweep while the SMs already run the next search).  ``release`` is called after the
    searches are known to be finished and before the copy is queued (e.g. ``EnvBatch.close``, which synchronises the
    device and would otherwise wait for the copy).  ``slot`` picks one of two pinned buffers, so the arrays of the
    previous sweep stay valid while this one lands.  Returns a handle whose ``finish()`` gives gather_sweep's dict."""
    import torch
    import torch.distributed as dist
    res = out["results"]
    dev = res.device
    n = int(out["n"])
    keys_cap, path_cap = out["expanded"].shape[0], out["x"].shape[0]
    mine = torch.cat([out["kcursor"].reshape(1).to(torch.int64), out["cursor"].reshape(1).to(torch.int64),
                      torch.tensor([n], dtype=torch.int64, device=dev)])
    if world_size > 1:
        sizes = torch.empty(3 * world_size, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, mine)
    else:
        sizes = mine
    sizes = sizes.cpu().numpy().reshape(world_size, 3)               # the step's one host sync on every rank
    off, (kused, pused, nrec), total = _gather_layout(sizes, keys_cap, path_cap)
    count_of = {"results": nrec, "expanded": kused}
    if rank != 0:
        ops_ = []
        for name, _ in _SECTIONS:
            c = int(count_of.get(name, pused)[rank])
            if c:
                ops_.append(dist.P2POp(dist.isend, _piece(out, name, c), 0))
        for w in (dist.batch_isend_irecv(ops_) if ops_ else []):
            w.wait()
        if release is not None:
            release()
        return _Pending(out, None, None)
    big = torch.empty(total, dtype=torch.uint8, device=dev)
    ops_ = []
    for name, _ in _SECTIONS:
        o = off[name]
        c0 = int(count_of.get(name, pused)[0])
        if c0:
            big[int(o[0]):int(o[1])].copy_(_piece(out, name, c0))
        for r in range(1, world_size):
            if o[r + 1] > o[r]:
                ops_.append(dist.P2POp(dist.irecv, big[int(o[r]):int(o[r + 1])], r))
    for w in (dist.batch_isend_irecv(ops_) if ops_ else []):
        w.wait()
    if release is not None:
        release()
    event = None
    if big.is_cuda:
        host = _pinned(total, slot)
        ready = torch.cuda.Event()
        ready.record()                                              # pools complete on the current stream
        st = stream if stream is not None else torch.cuda.current_stream()
        with torch.cuda.stream(st):
            st.wait_event(ready)
            host.copy_(big, non_blocking=True)
            event = torch.cuda.Event()
            event.record()
        host = host.numpy()
    else:
        host = big.numpy()

    def build():
        g = {"counts": np.stack([kused, pused, nrec], axis=1), "keys_produced": sizes[:, 0].copy(),
             "poses_produced": sizes[:, 1].copy()}
        for name, w in _SECTIONS:
            sec = host[int(off[name][0]):int(off[name][-1])]
            if name == "results":
                g[name] = sec.view(_lib.RESULT_DTYPE)
            elif name == "expanded":
                g[name] = sec.view(np.int32).reshape(-1, 3)
            else:
                g[name] = sec.view(np.float64 if w == 8 else np.int8)
        return g
    return _Pending((out, big), event, build)


class SweepDownloader:
    """Pipelined download for back-to-back sweeps (the counterpart of ``UploadPrefetcher``): the device -> host copy
    of sweep k runs on a side stream (copy engine) while the SMs search sweep k+1, so only the size exchange and the
    NVLink gatherv stay between two searches.

        dl = SweepDownloader(world_size, rank)
        pending = None
        for k in range(n):
            out = ops.hybrid_astar_batch(envs_k, scen_k, params, to_host=False)       # queued, asynchronous
            if pending is not None: results_of_k_minus_1 = dl.finish(pending, n_total)
            pending = dl.begin(out, release=envs_k.close)     # returns when search k is done; its copy is in flight
        results_of_last = dl.finish(pending, n_total)

    ``finish`` returns (rank 0) the merged host dict of ``merge_shards`` -- scenario order, shaped like
    ``ops.hybrid_astar_batch(to_host=True)``; its arrays are views of one of two alternating pinned buffers and stay
    valid until the sweep after the next one lands."""

    def __init__(self, world_size=1, rank=0, device=None):
        import torch
        self.world_size, self.rank = int(world_size), int(rank)
        self.stream = torch.cuda.Stream(device) if torch.cuda.is_available() else None
        self._k = 0

    def begin(self, out, release=None):
        self._k ^= 1
        return gather_sweep_begin(out, self.world_size, self.rank, stream=self.stream, release=release, slot=self._k)

    def finish(self, pending, n_total):
        g = pending.finish()
        return merge_shards(g, n_total, self.world_size) if g is not None else None


_pinned_cache = {}


def _pinned(nbytes, slot=0):
    """Grow-only pinned host buffers for the gathered sweep (cudaHostAlloc per step would cost more than the copy)."""
    import torch
    t = _pinned_cache.get(slot)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes + nbytes // 4, 1 << 20), dtype=torch.uint8).pin_memory()
        _pinned_cache[slot] = t
    return t[:nbytes]


def merge_shards(gathered, n_total, world_size):
    """Undo the interleaved sharding (record i of rank r is scenario r + i*world_size).  The key and path pools stay
    rank-major exactly as ``gather_sweep`` received them (views, not copies); only the 80-byte result records are
    put back into scenario order, their ``keys_offset`` / ``path_offset`` rebased into the pools.  Returns a host dict
    shaped like ``ops.hybrid_astar_batch(to_host=True)`` (``ops.expanded_of`` / ``hybrid_a_star_search.unpack_path``
    work on it)."""
    counts = gathered["counts"]
    res = np.zeros(n_total, dtype=_lib.RESULT_DTYPE)
    kbase = pbase = rbase = 0
    for r in range(world_size):
        idx = np.arange(r, n_total, world_size)
        rr = np.array(gathered["results"][rbase:rbase + len(idx)])
        rr["keys_offset"] += kbase
        rr["path_offset"] += pbase
        res[idx] = rr
        kbase += int(counts[r, 0]); pbase += int(counts[r, 1]); rbase += int(counts[r, 2])
    merged = {"results": res, "n": n_total, "keys_used": kbase, "used": pbase, "expanded": gathered["expanded"]}
    for name, _ in _PATH_FIELDS:
        merged[name] = gathered[name]
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
