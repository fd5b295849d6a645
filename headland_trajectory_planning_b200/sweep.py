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


def search_params(car, step_size=0.2, max_nodes=400, max_path_poses=16384):
    p, _ = make_search_params(car, "King", plan_resolution=step_size, max_nodes=max_nodes,
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


def gather_results(results_np, expanded_np, world_size, rank, device=None):
    """Gather of the result records and pooled expanded keys to rank 0 -- the only collective of a sweep
    (torch.distributed must be initialised).  Records are fixed-stride; the key pools are padded to the
    largest pool (one all_reduce of a scalar).  On the GPU box the tensors live on the device so NCCL moves
    them over NVLink; ``device="cpu"`` is the gloo path the CPU tests use."""
    import torch
    import torch.distributed as dist
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    t_res = torch.from_numpy(np.ascontiguousarray(results_np).view(np.uint8).reshape(-1).copy()).to(dev)
    rows = torch.tensor([len(expanded_np)], dtype=torch.int64, device=dev)
    dist.all_reduce(rows, op=dist.ReduceOp.MAX)
    pad = np.zeros((int(rows.item()), 3), dtype=np.int32)
    pad[:len(expanded_np)] = expanded_np
    t_exp = torch.from_numpy(pad).to(dev)
    if rank == 0:
        g_res = [torch.empty_like(t_res) for _ in range(world_size)]
        g_exp = [torch.empty_like(t_exp) for _ in range(world_size)]
        dist.gather(t_res, g_res, dst=0)
        dist.gather(t_exp, g_exp, dst=0)
        return ([g.cpu().numpy().view(_lib.RESULT_DTYPE) for g in g_res], [g.cpu().numpy() for g in g_exp])
    dist.gather(t_res, None, dst=0)
    dist.gather(t_exp, None, dst=0)
    return None, None


def merge_shards(per_rank_results, per_rank_expanded, n_total, world_size):
    """Undo the interleaved sharding: record i of rank r is scenario r + i*world_size.  Returns the
    merged records (keys_offset rewritten) and one pooled key array in scenario order."""
    res = np.zeros(n_total, dtype=_lib.RESULT_DTYPE)
    chunks = [None] * n_total
    for r in range(world_size):
        idx = shard_indices(n_total, r, world_size)
        rr = per_rank_results[r][:len(idx)]
        res[idx] = rr
        for k, i in enumerate(idx):
            a = int(rr[k]["keys_offset"])
            chunks[i] = per_rank_expanded[r][a:a + int(rr[k]["n_expanded"])]
    off = np.concatenate([[0], np.cumsum([len(c) for c in chunks])])
    res["keys_offset"] = off[:-1]
    exp = np.concatenate(chunks) if n_total else np.zeros((0, 3), np.int32)
    return res, exp


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

    def __init__(self):
        from concurrent.futures import ThreadPoolExecutor
        self._ex = ThreadPoolExecutor(max_workers=1)
        self._fut = None

    def submit(self, records, structs=None):
        from .env_batch import EnvBatch
        self._fut = self._ex.submit(EnvBatch, records, structs=structs)

    def result(self):
        fut, self._fut = self._fut, None
        return fut.result()

    def close(self):
        self._ex.shutdown(wait=True)
