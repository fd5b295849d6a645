"""CPU, world_size 2, gloo: the N > 1 path of a sweep -- interleaved sharding and the
single gather of fixed-stride result records to rank 0 (NCCL on the GPU box)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from headland_trajectory_planning_b200 import _lib, sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = sweep.shard_indices(n_total, rank, world)
    # stand-in shard results: fields derived from the global scenario id
    res = np.zeros(len(sweep.shard_indices(n_total, 0, world)), dtype=_lib.RESULT_DTYPE)
    rows = []
    for k, i in enumerate(idx):
        res[k]["status"] = i % 4
        res[k]["counter"] = 1000 + i
        res[k]["path_offset"] = 7 * i
        res[k]["goal_cost"] = 0.5 * i
        res[k]["keys_offset"] = sum(len(r) for r in rows)
        res[k]["n_expanded"] = 1 + i % 3                      # ragged key slices
        rows.append(np.full((1 + i % 3, 3), i, dtype=np.int32))
    exp = np.concatenate(rows)
    g_res, g_exp = sweep.gather_results(res, exp, world, rank, device="cpu")
    if rank == 0:
        m_res, m_exp = sweep.merge_shards(g_res, g_exp, n_total, world)
        np.save(os.path.join(tmp, "res.npy"), m_res)
        np.save(os.path.join(tmp, "exp.npy"), m_exp)
    else:
        assert g_res is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    n_total = 11                                   # ragged: rank 0 gets 6, rank 1 gets 5
    mp.spawn(_worker, args=(2, 29512, n_total, str(tmp_path)), nprocs=2, join=True)
    res = np.load(os.path.join(tmp_path, "res.npy"))
    exp = np.load(os.path.join(tmp_path, "exp.npy"))
    assert list(res["counter"]) == [1000 + i for i in range(n_total)]
    assert list(res["status"]) == [i % 4 for i in range(n_total)]
    assert list(res["path_offset"]) == [7 * i for i in range(n_total)]
    assert list(res["n_expanded"]) == [1 + i % 3 for i in range(n_total)]
    for i in range(n_total):
        a = int(res["keys_offset"][i])
        assert (exp[a:a + int(res["n_expanded"][i])] == i).all()
    assert len(exp) == int(res["n_expanded"].sum())
