"""CPU, world_size 2, gloo: the N > 1 path of a sweep -- interleaved sharding, the packed gather of every
rank's search output to rank 0 (``sweep.gather_sweep``: NCCL on the GPU box, the very same code over gloo here) and
the merge back into scenario order.  The shard outputs are laid out exactly like ``ops.hybrid_astar_batch(...,
to_host=False)`` returns them (uint8 record bytes, pooled keys / paths with cursors, spare capacity behind the used
part); the kernels themselves need a GPU -- the 2-GPU end-to-end check is tests/test_multigpu_gpu.py."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_output(idx, n_rows_cap):
    """Stand-in for one rank's device output; every field derives from the global scenario id."""
    import torch
    from headland_trajectory_planning_b200 import _lib
    n = len(idx)
    res = np.zeros(n, dtype=_lib.RESULT_DTYPE)
    keys, xs, dirs = [], [], []
    for k, i in enumerate(idx):
        res[k]["status"] = i % 4
        res[k]["counter"] = 1000 + i
        res[k]["goal_cost"] = 0.5 * i
        res[k]["keys_offset"] = sum(len(r) for r in keys)
        res[k]["n_expanded"] = 1 + i % 3                      # ragged key slices
        keys.append(np.full((1 + i % 3, 3), i, dtype=np.int32))
        res[k]["path_offset"] = sum(len(r) for r in xs)
        res[k]["path_len"] = (i * 7) % 5                      # ragged paths, some empty
        xs.append(np.full((i * 7) % 5, float(i)))
        dirs.append(np.full((i * 7) % 5, 1 if i % 2 else -1, dtype=np.int8))
    keys, xs, dirs = np.concatenate(keys), np.concatenate(xs), np.concatenate(dirs)
    kcap, pcap = n_rows_cap * 4, n_rows_cap * 6               # capacity > used, like the real buffers

    def pad(a, cap, dtype):
        o = np.full((cap,) + a.shape[1:], 99, dtype=dtype)
        o[:len(a)] = a
        return torch.from_numpy(o)
    return dict(results=torch.from_numpy(res.view(np.uint8).reshape(-1).copy()), n=n,
                expanded=pad(keys, kcap, np.int32), kcursor=torch.tensor([len(keys)], dtype=torch.int64),
                x=pad(xs, pcap, np.float64), y=pad(xs + 0.25, pcap, np.float64), yaw=pad(-xs, pcap, np.float64),
                k=pad(xs * 0.5, pcap, np.float64), dir=pad(dirs, pcap, np.int8),
                cursor=torch.tensor([len(xs)], dtype=torch.int64))


def _worker(rank, world, port, n_total, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from headland_trajectory_planning_b200 import sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = sweep.shard_indices(n_total, rank, world)
    out = _shard_output(idx, len(sweep.shard_indices(n_total, 0, world)))
    shards = sweep.gather_sweep(out, world, rank)
    if rank == 0:
        m = sweep.merge_shards(shards, n_total, world)
        np.savez(os.path.join(tmp, "merged.npz"), **{k: v for k, v in m.items() if isinstance(v, np.ndarray)})
    else:
        assert shards is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    sys.path.insert(0, ROOT)
    from headland_trajectory_planning_b200 import ops
    from headland_trajectory_planning_b200.hybrid_a_star_search import unpack_path
    n_total = 11                                   # ragged: rank 0 gets 6, rank 1 gets 5
    mp.spawn(_worker, args=(2, 29512, n_total, str(tmp_path)), nprocs=2, join=True)
    m = dict(np.load(os.path.join(tmp_path, "merged.npz")))
    res = m["results"]
    assert list(res["counter"]) == [1000 + i for i in range(n_total)]
    assert list(res["status"]) == [i % 4 for i in range(n_total)]
    assert list(res["n_expanded"]) == [1 + i % 3 for i in range(n_total)]
    assert list(res["path_len"]) == [(i * 7) % 5 for i in range(n_total)]
    assert len(m["expanded"]) == int(res["n_expanded"].sum()) and len(m["x"]) == int(res["path_len"].sum())
    for i in range(n_total):
        assert (ops.expanded_of(m, i) == i).all() and len(ops.expanded_of(m, i)) == 1 + i % 3
        x, y, yaw, dirs, ks = unpack_path(m, i)
        assert len(x) == (i * 7) % 5
        assert all(v == float(i) for v in x) and all(v == i + 0.25 for v in y) and all(v == -float(i) for v in yaw)
        assert all(v == 0.5 * i for v in ks) and all(d == (1 if i % 2 else -1) for d in dirs)
