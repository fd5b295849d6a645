"""Multi-GPU end to end (needs >= 2 CUDA devices; skipped on a 1-GPU box): the regression of round 1 was a
prefetch thread uploading every rank's environments to GPU 0.  These tests run the SAME code path as
``bench.py --gpus N``'s end-to-end arm -- UploadPrefetcher worker thread, search on the rank's own device, NCCL
gather of the device-resident output, merge on rank 0 -- and compare with one single-GPU sweep."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_two():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")


def _golden(n):
    from headland_trajectory_planning_b200 import scenarios as SC
    g = np.load(os.path.join(ROOT, "tests", "golden", "astar_golden.npz"))
    specs = [SC.scenario_spec(int(i)) for i in g["index"][:n]]
    return [SC.finalize(sp, f) for sp, f in zip(specs, g["feas"][:n])], g


def test_prefetcher_uploads_to_the_callers_device(built_library):
    """UploadPrefetcher + hybrid_astar_batch on cuda:1 == the same on cuda:0 (the worker thread must inherit the
    device of the thread that created it, not CUDA's per-thread default 0)."""
    _need_two()
    import torch
    from headland_trajectory_planning_b200 import ops, sweep
    from headland_trajectory_planning_b200.env_batch import pack_structs
    scns, g = _golden(64)
    recs, scen, car = sweep.build_records(scns)
    params = sweep.search_params(car)
    structs = pack_structs(recs)
    outs = []
    for d in (0, 1):
        with torch.cuda.device(d):
            pf = sweep.UploadPrefetcher()
            assert pf.device == d
            pf.submit(recs, structs)
            envs = pf.result()
            assert envs.device == d
            outs.append(ops.hybrid_astar_batch(envs, scen, params, path_capacity=2048 * 64))
            envs.close()
            pf.close()
    a, b = outs
    for f in ("status", "counter", "n_expanded", "arrival", "path_len", "rs_word", "goal_cost", "n_pose_checks_ref"):
        assert np.array_equal(a["results"][f], b["results"][f]), f
    assert np.array_equal(a["results"]["status"], g["status"][:64])
    for i in range(64):
        assert np.array_equal(ops.expanded_of(a, i), ops.expanded_of(b, i))


def test_device_mismatch_is_rejected(built_library):
    """An environment batch of another device, or tensors of another device, raise HeadlandError (Python check and,
    behind it, hl_enter in the C ABI) instead of launching across devices."""
    _need_two()
    import ctypes as C
    import torch
    from headland_trajectory_planning_b200 import _lib, ops, sweep
    from headland_trajectory_planning_b200.env_batch import EnvBatch
    scns, _ = _golden(4)
    recs, scen, car = sweep.build_records(scns)
    params = sweep.search_params(car)
    envs1 = EnvBatch(recs, device=1)
    with torch.cuda.device(0):
        with pytest.raises(_lib.HeadlandError, match="cuda:1"):
            ops.hybrid_astar_batch(envs1, scen, params)
        # straight through the C ABI: context of device 0, batch of device 1
        lib = _lib.load_library()
        out = torch.zeros(8, dtype=torch.uint8, device="cuda:0")
        poses = torch.zeros((8, 3), dtype=torch.float64, device="cuda:0")
        rc = lib.hl_collision_check(_lib.get_ctx(0), envs1.handle, None, _lib.ptr(poses), None, 8, 1, _lib.ptr(out),
                                    None, _lib.stream_ptr())
        assert rc != 0 and "device 1" in _lib.last_error()
        # context of device 1 and its own batch, but an output buffer that lives on device 0
        rc = lib.hl_collision_check(_lib.get_ctx(1), envs1.handle, None, _lib.ptr(poses), None, 8, 1, _lib.ptr(out),
                                    None, C.c_void_p(0))
        assert rc != 0 and "device 0" in _lib.last_error()
    envs1.close()


def _rank_main(rank, world, port, n_total, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from headland_trajectory_planning_b200 import ops, sweep
    from headland_trajectory_planning_b200.env_batch import pack_structs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    scns, _ = _golden(n_total)
    mine = [scns[i] for i in sweep.shard_indices(n_total, rank, world)]
    recs, scen, car = sweep.build_records(mine)
    params = sweep.search_params(car)
    pf = sweep.UploadPrefetcher()
    pf.submit(recs, pack_structs(recs))
    envs = pf.result()
    out = ops.hybrid_astar_batch(envs, scen, params, path_capacity=2048 * len(mine), to_host=False)
    shards = sweep.gather_sweep(out, world, rank)
    if rank == 0:
        m = sweep.merge_shards(shards, n_total, world)
        np.savez(os.path.join(tmp, "merged.npz"), **{k: v for k, v in m.items() if isinstance(v, np.ndarray)})
    envs.close()
    pf.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_sweep_equals_single_gpu(built_library, tmp_path):
    _need_two()
    import torch.multiprocessing as mp
    from headland_trajectory_planning_b200 import ops, sweep
    from headland_trajectory_planning_b200.env_batch import EnvBatch
    from headland_trajectory_planning_b200.hybrid_a_star_search import unpack_path
    n_total = 97                                   # ragged shards
    mp.spawn(_rank_main, args=(2, 29533, n_total, str(tmp_path)), nprocs=2, join=True)
    m = dict(np.load(os.path.join(tmp_path, "merged.npz")))
    scns, g = _golden(n_total)
    recs, scen, car = sweep.build_records(scns)
    one = ops.hybrid_astar_batch(EnvBatch(recs), scen, sweep.search_params(car), path_capacity=2048 * n_total)
    for f in ("status", "counter", "n_expanded", "arrival", "path_len", "rs_word", "goal_cost", "n_pose_checks_ref"):
        assert np.array_equal(m["results"][f], one["results"][f]), f
    assert np.array_equal(m["results"]["status"], g["status"][:n_total])
    for i in range(n_total):
        assert np.array_equal(ops.expanded_of(m, i), ops.expanded_of(one, i))
        assert unpack_path(m, i) == unpack_path(one, i)


def test_pipelined_upload_search_download_single_gpu(built_library):
    """bench.py's end-to-end loop on one GPU: UploadPrefetcher + SweepDownloader keep three sweeps in flight (upload of
    k+1, search of k, device -> host copy of k-1 on a side stream); every sweep's merged host result must equal the
    plain ``to_host=True`` call, and the arrays of sweep k must stay intact while sweep k+1 lands."""
    from headland_trajectory_planning_b200 import ops, sweep
    from headland_trajectory_planning_b200.env_batch import EnvBatch, pack_structs
    from headland_trajectory_planning_b200.hybrid_a_star_search import unpack_path
    n = 96
    scns, g = _golden(n)
    subsets = [list(range(0, n)), list(range(0, n, 2)), list(range(1, n, 3)), list(range(5, 60))]
    built = []
    for sub in subsets:
        recs, scen, car = sweep.build_records([scns[i] for i in sub])
        built.append((recs, pack_structs(recs), scen, sweep.search_params(car)))
    want = []
    for recs, _, scen, params in built:
        envs = EnvBatch(recs)
        want.append(ops.hybrid_astar_batch(envs, scen, params, path_capacity=2048 * len(recs)))
        envs.close()
    def same(m, w):
        for f in ("status", "counter", "n_expanded", "arrival", "path_len", "rs_word", "goal_cost", "n_pose_checks_ref"):
            assert np.array_equal(m["results"][f], w["results"][f]), f
        for i in range(len(w["results"])):
            assert np.array_equal(ops.expanded_of(m, i), ops.expanded_of(w, i))
            assert unpack_path(m, i) == unpack_path(w, i)

    pf, dl = sweep.UploadPrefetcher(), sweep.SweepDownloader()
    done, pending, prev = 0, None, None
    pf.submit(built[0][0], built[0][1])
    for k, (recs, structs, scen, params) in enumerate(built):
        envs = pf.result()
        if k + 1 < len(built):
            pf.submit(built[k + 1][0], built[k + 1][1])
        out = ops.hybrid_astar_batch(envs, scen, params, path_capacity=2048 * len(recs), to_host=False)
        if pending is not None:
            prev = dl.finish(pending[0], pending[1])
            same(prev, want[done])
        pending = (dl.begin(out, release=envs.close), len(recs))
        if prev is not None:                    # the views of sweep k-1 stay valid while sweep k lands (two pinned buffers)
            same(prev, want[done])
            done += 1
            prev = None
    same(dl.finish(pending[0], pending[1]), want[done])
    pf.close()
