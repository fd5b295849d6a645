#!/usr/bin/env python
"""Short driver for ncu: one warm-up + one measured launch of each hot kernel
(k_hybrid_astar on N config-5 scenarios, k_collision on the canonical scenario,
k_rs_all_paths on random pose pairs).  Usage under gpurun:

    python tools/profile_kernels.py [n_scenarios] [n_poses] [n_pairs]
"""
import argparse
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from headland_trajectory_planning_b200 import ops, scenarios as SC, sweep
    from headland_trajectory_planning_b200.env_batch import EnvBatch
    ap = argparse.ArgumentParser()
    ap.add_argument("n_scen", nargs="?", type=int, default=512)
    ap.add_argument("n_poses", nargs="?", type=int, default=1 << 22)
    ap.add_argument("n_pairs", nargs="?", type=int, default=1 << 16)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    scns = SC.make_scenarios_gpu(list(range(a.n_scen)))
    recs, scen, car = sweep.build_records(scns)
    envs = EnvBatch(recs)
    params = sweep.search_params(car)
    d_scen = torch.from_numpy(scen.view(np.uint8).reshape(-1)).to(dev)
    for _ in range(2):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        out = ops.hybrid_astar_batch(envs, d_scen, params, path_capacity=1024 * a.n_scen, to_host=False)
        a1.record()
        torch.cuda.synchronize()
    print("k_hybrid_astar", a.n_scen, "scenarios:", a0.elapsed_time(a1), "ms")
    ph = ops.astar_phase_cycles()
    tot = sum(ph.values()) or 1
    print("phase share of thread-0 cycles:", {k: round(100.0 * v / tot, 1) for k, v in ph.items()})
    res = out["results"].cpu().numpy().view(__import__("headland_trajectory_planning_b200")._lib.RESULT_DTYPE)
    print("expansions", int(res["n_expanded"].sum()), "cycles/expansion", tot / 2 / max(1, int(res["n_expanded"].sum())))

    class A:
        collision_poses = a.n_poses
    peak = ops.measure_fp32_peak(0)
    print("fp32 peak TFLOP/s", peak)
    print("k_collision", bench.collision_microbench(A, dev, peak))
    rng = np.random.default_rng(0)
    sg = np.empty((a.n_pairs, 6))
    sg[:, [0, 1, 3, 4]] = rng.uniform(-10, 10, (a.n_pairs, 4))
    sg[:, [2, 5]] = rng.uniform(-math.pi, math.pi, (a.n_pairs, 2))
    d_sg = torch.from_numpy(sg).to(dev)
    for _ in range(2):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        ops.rs_all_paths(d_sg, math.tan(0.55) / 1.9, 0.1)
        a1.record()
        torch.cuda.synchronize()
    print("k_rs_all_paths", a.n_pairs, "pairs:", a0.elapsed_time(a1), "ms ->", a.n_pairs / a0.elapsed_time(a1) * 1e3, "pairs/s")


if __name__ == "__main__":
    main()
