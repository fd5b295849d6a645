// hl_rs.cu -- K2/K3: batched Reeds-Shepp word evaluation, sampling and shots.
//
// One warp per start/goal pair: lanes evaluate the 46 candidate rows, lane 0
// replays set_path's order-dependent dedup and the heapdict pop order, then one
// lane per accepted word builds its sampling plan and (when an environment is
// given) the whole warp samples each word and collision-checks it with ballot
// early exit.  Replaces reeds_shepp.calc_all_paths (reeds_shepp.py:39-65) and the
// per-path loop of _get_goal_extension_with_reeds_shepp_path
// (hybrid_a_star_search.py:249-287).
// Code bytes are what this kernel pays for (one pair walks ~100 KB of solver / sampler / filter code), so the
// helpers are built out of line with their loops rolled, like in the search kernels.
#define HL_SHARED_CODE 1
#include <cstring>
#include "hl_geom.cuh"
#include "hl_rs.cuh"

#define RS_WARPS 4
#define RS_PLANS 6              // sampling plans built per round: one lane per (word, segment), rs_make_plans_warp

struct RsWarpSmem {
    double lens[HL_RS_CANDIDATES][HL_RS_MAX_SEGS];
    double L[HL_RS_CANDIDATES], Lc[HL_RS_CANDIDATES];
    double prio[HL_RS_CANDIDATES];
    int acc[HL_RS_CANDIDATES];
    int order[HL_RS_CANDIDATES];
    unsigned char valid[HL_RS_CANDIDATES + 2], accept[HL_RS_CANDIDATES + 2];
    RsProblem prob;
    RsPlan plans[RS_PLANS];
};

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rs_all_paths(EnvBatchDev eb, int have_env, const int32_t* __restrict__ env_id,
               const double* __restrict__ sg, long long n, double maxc, double step, double max_steer,
               unsigned flags, HlRsWord* __restrict__ words, int32_t* __restrict__ count,
               int32_t* __restrict__ order_out) {
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsWarpSmem* sm = reinterpret_cast<RsWarpSmem*>(rs_smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    RsWarpSmem& W = sm[wid];
    const long long n_warps = (long long)gridDim.x * RS_WARPS;
    const double stepn = xmul(step, maxc);
    const float inv_maxc = (float)(1.0 / maxc);
    const bool body_only = !(flags & HL_CHECK_AUX);
    for (long long i = (long long)blockIdx.x * RS_WARPS + wid; i < n; i += n_warps) {
        double q0[3] = {sg[6 * i], sg[6 * i + 1], sg[6 * i + 2]};
        double q1[3] = {sg[6 * i + 3], sg[6 * i + 4], sg[6 * i + 5]};
        double cq, sq;
        {
            // generate_path's normalisation (rs_normalise): the two sine / cosine pairs are one sincos each on two lanes
            // (sincos == sin, cos and sin odd / cos even bit for bit, tools/sincos_check.cu); the start heading's pair
            // also gives cos(-yaw), sin(-yaw) of the local -> world rotation
            const double phi = xsub(q1[2], q0[2]);
            double sn, cs;
            m_sincos(lane == 1 ? phi : q0[2], &sn, &cs);
            const double c0 = __shfl_sync(0xffffffffu, cs, 0), s0 = __shfl_sync(0xffffffffu, sn, 0);
            const double cp = __shfl_sync(0xffffffffu, cs, 1), sp = __shfl_sync(0xffffffffu, sn, 1);
            cq = c0; sq = -s0;
            if (lane == 0) {
                RsProblem Pr;
                const double dx = xsub(q1[0], q0[0]), dy = xsub(q1[1], q0[1]);
                Pr.phi = phi;
                Pr.x = xmul(xadd(xmul(c0, dx), xmul(s0, dy)), maxc);
                Pr.y = xmul(xadd(xmul(-s0, dx), xmul(c0, dy)), maxc);
                Pr.sp = sp; Pr.cp = cp;
                Pr.xb = xadd(xmul(Pr.x, cp), xmul(Pr.y, sp));
                Pr.yb = xsub(xmul(Pr.x, sp), xmul(Pr.y, cp));
                W.prob = Pr;
            }
            __syncwarp();
        }
        // ---- the 46 word solvers: two passes of lock-step transcendental slots (rs_candidates_warp)
        rs_candidates_warp(W.prob, W.valid, W.lens, lane);
        // ---- set_path dedup per letter group (lane-parallel), compaction in evaluation order, costs, heapdict order
        if (lane < RS_N_GROUPS) rs_select_group(lane, W.valid, W.lens, W.accept, W.Lc);
        __syncwarp();
        int m;
        {
            const int a0 = W.accept[lane];
            const int a1 = (lane + 32 < HL_RS_CANDIDATES) ? W.accept[lane + 32] : 0;
            const unsigned b0 = __ballot_sync(0xffffffffu, a0 == 1), b1 = __ballot_sync(0xffffffffu, a1 == 1);
            const unsigned bad = __ballot_sync(0xffffffffu, a0 == 2 || a1 == 2);
            const unsigned lt = (1u << lane) - 1u;
            const int n0 = __popc(b0);
            if (a0 == 1) { int k = __popc(b0 & lt); W.acc[k] = lane; W.L[k] = W.Lc[lane]; }
            if (a1 == 1) { int k = n0 + __popc(b1 & lt); W.acc[k] = lane + 32; W.L[k] = W.Lc[lane + 32]; }
            m = bad ? -1 : n0 + __popc(b1);          // -1: the reference's `assert path.L >= 0.01` would fire
        }
        __syncwarp();
        const int mm = m < 0 ? 0 : m;
        for (int k = lane; k < mm; k += 32)
            W.prio[k] = rs_path_cost(0.0, W.acc[k], W.lens[W.acc[k]], max_steer, 5000.0, 1000.0, 1.0);
        __syncwarp();
        if (lane == 0) {
            if (mm > 0) heapdict_order(W.prio, mm, W.order);
            count[i] = m;
        }
        __syncwarp();
        HlRsWord* out = words + (size_t)i * HL_RS_CANDIDATES;
        if (order_out)
            for (int k = lane; k < HL_RS_CANDIDATES; k += 32)
                order_out[(size_t)i * HL_RS_CANDIDATES + k] = (k < mm) ? W.order[k] : -1;
        const int e = (have_env && env_id) ? env_id[i] : 0;
        const EnvDesc* Dp = have_env ? &eb.desc[e] : nullptr;
        EnvSmem E;
        if (have_env && mm > 0) {
            global_env(eb, *Dp, E);
            E.eps += 6e-5f;                               // float32 sampling error of rs_sample_world32
        }
        // ---- words in rounds of RS_PLANS: plans lane-parallel, records out, then each word sampled and checked
        HL_LOOP
        for (int k0 = 0; k0 < mm; k0 += RS_PLANS) {
            const int nk = (mm - k0) < RS_PLANS ? (mm - k0) : RS_PLANS;
            __syncwarp();
            rs_make_plans_warp(W.acc, W.lens, nullptr, k0, nk, W.plans, q0, cq, sq, have_env ? Dp->origin : nullptr, maxc, stepn, lane);
            if (lane < nk) {
                const int k = k0 + lane;
                const int c = W.acc[k];
                const RsPlan& plan = W.plans[lane];
                HlRsWord w;
                w.cand = c; w.n_seg = plan.nseg; w.npts = plan.npts; w.collide = -1;
                w.L = xdiv(W.L[k], maxc);
                w.cost = W.prio[k];
                for (int s = 0; s < HL_RS_MAX_SEGS; ++s) {
                    w.len[s] = (s < plan.nseg) ? xdiv(W.lens[c][s], maxc) : 0.0;
                    w.nlen[s] = (s < plan.nseg) ? W.lens[c][s] : 0.0;
                }
                out[k] = w;
            }
            __syncwarp();
            if (!have_env) continue;
            const EnvDesc& D = *Dp;
            HL_LOOP
            for (int r = 0; r < nk; ++r) {
                const RsPlan& plan = W.plans[r];
                const int npts = plan.npts;
                int hit = 0;
                if (body_only) {
                    // float32 world samples, poses strided over the word so that a collision anywhere shows up in
                    // the first pass; float64 only inside the error band (same predicate as the oracle's)
                    const int passes = (npts + 31) >> 5;
                    for (int pass = 0; pass < passes && !hit; ++pass) {
                        const int j = lane * passes + pass;
                        int st = HL_FREE;
                        unsigned amb = 0;
                        if (j < npts) {
                            float fx, fy, fc, fs;
                            rs_sample_world32(plan, j, inv_maxc, fx, fy, fc, fs);
                            if (fabsf(fx) > E.reach || fabsf(fy) > E.reach) st = far_status(flags, E.n_seg);
                            else if (!(fx == fx) || !(fy == fy) || !(fc == fc)) { st = HL_AMBIG; amb = flags; }
                            else st = filter_part(E, fx, fy, fc, fs, E.ext, flags, &amb);
                        }
                        hit = __any_sync(0xffffffffu, st == HL_HIT);
                        if (!hit && __any_sync(0xffffffffu, st == HL_AMBIG)) {
                            int bad = 0;
                            if (st == HL_AMBIG) {
                                double lx, ly, lyaw, wx, wy, wyaw;
                                int cs, dir;
                                rs_sample_local(plan, j, maxc, lx, ly, lyaw, cs, dir);
                                rs_to_world(q0, cq, sq, lx, ly, lyaw, wx, wy, wyaw);
                                Pose64 p64;
                                p64.x = wx; p64.y = wy; p64.c = m_cos(wyaw); p64.s = m_sin(wyaw);
                                bad = exact_part_check(p64, D.body_ext, eb, D, amb) ? 1 : 0;
                            }
                            hit = __any_sync(0xffffffffu, bad);
                        }
                    }
                } else {
                    // implement rectangles ride on every second pose (car_model.py:58): float64 samples
                    EnvSmem E0 = E;
                    E0.eps = D.eps;
                    for (int base = 0; base < npts && !hit; base += 32) {
                        const int j = base + lane;
                        int bad = 0;
                        if (j < npts) {
                            double lx, ly, lyaw, wx, wy, wyaw;
                            int cs, dir;
                            rs_sample_local(plan, j, maxc, lx, ly, lyaw, cs, dir);
                            rs_to_world(q0, cq, sq, lx, ly, lyaw, wx, wy, wyaw);
                            bad = pose_infeasible(eb, D, E0, wx, wy, wyaw, (j & 1) == 0, flags, nullptr) ? 1 : 0;
                        }
                        hit = __any_sync(0xffffffffu, bad);
                    }
                }
                if (lane == 0) out[k0 + r].collide = hit ? 1 : 0;
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rs_sample(const double* __restrict__ start, const HlRsWord* __restrict__ words, long long m, double maxc,
            double step, const long long* __restrict__ offset, double* __restrict__ ox, double* __restrict__ oy,
            double* __restrict__ oyaw, double* __restrict__ ocs, int8_t* __restrict__ odir) {
    __shared__ RsPlan plans[RS_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long n_warps = (long long)gridDim.x * RS_WARPS;
    for (long long i = (long long)blockIdx.x * RS_WARPS + wid; i < m; i += n_warps) {
        const HlRsWord w = words[i];
        double q0[3] = {start[3 * i], start[3 * i + 1], start[3 * i + 2]};
        __syncwarp();
        if (lane == 0) {
            double lens[HL_RS_MAX_SEGS];
            for (int s = 0; s < HL_RS_MAX_SEGS; ++s) lens[s] = w.nlen[s];
            rs_make_plan(w.cand, lens, maxc, xmul(step, maxc), plans[wid]);
        }
        __syncwarp();
        const RsPlan& P = plans[wid];
        const long long o = offset[i];
        const long long cap = offset[i + 1] - o;
        const double cq = cos(-q0[2]), sq = sin(-q0[2]);
        for (int j = lane; j < P.npts && j < cap; j += 32) {
            double lx, ly, lyaw, wx, wy, wyaw;
            int cs, dir;
            rs_sample_local(P, j, maxc, lx, ly, lyaw, cs, dir);
            rs_to_world(q0, cq, sq, lx, ly, lyaw, wx, wy, wyaw);
            ox[o + j] = wx; oy[o + j] = wy; oyaw[o + j] = wyaw;
            ocs[o + j] = cs == 0 ? 0.0 : (cs > 0 ? maxc : -maxc);
            odir[o + j] = (int8_t)dir;
        }
    }
}

extern "C" int hl_rs_all_paths(hl_ctx* ctx, const hl_env_batch* envs, const int32_t* d_env_id,
                               const double* d_start_goal, int64_t n, double maxc, double step,
                               double max_steer, uint32_t flags, HlRsWord* d_words, int32_t* d_count,
                               int32_t* d_order, void* stream) {
    if (!ctx || !d_start_goal || !d_words || !d_count || n < 0) { hl_set_error("hl_rs_all_paths: bad arguments"); return 1; }
    if (n == 0) return 0;
    if (hl_enter(ctx, envs, d_words, "hl_rs_all_paths")) return 1;
    long long blocks = (n + RS_WARPS - 1) / RS_WARPS;
    long long cap = (long long)ctx->sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    EnvBatchDev eb;
    memset(&eb, 0, sizeof(eb));
    if (envs) eb = envs->dev;
    k_rs_all_paths<<<grid, RS_WARPS * 32, sizeof(RsWarpSmem) * RS_WARPS, (cudaStream_t)stream>>>(
        eb, envs ? 1 : 0, d_env_id, d_start_goal, (long long)n, maxc, step, max_steer, flags, d_words,
        d_count, d_order);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int hl_rs_sample(hl_ctx* ctx, const double* d_start, const HlRsWord* d_words, int64_t m,
                            double maxc, double step, const int64_t* d_offset, double* d_x, double* d_y,
                            double* d_yaw, double* d_cs, int8_t* d_dir, void* stream) {
    if (!ctx || !d_start || !d_words || !d_offset || !d_x || !d_y || !d_yaw || !d_cs || !d_dir || m < 0) {
        hl_set_error("hl_rs_sample: bad arguments"); return 1;
    }
    if (m == 0) return 0;
    if (hl_enter(ctx, nullptr, d_x, "hl_rs_sample")) return 1;
    long long blocks = (m + RS_WARPS - 1) / RS_WARPS;
    long long cap = (long long)ctx->sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    k_rs_sample<<<grid, RS_WARPS * 32, 0, (cudaStream_t)stream>>>(
        d_start, d_words, (long long)m, maxc, step, (const long long*)d_offset, d_x, d_y, d_yaw, d_cs, d_dir);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}
