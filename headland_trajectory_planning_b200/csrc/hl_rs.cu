// hl_rs.cu -- K2/K3: batched Reeds-Shepp word evaluation, sampling and shots.
//
// One warp per start/goal pair: lanes evaluate the 46 candidate rows, lane 0
// replays set_path's order-dependent dedup and the heapdict pop order, then one
// lane per accepted word builds its sampling plan and (when an environment is
// given) the whole warp samples each word and collision-checks it with ballot
// early exit.  Replaces reeds_shepp.calc_all_paths (reeds_shepp.py:39-65) and the
// per-path loop of _get_goal_extension_with_reeds_shepp_path
// (hybrid_a_star_search.py:249-287).
#include <cstring>
#include "hl_geom.cuh"
#include "hl_rs.cuh"

#define RS_WARPS 4

struct RsWarpSmem {
    double lens[HL_RS_CANDIDATES][HL_RS_MAX_SEGS];
    double L[HL_RS_CANDIDATES];
    double prio[HL_RS_CANDIDATES];
    int acc[HL_RS_CANDIDATES];
    int order[HL_RS_CANDIDATES];
    unsigned char valid[HL_RS_CANDIDATES + 2];
    int n;
    RsPlan plan;
};

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rs_all_paths(EnvBatchDev eb, int have_env, const int32_t* __restrict__ env_id,
               const double* __restrict__ sg, long long n, double maxc, double step, double max_steer,
               unsigned flags, HlRsWord* __restrict__ words, int32_t* __restrict__ count,
               int32_t* __restrict__ order_out) {
    __shared__ RsWarpSmem sm[RS_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    RsWarpSmem& W = sm[wid];
    const long long n_warps = (long long)gridDim.x * RS_WARPS;
    for (long long i = (long long)blockIdx.x * RS_WARPS + wid; i < n; i += n_warps) {
        double q0[3] = {sg[6 * i], sg[6 * i + 1], sg[6 * i + 2]};
        double q1[3] = {sg[6 * i + 3], sg[6 * i + 4], sg[6 * i + 5]};
        RsProblem P = rs_normalise(q0, q1, maxc);
        for (int c = lane; c < HL_RS_CANDIDATES; c += 32) {
            double l[HL_RS_MAX_SEGS] = {0, 0, 0, 0, 0};
            bool ok = rs_candidate(c, P, l);
            W.valid[c] = ok ? 1 : 0;
            for (int k = 0; k < HL_RS_MAX_SEGS; ++k) W.lens[c][k] = l[k];
        }
        __syncwarp();
        if (lane == 0) {
            int m = rs_select(W.valid, W.lens, W.acc, W.L);
            W.n = m;
            for (int k = 0; k < m; ++k)
                W.prio[k] = rs_path_cost(0.0, W.acc[k], W.lens[W.acc[k]], max_steer, 5000.0, 1000.0, 1.0);
            if (m > 0) heapdict_order(W.prio, m, W.order);
            count[i] = m;
        }
        __syncwarp();
        const int m = W.n;
        HlRsWord* out = words + (size_t)i * HL_RS_CANDIDATES;
        for (int k = lane; k < m; k += 32) {
            const int c = W.acc[k];
            RsPlan plan;
            rs_make_plan(c, W.lens[c], maxc, xmul(step, maxc), plan);
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
        if (order_out)
            for (int k = lane; k < HL_RS_CANDIDATES; k += 32)
                order_out[(size_t)i * HL_RS_CANDIDATES + k] = (k < m) ? W.order[k] : -1;
        if (have_env && m > 0) {
            const int e = env_id ? env_id[i] : 0;
            const EnvDesc& D = eb.desc[e];
            EnvSmem E;
            global_env(eb, D, E);
            const double cq = cos(-q0[2]), sq = sin(-q0[2]);
            for (int k = 0; k < m; ++k) {
                __syncwarp();
                if (lane == 0) rs_make_plan(W.acc[k], W.lens[W.acc[k]], maxc, xmul(step, maxc), W.plan);
                __syncwarp();
                const int npts = W.plan.npts;
                int hit = 0;
                for (int base = 0; base < npts && !hit; base += 32) {
                    int j = base + lane;
                    int bad = 0;
                    if (j < npts) {
                        double lx, ly, lyaw, wx, wy, wyaw;
                        int cs, dir;
                        rs_sample_local(W.plan, j, maxc, lx, ly, lyaw, cs, dir);
                        rs_to_world(q0, cq, sq, lx, ly, lyaw, wx, wy, wyaw);
                        bad = pose_infeasible(eb, D, E, wx, wy, wyaw, (j & 1) == 0, flags, nullptr) ? 1 : 0;
                    }
                    hit = __any_sync(0xffffffffu, bad);
                }
                if (lane == 0) out[k].collide = hit ? 1 : 0;
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
    HL_CUDA_OK(cudaSetDevice(ctx->device));
    long long blocks = (n + RS_WARPS - 1) / RS_WARPS;
    long long cap = (long long)ctx->sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    EnvBatchDev eb;
    memset(&eb, 0, sizeof(eb));
    if (envs) eb = envs->dev;
    k_rs_all_paths<<<grid, RS_WARPS * 32, 0, (cudaStream_t)stream>>>(
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
    HL_CUDA_OK(cudaSetDevice(ctx->device));
    long long blocks = (m + RS_WARPS - 1) / RS_WARPS;
    long long cap = (long long)ctx->sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    k_rs_sample<<<grid, RS_WARPS * 32, 0, (cudaStream_t)stream>>>(
        d_start, d_words, (long long)m, maxc, step, (const long long*)d_offset, d_x, d_y, d_yaw, d_cs, d_dir);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}
