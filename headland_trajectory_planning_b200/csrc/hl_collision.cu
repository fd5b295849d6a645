// hl_collision.cu -- K1: per-pose footprint collision check.
//
// Replaces OrchardGeometryEnvironment.check_path_feasibility
// (orchard_geometry_environment.py:423-458) + CarModel.get_path_poly
// (car_model.py:39-73) and ReferenceLineHeuristic.check_path_feasibility
// (reference_line_heuristic.py:105-118), decomposed per pose.
//
// One thread per pose.  The CTA stages the float32 half-planes / polygon /
// lane segments of its tile's environment in shared memory (every thread reads the
// same obstacle at the same time -> broadcast), the float32 filter decides clear
// cases, and only poses inside the error band fall through to the float64
// predicates.  Roofline: FP32 ALU (12..24 B of HBM traffic per ~2.5 kflop check).
#include "hl_geom.cuh"

#ifndef K1_THREADS
#define K1_THREADS 256
#endif

// Ambiguous poses of a warp are resolved one at a time by the WHOLE warp (warp_exact_part_check): the pose
// is broadcast with shuffles and the exact float64 predicate runs on 32 lanes.
__device__ __forceinline__ void warp_resolve(bool need, double x, double y, double yaw, int env, const double* ext,
                                             unsigned amb, bool& bad, const EnvBatchDev& eb,
                                             unsigned long long* n_exact, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, need);
    if (m && n_exact && lane == 0) atomicAdd(n_exact, (unsigned long long)__popc(m));
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        Pose64 p;
        p.x = __shfl_sync(0xffffffffu, x, src);
        p.y = __shfl_sync(0xffffffffu, y, src);
        const double byaw = __shfl_sync(0xffffffffu, yaw, src);
        p.c = cos(byaw); p.s = sin(byaw);
        double e4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) e4[k] = __shfl_sync(0xffffffffu, ext[k], src);
        const int be = __shfl_sync(0xffffffffu, env, src);
        const unsigned ba = __shfl_sync(0xffffffffu, amb, src);
        const bool res = warp_exact_part_check(p, e4, eb, eb.desc[be], ba, lane);
        if (lane == src) bad = res;
    }
}

#ifndef K1_STAGED
#define K1_STAGED 0        // 1: three compacted filter passes per tile, 0: monolithic filter per pose.
                           // Measured (r1g): staged 9.1 G random / 9.9 G path-ordered checks/s vs monolithic 9.7 / 12.6 --
                           // four CTA barriers per 256-pose tile and sparse passes cost more than the divergence they
                           // remove at 4 CTAs per SM, so the monolithic filter stays the default.
#endif
#ifndef K1_MIN_CTAS
#define K1_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(K1_THREADS, K1_MIN_CTAS)
k_collision(EnvBatchDev eb, const int32_t* __restrict__ env_id, const double* __restrict__ poses,
            const int32_t* __restrict__ pose_idx, long long n, unsigned flags,
            uint8_t* __restrict__ out, unsigned long long* n_exact, int smem_floats) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31;
    const long long n_tiles = (n + K1_THREADS - 1) / K1_THREADS;
    int staged_env = -1;
    EnvSmem Es;
    bool staged = false;
    // software pipeline: the pose of the NEXT tile is loaded before this tile's arithmetic starts, so the
    // HBM latency of the (only) streaming input overlaps ~1.4k instructions of filter work
    double nx = 0.0, ny = 0.0, nyaw = 0.0;
    {
        const long long i0 = (long long)blockIdx.x * K1_THREADS + threadIdx.x;
        if (blockIdx.x < n_tiles && i0 < n) { nx = poses[3 * i0]; ny = poses[3 * i0 + 1]; nyaw = poses[3 * i0 + 2]; }
    }
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        long long base = tile * K1_THREADS;
        int e0 = env_id ? env_id[base] : 0;
        if (e0 != staged_env) {
            __syncthreads();
            stage_env(eb, eb.desc[e0], sm, smem_floats, Es, staged);
            staged_env = e0;
            __syncthreads();
        }
        const long long i = base + threadIdx.x;
        const bool active = i < n;
        const int e = active ? (env_id ? env_id[i] : 0) : e0;
        const EnvDesc& D = eb.desc[e];
        EnvSmem Eg;
        if (e != e0) global_env(eb, D, Eg);
        const EnvSmem& E = (e == e0) ? Es : Eg;
        const double x = nx, y = ny, yaw = nyaw;
        {
            const long long in = (tile + gridDim.x) * K1_THREADS + threadIdx.x;
            if (tile + gridDim.x < n_tiles && in < n) { nx = poses[3 * in]; ny = poses[3 * in + 1]; nyaw = poses[3 * in + 2]; }
        }
        bool with_aux = false;
        if (active) with_aux = pose_idx ? ((pose_idx[i] & 1) == 0) : true;
        const float px = (float)(x - D.origin[0]), py = (float)(y - D.origin[1]);
        const bool beyond = fabsf(px) > E.reach || fabsf(py) > E.reach;    // decided without a test (far_status)
        const bool far = !beyond && (!(fabs(yaw) < 1e6) || !(px == px) || !(py == py));
        float sf, cf;
        sincosf((float)yaw, &sf, &cf);
        // ---- body rectangle
        bool bad = false;
        unsigned amb = flags;
        int r = HL_FREE;
#if K1_STAGED
        // The lanes of a warp hold unrelated poses, so a monolithic filter makes every warp pay for every
        // section (obstacles, nearby field edges, lane, corners) as soon as ONE lane needs it.  The tile runs
        // the filter in three passes instead and COMPACTS the poses that still need the next pass, so the
        // rare sections execute on dense warps:  1) obstacles + field parity (all poses)  2) nearby field
        // edges (poses with a non-empty edge mask)  3) lane (poses that survived 1 and 2).
        __shared__ float s_px[K1_THREADS], s_py[K1_THREADS], s_c[K1_THREADS], s_s[K1_THREADS];
        __shared__ unsigned s_near[K1_THREADS];
        __shared__ unsigned char s_st[K1_THREADS], s_amb[K1_THREADS];     // st: 1 hit, 2 inside, 4 overflow, 8 lane ambiguous
        __shared__ unsigned short s_list2[K1_THREADS], s_list3[K1_THREADS];
        __shared__ int s_cnt[2];
        const int t = threadIdx.x;
        if (t < 2) s_cnt[t] = 0;
        __syncthreads();
        const bool staged_pose = active && !beyond && !far && e == e0;
        if (active && !staged_pose)
            r = beyond ? far_status(flags, E.n_seg) : (far ? HL_AMBIG : filter_part(E, px, py, cf, sf, E.ext, flags, &amb));
        if (staged_pose) {
            FiltState F;
            filt_stage1(Es, px, py, cf, sf, Es.ext, flags, F);
            bool need2 = false;
            if ((flags & HL_CHECK_BOUNDARY) && !F.hit) {
                if (F.near_mask == 0 && !F.overflow) { if (!F.inside) F.hit = true; }     // every edge clear: parity decides
                else need2 = true;
            }
            s_px[t] = px; s_py[t] = py; s_c[t] = cf; s_s[t] = sf;
            s_near[t] = F.near_mask;
            s_st[t] = (unsigned char)((F.hit ? 1 : 0) | (F.inside ? 2 : 0) | (F.overflow ? 4 : 0));
            s_amb[t] = (unsigned char)F.amb;
            if (need2) s_list2[atomicAdd(&s_cnt[0], 1)] = (unsigned short)t;
        }
        __syncthreads();
        if (t < s_cnt[0]) {
            const int q = s_list2[t];
            FiltState F;
            F.near_mask = s_near[q]; F.amb = s_amb[q];
            F.hit = false; F.inside = (s_st[q] & 2) != 0; F.overflow = (s_st[q] & 4) != 0;
            filt_field2(Es, s_px[q], s_py[q], s_c[q], s_s[q], Es.ext, F);
            s_st[q] = (unsigned char)((s_st[q] & ~1) | (F.hit ? 1 : 0));
            s_amb[q] = (unsigned char)F.amb;
        }
        __syncthreads();
        if (staged_pose && !(s_st[t] & 1) && (flags & HL_CHECK_LANE) && Es.n_seg > 0)
            s_list3[atomicAdd(&s_cnt[1], 1)] = (unsigned short)t;
        __syncthreads();
        if (t < s_cnt[1]) {
            const int q = s_list3[t];
            bool lane_amb;
            const int rl = filt_lane(Es, s_px[q], s_py[q], s_c[q], s_s[q], Es.ext, &lane_amb);
            if (rl == HL_HIT) s_st[q] |= 1;
            else if (lane_amb) s_amb[q] |= HL_CHECK_LANE;
        }
        __syncthreads();
        if (staged_pose) {
            if (s_st[t] & 1) r = HL_HIT;
            else if (s_amb[t]) { r = HL_AMBIG; amb = s_amb[t]; }
            else r = HL_FREE;
        }
#else
        if (active) r = beyond ? far_status(flags, E.n_seg) : (far ? HL_AMBIG : filter_part(E, px, py, cf, sf, E.ext, flags, &amb));
#endif
        if (r == HL_HIT) bad = true;
        warp_resolve(r == HL_AMBIG, x, y, yaw, e, D.body_ext, amb, bad, eb, n_exact, lane);
        // ---- implement rectangles: obstacles + field polygon, never the lane, poses 0,2,4,.. of a path
        // (orchard_geometry_environment.py:439-456; car_model.py:58)
        if (flags & HL_CHECK_AUX) {
            const unsigned aflags = flags & (HL_CHECK_OBSTACLES | HL_CHECK_BOUNDARY);
            int na = (active && with_aux && !beyond) ? D.n_aux : 0;
            int na_max = na;
            for (int o = 16; o; o >>= 1) na_max = max(na_max, __shfl_xor_sync(0xffffffffu, na_max, o));
            for (int a = 0; a < na_max; ++a) {
                const bool mine = a < na && !bad;
                const double* ext64 = eb.aux64 + 4 * (size_t)(D.aux_off + (a < D.n_aux ? a : 0));
                unsigned amb2 = aflags;
                int r2 = HL_FREE;
                if (mine) {
                    float ext32[4] = {(float)ext64[0], (float)ext64[1], (float)ext64[2], (float)ext64[3]};
                    r2 = far ? HL_AMBIG : filter_part(E, px, py, cf, sf, ext32, aflags, &amb2);
                    if (r2 == HL_HIT) bad = true;
                }
                bool bad2 = false;
                warp_resolve(r2 == HL_AMBIG, x, y, yaw, e, D.n_aux > 0 ? ext64 : D.body_ext, amb2, bad2, eb, n_exact, lane);
                bad = bad || bad2;
            }
        }
        if (active) out[i] = bad ? 1 : 0;
    }
}

__global__ void k_path_reduce(const uint8_t* __restrict__ pose_bad, const long long* __restrict__ path_start,
                              long long n_paths, uint8_t* __restrict__ path_bad) {
    // one warp per path: OR over its pose flags
    long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= n_paths) return;
    long long a = path_start[warp], b = path_start[warp + 1];
    int any = 0;
    for (long long i = a + lane; i < b; i += 32) any |= pose_bad[i];
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) path_bad[warp] = any ? 1 : 0;
}

extern "C" int hl_collision_check(hl_ctx* ctx, const hl_env_batch* envs, const int32_t* d_env_id,
                                  const double* d_poses, const int32_t* d_pose_idx, int64_t n,
                                  uint32_t flags, uint8_t* d_out, unsigned long long* d_n_exact,
                                  void* stream) {
    if (!ctx || !envs || !d_poses || !d_out || n < 0) { hl_set_error("hl_collision_check: bad arguments"); return 1; }
    if (n == 0) return 0;
    if (hl_enter(ctx, envs, d_out, "hl_collision_check")) return 1;
    const int smem_bytes = 32 * 1024;
    long long tiles = (n + K1_THREADS - 1) / K1_THREADS;
    int grid = (int)(tiles < (long long)ctx->sm_count * 8 ? tiles : (long long)ctx->sm_count * 8);
    k_collision<<<grid, K1_THREADS, smem_bytes, (cudaStream_t)stream>>>(
        envs->dev, d_env_id, d_poses, d_pose_idx, (long long)n, flags, d_out,
        d_n_exact, smem_bytes / 4);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int hl_path_reduce(hl_ctx* ctx, const uint8_t* d_pose_bad, const int64_t* d_path_start,
                              int64_t n_paths, uint8_t* d_path_bad, void* stream) {
    if (!ctx || !d_pose_bad || !d_path_start || !d_path_bad || n_paths < 0) { hl_set_error("hl_path_reduce: bad arguments"); return 1; }
    if (n_paths == 0) return 0;
    if (hl_enter(ctx, nullptr, d_path_bad, "hl_path_reduce")) return 1;
    long long threads = n_paths * 32;
    int grid = (int)((threads + 255) / 256);
    k_path_reduce<<<grid, 256, 0, (cudaStream_t)stream>>>(d_pose_bad, (const long long*)d_path_start, n_paths, d_path_bad);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}
