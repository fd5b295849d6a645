// hl_refpath.cu -- K8: warm-start path -> OBCA initial guess.
//
// Replaces get_init_ref_path (obca_py/util.py:62-113) with calc_spline_course / Spline2D
// (path_planner/utils/cubic_spline.py:19-112) for a whole sweep: every planner path is split where the driving
// direction changes, each piece gets scipy's not-a-knot cubic splines x(s), y(s) over its chord length, is
// resampled at ds and emitted as rows (x, y, v, yaw, steer); the yaw column is then unwrapped along the path.
// Input is exactly the pooled path output of hl_hybrid_astar_batch.
//
// Two launches: k_refpath_count (thread per path: pieces, chord lengths, sample counts -> the caller's prefix sum)
// and k_refpath_fill (warp per path: lane 0 assembles the knots and solves the tridiagonal systems -- the Thomas
// recurrence is sequential --, all lanes evaluate the samples, lane 0 unwraps the yaw).  float64 throughout.
#include "hl_spline.cuh"

namespace {

#define RP_WS 8                         // workspace doubles per input pose

__device__ __forceinline__ double rp_wrap(double a) {          // obca_py/util.py:7-13, Python's floored %
    return xsub(py_mod_pos(xadd(a, HL_PI), 2.0 * HL_PI), HL_PI);
}

__global__ void k_refpath_count(const double* __restrict__ px, const double* __restrict__ py,
                                const int8_t* __restrict__ pdir, const long long* __restrict__ io, long long n, double ds,
                                long long* __restrict__ counts, int* __restrict__ status) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const long long a = io[p], b = io[p + 1];
        long long total = 0;
        int st = (b - a) < 1 ? 1 : 0;
        long long seg = a;
        while (seg < b && !st) {
            long long end = seg + 1;
            while (end < b && pdir[end] == pdir[end - 1]) ++end;
            // chord length of the piece, repeated poses dropped
            int m = 0;
            double s = 0.0, lx = 0.0, ly = 0.0;
            for (long long i = seg; i < end; ++i) {
                if (i + 1 < end && px[i + 1] == px[i] && py[i + 1] == py[i]) continue;
                if (m > 0) s = xadd(s, hypot_cr(xsub(px[i], lx), xsub(py[i], ly)));
                lx = px[i]; ly = py[i];
                ++m;
            }
            if (m < 2) st = 1;                        // scipy: "`x` must contain at least 2 elements"
            else total += rp_count(s, ds);
            seg = end;
        }
        counts[p] = st ? 0 : total;
        status[p] = st;
    }
}

__global__ void __launch_bounds__(128) k_refpath_fill(const double* __restrict__ px, const double* __restrict__ py,
                                                      const int8_t* __restrict__ pdir, const long long* __restrict__ io,
                                                      const long long* __restrict__ oo, const int* __restrict__ status,
                                                      long long n, double wheel_base, double desired_v, double ds,
                                                      double* __restrict__ ws, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < n; p += n_warps) {
        if (status[p]) continue;
        const long long a = io[p], b = io[p + 1];
        const long long cap = b - a;
        double* kx = ws + RP_WS * a;
        double* ky = kx + cap; double* ks = ky + cap; double* dx = ks + cap; double* dy = dx + cap;
        double* cp = dy + cap; double* bx = cp + cap; double* by = bx + cap;
        double* o = out + 5 * oo[p];
        long long row = 0;
        long long seg = a;
        while (seg < b) {
            long long end = seg + 1;
            while (end < b && pdir[end] == pdir[end - 1]) ++end;
            int m = 0;
            if (lane == 0) {
                m = rp_knots(px, py, seg, end, kx, ky, ks);
                rp_derivs(ks, kx, ky, m, dx, dy, cp, bx, by);
            }
            m = __shfl_sync(0xffffffffu, m, 0);
            __syncwarp();
            const long long cnt = rp_count(ks[m - 1], ds);
            const double dir0 = (double)pdir[seg], dirl = (double)pdir[end - 1];
            const bool reversed = dirl < 0.0;                  // util.py:93-97
            for (long long i = lane; i < cnt; i += 32) {
                const double t = xmul((double)i, ds);           // np.arange(0, s_end + ds, ds)[i]
                double x, x1, x2, y, y1, y2;
                rp_eval(ks, kx, dx, m, t, x, x1, x2);
                rp_eval(ks, ky, dy, m, t, y, y1, y2);
                double yaw = m_atan2(y1, x1);
                const double q = x1 * x1 + y1 * y1;
                const double k = (y2 * x1 - x2 * y1) / (q * sqrt(q));
                if (reversed) yaw = rp_wrap(yaw + HL_PI);
                double steer = atan(wheel_base * k) * (reversed ? -1.0 : 1.0);
                double v = dir0 * desired_v;
                if (i == 0) { v = 0.0; steer = 0.0; }
                double* r = o + 5 * (row + i);
                r[0] = x; r[1] = y; r[2] = v; r[3] = yaw; r[4] = steer;
            }
            row += cnt;
            seg = end;
            __syncwarp();
        }
        __syncwarp();
        if (lane == 0 && row > 0) {                             // process_angle (util.py:16-44) + end speeds
            double prev_raw = rp_wrap(o[3]);
            double acc = prev_raw;
            o[3] = acc;
            for (long long i = 1; i < row; ++i) {
                const double raw = rp_wrap(o[5 * i + 3]);
                acc = acc + rp_wrap(raw - prev_raw);
                prev_raw = raw;
                o[5 * i + 3] = acc;
            }
            o[2] = 0.0;
            o[5 * (row - 1) + 2] = 0.0;
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int hl_ref_path_count(hl_ctx* ctx, const double* d_x, const double* d_y, const int8_t* d_dir,
                                 const int64_t* d_in_offsets, int64_t n_paths, double ds, int64_t* d_counts,
                                 int32_t* d_status, void* stream) {
    if (!ctx || !d_x || !d_y || !d_dir || !d_in_offsets || !d_counts || !d_status || n_paths < 0 || !(ds > 0.0)) {
        hl_set_error("hl_ref_path_count: bad arguments"); return 1;
    }
    if (n_paths == 0) return 0;
    if (hl_enter(ctx, nullptr, d_counts, "hl_ref_path_count")) return 1;
    const int grid = (int)((n_paths + 127) / 128 < 1184 ? (n_paths + 127) / 128 : 1184);
    k_refpath_count<<<grid, 128, 0, (cudaStream_t)stream>>>(d_x, d_y, d_dir, (const long long*)d_in_offsets, (long long)n_paths,
                                                            ds, (long long*)d_counts, d_status);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int hl_ref_path_fill(hl_ctx* ctx, const double* d_x, const double* d_y, const int8_t* d_dir,
                                const int64_t* d_in_offsets, const int64_t* d_out_offsets, const int32_t* d_status,
                                int64_t n_paths, double wheel_base, double desired_v, double ds, double* d_workspace,
                                double* d_out, void* stream) {
    if (!ctx || !d_x || !d_y || !d_dir || !d_in_offsets || !d_out_offsets || !d_status || !d_workspace || !d_out ||
        n_paths < 0 || !(ds > 0.0)) {
        hl_set_error("hl_ref_path_fill: bad arguments"); return 1;
    }
    if (n_paths == 0) return 0;
    if (hl_enter(ctx, nullptr, d_out, "hl_ref_path_fill")) return 1;
    long long want = (n_paths + 3) / 4;
    const long long cap = (long long)ctx->sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    k_refpath_fill<<<grid, 128, 0, (cudaStream_t)stream>>>(d_x, d_y, d_dir, (const long long*)d_in_offsets,
                                                           (const long long*)d_out_offsets, d_status, (long long)n_paths,
                                                           wheel_base, desired_v, ds, d_workspace, d_out);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}
