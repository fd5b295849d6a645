// hl_ypark.cu -- K5: candidate paths of the Y-type parking sweep.
//
// Replaces the body of the 4-deep loop of search_y_type_parking_path
// (path_planner/headland_path_planning.py:405-427): get_y_type_parking_path (:488-516) =
// two calculate_motion_path rollouts (:455-485) planned inversely from the end pose, then
// get_path_in_odom (:519-527).  All candidates of a sweep (and of many sweeps) are generated in
// ONE launch, checked by ONE k_collision launch and reduced by k_path_reduce; the first feasible
// candidate in loop order is the reference's answer.
//
// One warp per candidate.  The per-step terms step*cos(yaw)*dir are computed lane-parallel; the
// running sums are then added by lane 0 in sequence, exactly like np.cumsum, so the poses agree
// with numpy to the last bit wherever CUDA's cos/sin/tan agree with libm.
#include "hl_common.cuh"
#include "hl_crmath.cuh"

#define YP_MAX_STEPS 512        // per rollout (5 m at 0.01 m); longer candidates are rejected by the host side

namespace {

// np.linspace(start, start + yaw_step*div, div + 1)[i] followed by angle_wrap
__device__ __forceinline__ double yp_yaw(double init_yaw, double stop, double lstep, double delta, int div, int i) {
    double v;
    if (div > 0 && i == div) v = stop;
    else if (lstep != 0.0) v = xadd(xmul((double)i, lstep), init_yaw);
    else v = xadd(xmul(xdiv((double)i, (double)(div > 0 ? div : 1)), delta), init_yaw);
    return angle_wrap(v);
}

// One calculate_motion_path: poses 1..n of the rollout into (xs, ys, yw) [n]; returns the last pose.
// stop_mult: the linspace end is init_yaw + yaw_step * stop_mult (n in headland_path_planning.py:466, n + 1 in
// CarModel.calculate_motion_path, car_model.py:213-215)
__device__ void yp_rollout(double ix, double iy, double iyaw, double steer, double dir, int n, int stop_mult,
                           double wheel_base, double step, double* tx, double* ty, double* yw, int lane) {
    const double yaw_step = xmul(xdiv(xmul(dir, step), wheel_base), m_tan(steer));
    const double init_yaw = angle_wrap(xadd(iyaw, yaw_step));
    const double stop = xadd(init_yaw, xmul(yaw_step, (double)stop_mult));
    const double delta = xsub(stop, init_yaw);
    const double lstep = n > 0 ? xdiv(delta, (double)n) : 0.0;
    for (int i = lane; i < n; i += 32) {
        const double y0 = yp_yaw(init_yaw, stop, lstep, delta, n, i);        // yaws[:-1][i]
        double sn, cs;
        cr_sincos(y0, &sn, &cs);       // correctly rounded like numpy's / glibc's (hl_crmath.cuh): these kernels are tiny
        tx[i] = xmul(xmul(step, cs), dir);
        ty[i] = xmul(xmul(step, sn), dir);
        yw[i] = yp_yaw(init_yaw, stop, lstep, delta, n, i + 1);              // yaws[1:][i]
    }
    __syncwarp();
    if (lane == 0) {
        double ax = 0.0, ay = 0.0;
        for (int i = 0; i < n; ++i) {
            ax = (i == 0) ? tx[0] : xadd(ax, tx[i]);
            ay = (i == 0) ? ty[0] : xadd(ay, ty[i]);
            tx[i] = xadd(ix, ax);
            ty[i] = xadd(iy, ay);
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(128) k_ypark_paths(const double* __restrict__ cand, const long long* __restrict__ offs,
                                                     long long n, double step, double* __restrict__ poses) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* bx = sm + (size_t)wib * 6 * YP_MAX_STEPS;
    double* by = bx + YP_MAX_STEPS; double* bw = by + YP_MAX_STEPS;
    double* fx = bw + YP_MAX_STEPS; double* fy = fx + YP_MAX_STEPS; double* fw = fy + YP_MAX_STEPS;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = warp; c < n; c += n_warps) {
        const double* q = cand + 8 * c;
        const double bl = q[0], fl = q[1], sb = q[2], sf = q[3], ex = q[4], ey = q[5], eyaw = q[6], wb = q[7];
        const int nb = (int)rint(xdiv(bl, step)), nf = (int)rint(xdiv(fl, step));
        double* out = poses + 3 * offs[c];
        const long long room = offs[c + 1] - offs[c];
        if (nb < 0 || nf < 0 || nb > YP_MAX_STEPS || nf > YP_MAX_STEPS || room != (long long)nb + nf + 2) {
            // host and device disagree on the pose count: poison the candidate (NaN poses are infeasible)
            for (long long i = lane; i < 3 * room; i += 32) out[i] = __longlong_as_double(0x7ff8000000000000LL);
            continue;
        }
        yp_rollout(0.0, 0.0, 0.0, sb, -1.0, nb, nb, wb, step, bx, by, bw, lane);
        const double jx = nb ? bx[nb - 1] : 0.0, jy = nb ? by[nb - 1] : 0.0, jw = nb ? bw[nb - 1] : 0.0;
        yp_rollout(jx, jy, jw, sf, 1.0, nf, nf, wb, step, fx, fy, fw, lane);
        // odom frame (transformation.py:7-61, navigation_utils.py:196-203)
        double se, ce;
        cr_sincos(eyaw, &se, &ce);
        const double dyaw = m_atan2(se, ce);
        const int total = nb + nf + 2;
        for (int r = lane; r < total; r += 32) {
            // rows: forward path reversed (nf .. 0), then backward path reversed (nb .. 0); index 0 = rollout start
            double lx, ly, lw;
            if (r <= nf) {
                const int k = nf - r;                 // k-th pose of the forward rollout, 0 = its init pose
                if (k == 0) { lx = jx; ly = jy; lw = jw; } else { lx = fx[k - 1]; ly = fy[k - 1]; lw = fw[k - 1]; }
            } else {
                const int k = nb - (r - nf - 1);
                if (k == 0) { lx = 0.0; ly = 0.0; lw = 0.0; } else { lx = bx[k - 1]; ly = by[k - 1]; lw = bw[k - 1]; }
            }
            out[3 * r] = xadd(xadd(xmul(ce, lx), xmul(-se, ly)), ex);
            out[3 * r + 1] = xadd(xadd(xmul(se, lx), xmul(ce, ly)), ey);
            out[3 * r + 2] = xadd(lw, dyaw);
        }
        __syncwarp();
    }
}

// CarModel.calculate_motion_path (car_model.py:202-234) for many (init pose, command) rows: [init pose; n poses].
__global__ void __launch_bounds__(128) k_arc_paths(const double* __restrict__ cand, const long long* __restrict__ offs,
                                                   long long n, double step, double* __restrict__ poses) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* ax = sm + (size_t)wib * 3 * YP_MAX_STEPS;
    double* ay = ax + YP_MAX_STEPS; double* aw = ay + YP_MAX_STEPS;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = warp; c < n; c += n_warps) {
        const double* q = cand + 8 * c;
        const double ix = q[0], iy = q[1], iyaw = q[2], steer = q[3], dir = q[4], len = q[5], wb = q[6];
        const int ns = (int)rint(xdiv(len, step));
        double* out = poses + 3 * offs[c];
        const long long room = offs[c + 1] - offs[c];
        if (ns < 1 || ns > YP_MAX_STEPS || room != (long long)ns + 1) {
            for (long long i = lane; i < 3 * room; i += 32) out[i] = __longlong_as_double(0x7ff8000000000000LL);
            continue;
        }
        yp_rollout(ix, iy, iyaw, steer, dir, ns, ns + 1, wb, step, ax, ay, aw, lane);
        if (lane == 0) { out[0] = ix; out[1] = iy; out[2] = iyaw; }
        for (int r = lane; r < ns; r += 32) { out[3 * (r + 1)] = ax[r]; out[3 * (r + 1) + 1] = ay[r]; out[3 * (r + 1) + 2] = aw[r]; }
        __syncwarp();
    }
}

}  // namespace

extern "C" int hl_arc_paths(hl_ctx* ctx, const double* d_cand, const int64_t* d_offsets, int64_t n, double step,
                            double* d_poses, void* stream) {
    if (!ctx || !d_cand || !d_offsets || !d_poses || n < 0 || !(step > 0.0)) {
        hl_set_error("hl_arc_paths: bad arguments"); return 1;
    }
    if (n == 0) return 0;
    if (hl_enter(ctx, nullptr, d_poses, "hl_arc_paths")) return 1;
    const int threads = 128;
    const size_t smem = (size_t)(threads / 32) * 3 * YP_MAX_STEPS * sizeof(double);      // 48 KB
    HL_CUDA_OK(cudaFuncSetAttribute(k_arc_paths, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long want = (n + 3) / 4;
    const long long cap = (long long)ctx->sm_count * 4;
    const int grid = (int)(want < cap ? want : cap);
    k_arc_paths<<<grid, threads, smem, (cudaStream_t)stream>>>(d_cand, (const long long*)d_offsets, (long long)n, step, d_poses);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int hl_ypark_paths(hl_ctx* ctx, const double* d_cand, const int64_t* d_offsets, int64_t n, double step,
                              double* d_poses, void* stream) {
    if (!ctx || !d_cand || !d_offsets || !d_poses || n < 0 || !(step > 0.0)) {
        hl_set_error("hl_ypark_paths: bad arguments"); return 1;
    }
    if (n == 0) return 0;
    if (hl_enter(ctx, nullptr, d_poses, "hl_ypark_paths")) return 1;
    const int threads = 128;
    const size_t smem = (size_t)(threads / 32) * 6 * YP_MAX_STEPS * sizeof(double);      // 96 KB
    HL_CUDA_OK(cudaFuncSetAttribute(k_ypark_paths, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long want = (n + 3) / 4;
    const long long cap = (long long)ctx->sm_count * 2;
    const int grid = (int)(want < cap ? want : cap);
    k_ypark_paths<<<grid, threads, smem, (cudaStream_t)stream>>>(d_cand, (const long long*)d_offsets, (long long)n, step, d_poses);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}
