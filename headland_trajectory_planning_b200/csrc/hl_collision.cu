// hl_collision.cu -- K1: per-pose footprint collision check.
//
// Replaces OrchardGeometryEnvironment.check_path_feasibility
// (orchard_geometry_environment.py:423-458) + CarModel.get_path_poly
// (car_model.py:39-73) and ReferenceLineHeuristic.check_path_feasibility
// (reference_line_heuristic.py:105-118), decomposed per pose.
//
// Round-2 structure:
//   * a CTA of 8 warps walks 1024-pose tiles; every warp owns a 128-pose slice (4 poses per lane);
//   * the slice's poses (3 KB, contiguous) are fetched by ONE cp.async.bulk (TMA, 1-D) into the warp's landing zone,
//     completion on the warp's mbarrier; the copy of the NEXT slice is issued as soon as the current one has been
//     converted to the float32 frame, so HBM latency hides behind the filter stages;
//   * the environment's float32 records are staged per CTA by three bulk copies on a CTA mbarrier and read with
//     ld.shared.v4 (LDS, broadcast) -- never through a generic pointer;
//   * the float32 filter runs as STAGES over the slice, cheapest and most decisive first -- lane centre test,
//     obstacles (box-box SAT), field-edge pass, nearby field edges, lane corners -- and between stages the warp
//     COMPACTS the poses that are still undecided (ballot + popc into a byte list), so the expensive stages run on
//     dense warps instead of dragging 32 unrelated poses through every section;
//   * poses inside the float32 error band are resolved by the whole warp with the float64 predicates (hl_geom.cuh).
// Slices that cannot use this path (environment larger than the staging area, several environments in one slice,
// non-rectangular obstacle quads) take the monolithic per-pose filter on global memory (k1_slow_slice).
// Roofline: FP32 ALU (24 B of HBM traffic per ~2.5 kflop check).
#include "hl_geom.cuh"

#define K1_WARPS 8
#define K1_THREADS (K1_WARPS * 32)
#define K1_TILE 128                       // poses per warp slice
#define K1_PER_LANE (K1_TILE / 32)
#define K1_CTA_TILE (K1_WARPS * K1_TILE)
#define K1_PAIRS 256                      // capacity of a warp's (pose, obstacle) / (pose, field edge) pair list
#define K1_ENV_FLOATS 1024                // staged float32 records (canonical environment: 432 floats)
#ifndef K1_MIN_CTAS
#define K1_MIN_CTAS 4
#endif

// ------------------------------------------------------------------ PTX: mbarrier + 1-D bulk copy (TMA)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// generic-proxy reads of a buffer must be ordered before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte load of a staged record.  The pointers below all derive from the kernel's `extern __shared__` block, so
// the compiler proves the address space and emits LDS.128 (checked in the SASS, profiles/r2_k1_sass.txt).
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ------------------------------------------------------------------ shared-memory layout
struct __align__(16) K1Warp {
    double raw[K1_TILE * 3];              // TMA landing zone: x, y, yaw of the slice
    float4 pose[K1_TILE];                 // px, py (relative to the environment origin), cos, sin
    unsigned short pairs[K1_PAIRS];       // (pose q | record k << 8) work items of the obstacle / field-edge passes
    unsigned char st[K1_TILE];            // K1S_* bits
    unsigned char amb[K1_TILE];           // HL_CHECK_* bits inside the float32 band
    unsigned char la[K1_TILE], lb[K1_TILE], lc[K1_TILE];     // compacted pose lists
    unsigned long long bar;               // mbarrier of the landing zone
};
struct __align__(16) K1Cta {
    float env[K1_ENV_FLOATS];             // obstacle records | field-edge records | lane segments
    float4 segd[HL_MAX_SEGS];             // per lane segment: ex, ey, 1/len^2, 1/len (derived once per staging)
    float4 fld[32][2];                    // per field edge, for the first pass: (Ax, Ay, Ex, Ey) with the endpoints
                                          // ordered so that Ey >= 0, and (nx, ny, c, By) -- see k1_field1
    unsigned long long bar;
    K1Warp w[K1_WARPS];
};
enum { K1S_HIT = 1, K1S_INSIDE = 2, K1S_NOTCLEAR = 4, K1S_CORNERS = 8, K1S_DONE = 16, K1S_FAR = 32, K1S_OFF = 64 };

struct K1Env {                            // staged environment (shared memory, uniform per CTA)
    const float* obs_a; const float* field_a; const float* seg_a; const float* segd_a; const float* fld_a;
    int n_obs, n_field, n_seg;
    float eps;
};

// Ambiguous poses of a warp are resolved one at a time by the WHOLE warp (warp_exact_part_check): the pose
// is broadcast with shuffles and the exact float64 predicate runs on 32 lanes.
__device__ __noinline__ void warp_resolve(bool need, double x, double y, double yaw, int env, const double* ext,
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

// OR `bit` into byte q of a shared byte array (several lanes may hold pairs of the same pose)
__device__ __forceinline__ void st_or(unsigned char* arr, int q, unsigned bit) {
    atomicOr(reinterpret_cast<unsigned*>(arr) + (q >> 2), bit << (8 * (q & 3)));
}

// appends q to `list` for every lane with `pred`; returns the new (warp-uniform) count
__device__ __forceinline__ int warp_append(unsigned char* list, int cnt, bool pred, int q, int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (pred) list[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned char)q;
    return cnt + __popc(m);
}

// ------------------------------------------------------------------ filter stages (float32, staged environment)
// Same formulations and the same conservative band as filt_stage1 / filt_field2 / filt_lane (hl_geom.cuh), split so
// that every stage is a short uniform loop over shared-memory records.
struct K1Rect { float hx, hy, mx, my, x0, x1, y0, y1; };
__device__ __forceinline__ K1Rect k1_rect(const float* ext) {
    K1Rect R;
    R.hx = 0.5f * (ext[1] - ext[0]); R.hy = 0.5f * (ext[3] - ext[2]);
    R.mx = 0.5f * (ext[1] + ext[0]); R.my = 0.5f * (ext[3] + ext[2]);
    R.x0 = ext[0]; R.x1 = ext[1]; R.y0 = ext[2]; R.y1 = ext[3];
    return R;
}

// Stage A: rectangle centre against every capsule axis.  0 = the whole rectangle is inside one capsule,
// 1 = every point of it is outside every capsule (HIT), 2 = the corners decide.
__device__ __forceinline__ int k1_lane_centre(const K1Env& E, const K1Rect& R, float4 P, float rho) {
    const float Cx = fmaf(P.z, R.mx, fmaf(-P.w, R.my, P.x)), Cy = fmaf(P.w, R.mx, fmaf(P.z, R.my, P.y));
    const float rin = (float)HL_LANE_RIN - E.eps, rout = (float)HL_LANE_R + E.eps;
    bool accepted = false, need = false;
    for (int i = 0; i < E.n_seg; ++i) {
        const float4 sg = lds4(E.seg_a + 4 * i), sd = lds4(E.segd_a + 4 * i);
        const float qx = Cx - sg.x, qy = Cy - sg.y;
        const float t = fminf(fmaxf(fmaf(qx, sd.x, qy * sd.y) * sd.z, 0.f), 1.f);
        const float ddx = fmaf(-t, sd.x, qx), ddy = fmaf(-t, sd.y, qy);
        const float d = f_sqrt(fmaf(ddx, ddx, ddy * ddy));
        accepted |= d + rho <= rin;
        need |= d - rho <= rout;
    }
    return accepted ? 0 : (need ? 2 : 1);
}

// Stage B, first pass: which obstacles can the rectangle touch at all?  Centre of the rectangle against the
// obstacle box grown by the circumradius (+ band), in the obstacle's own frame: ~12 instructions per obstacle.  A box
// that fails this is separated from the rectangle by more than eps along one of its own axes.
__device__ __forceinline__ unsigned k1_obstacle_mask(const K1Env& E, const K1Rect& R, float4 P, float rho_eps) {
    const float c = P.z, s = P.w;
    const float Cx = fmaf(c, R.mx, fmaf(-s, R.my, P.x)), Cy = fmaf(s, R.mx, fmaf(c, R.my, P.y));
    unsigned m = 0;
    for (int k = 0; k < E.n_obs; ++k) {
        const float4 b0 = lds4(E.obs_a + HL_OBS32_STRIDE * k + 20);      // flag, cx, cy, ax
        const float4 b1 = lds4(E.obs_a + HL_OBS32_STRIDE * k + 24);      // ay, ha, hb, pad
        const float dx = b0.y - Cx, dy = b0.z - Cy;
        const float da = fabsf(fmaf(dx, b0.w, dy * b1.x)), db = fabsf(fmaf(dy, b0.w, -dx * b1.x));
        m |= (unsigned)(!(da > b1.y + rho_eps) && !(db > b1.z + rho_eps)) << k;
    }
    return m;
}

// Stage B, second pass: one (pose, obstacle) pair -- 4-axis box-box separating test on (centre, axis, half extents).
__device__ __forceinline__ void k1_obstacle_one(const K1Env& E, const K1Rect& R, float4 P, int k, bool& hit, bool& amb) {
    const float c = P.z, s = P.w;
    const float Cx = fmaf(c, R.mx, fmaf(-s, R.my, P.x)), Cy = fmaf(s, R.mx, fmaf(c, R.my, P.y));
    const float4 b0 = lds4(E.obs_a + HL_OBS32_STRIDE * k + 20);      // flag, cx, cy, ax
    const float4 b1 = lds4(E.obs_a + HL_OBS32_STRIDE * k + 24);      // ay, ha, hb, pad
    const float dx = b0.y - Cx, dy = b0.z - Cy;
    const float ax = b0.w, ay = b1.x, ha = b1.y, hb = b1.z;
    const float p = fabsf(fmaf(ax, c, ay * s)), q = fabsf(fmaf(ay, c, -ax * s));
    const float du = fabsf(fmaf(dx, c, dy * s)), dv = fabsf(fmaf(dy, c, -dx * s));
    const float da = fabsf(fmaf(dx, ax, dy * ay)), db = fabsf(fmaf(dy, ax, -dx * ay));
    const float sep = fmaxf(fmaxf(du - fmaf(ha, p, fmaf(hb, q, R.hx)), dv - fmaf(ha, q, fmaf(hb, p, R.hy))),
                            fmaxf(da - fmaf(R.hx, p, fmaf(R.hy, q, ha)), db - fmaf(R.hx, q, fmaf(R.hy, p, hb))));
    hit = sep < -E.eps;
    amb = fabsf(sep) <= E.eps;
}

// Appends this lane's pairs (q | k << 8 for every set bit k of `m`) to the warp's pair list, flushing the list through
// `process` first when it would overflow.  Returns false when this chunk alone exceeds the capacity (caller falls
// back to a per-pose loop).
template <class F>
__device__ __forceinline__ bool k1_push_pairs(K1Warp& W, int& n_pairs, unsigned m, int q, int lane, F&& process) {
    int cnt = __popc(m), off = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += t; }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    if (total > K1_PAIRS) return false;
    if (n_pairs + total > K1_PAIRS) { process(n_pairs); n_pairs = 0; __syncwarp(); }
    off += n_pairs - cnt;
    while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1;
        W.pairs[off++] = (unsigned short)(q | (k << 8));
    }
    n_pairs += total;
    return true;
}

// Stage C: one pass over the field edges -- crossing parity of the centre + mask of the edges whose LINE passes
// within the circumradius of the rectangle (the support-radius refinement of those few edges is stage D's first
// test).  Branch-free: 2 LDS.128 + ~17 ALU instructions per edge.  The derived records have the endpoints ordered
// by y (Ey >= 0), so the crossing rule is one comparison; the straddle test uses the very same float values for a
// vertex in both of its edges, which keeps the half-open rule consistent.
__device__ __forceinline__ void k1_field1(const K1Env& E, const K1Rect& R, float4 P, float rho_eps,
                                          unsigned& near_mask, bool& inside) {
    const float c = P.z, s = P.w;
    const float Cx = fmaf(c, R.mx, fmaf(-s, R.my, P.x)), Cy = fmaf(s, R.mx, fmaf(c, R.my, P.y));
    unsigned nm = 0, par = 0;
    for (int i = 0; i < E.n_field; ++i) {
        const float4 f0 = lds4(E.fld_a + 8 * i);          // Ax, Ay, Ex, Ey   (Ay <= By)
        const float4 f1 = lds4(E.fld_a + 8 * i + 4);      // nx, ny, c, By
        const float lhs = (Cx - f0.x) * f0.w, rhs = f0.z * (Cy - f0.y);
        par ^= (unsigned)(!(f0.y > Cy) && (f1.w > Cy) && (lhs < rhs));
        const float sd = fmaf(f1.x, Cx, fmaf(f1.y, Cy, -f1.z));
        nm |= (unsigned)(!(fabsf(sd) > rho_eps)) << i;
    }
    near_mask = nm; inside = par != 0;
}

// Stage D: one (pose, nearby field edge) pair.  0 = the edge is clear of the rectangle, 1 = it may touch it (band),
// 2 = it definitely cuts it (Liang-Barsky against the rectangle shrunk by eps).
__device__ __forceinline__ int k1_field2_one(const K1Env& E, const K1Rect& R, float4 P, int i) {
    const float eps = E.eps;
    const float px = P.x, py = P.y, c = P.z, s = P.w;
    const float Cx = fmaf(c, R.mx, fmaf(-s, R.my, px)), Cy = fmaf(s, R.mx, fmaf(c, R.my, py));
    const float4 r0 = lds4(E.field_a + HL_FIELD32_STRIDE * i);        // Ax, Ay, Ex, Ey
    const float4 r1 = lds4(E.field_a + HL_FIELD32_STRIDE * i + 4);    // nx, ny, c, By
    const float4 r2 = lds4(E.field_a + HL_FIELD32_STRIDE * i + 8);    // t.A, t.B
    const float Ax = r0.x, Ay = r0.y, Bx = Ax + r0.z, By = r1.w, nx = r1.x, ny = r1.y;
    const float nu = fmaf(nx, c, ny * s), nv = fmaf(ny, c, -nx * s);
    const float sd = fmaf(nx, Cx, fmaf(ny, Cy, -r1.z));                  // signed distance of the centre to the line
    if (fabsf(sd) > fmaf(R.hx, fabsf(nu), R.hy * fabsf(nv)) + eps) return 0;   // clear by the support radius along n
    const float ct = fmaf(-ny, Cx, nx * Cy);
    const float rt = fmaf(R.hx, fabsf(nv), R.hy * fabsf(nu));
    if (ct - rt > fmaxf(r2.x, r2.y) + eps || ct + rt < fminf(r2.x, r2.y) - eps) return 0;
    const float dxa = Ax - px, dya = Ay - py, dxb = Bx - px, dyb = By - py;
    const float ua = fmaf(c, dxa, s * dya), wa = fmaf(c, dya, -s * dxa);
    const float ub = fmaf(c, dxb, s * dyb), wb = fmaf(c, dyb, -s * dxb);
    if ((fminf(ua, ub) > R.x1 + eps) || (fmaxf(ua, ub) < R.x0 - eps) ||
        (fminf(wa, wb) > R.y1 + eps) || (fmaxf(wa, wb) < R.y0 - eps)) return 0;
    float t0 = 0.f, t1 = 1.f;
    bool dead = false;
    const float a2[2] = {ua, wa}, d2v[2] = {ub - ua, wb - wa};
    const float lo2[2] = {R.x0 + eps, R.y0 + eps}, hi2[2] = {R.x1 - eps, R.y1 - eps};
#pragma unroll
    for (int ax = 0; ax < 2; ++ax) {
        if (fabsf(d2v[ax]) < 1e-12f) {
            if (a2[ax] <= lo2[ax] || a2[ax] >= hi2[ax]) dead = true;
        } else {
            const float inv = f_rcp(d2v[ax]);
            const float tl = (lo2[ax] - a2[ax]) * inv, th = (hi2[ax] - a2[ax]) * inv;
            t0 = fmaxf(t0, fminf(tl, th));
            t1 = fminf(t1, fmaxf(tl, th));
        }
    }
    return (!dead && (t1 - t0) * f_sqrt(fmaf(d2v[0], d2v[0], d2v[1] * d2v[1])) > 8.0f * eps) ? 2 : 1;
}

// corner_in_capsule (hl_geom.cuh) on the staged segment + its derived record
__device__ __forceinline__ int k1_corner_in_capsule(float4 sg, float4 sd, float wx, float wy, float eps) {
    const float ex = sd.x, ey = sd.y;
    const float qx = wx - sg.x, qy = wy - sg.y;
    const float tt = fmaf(qx, ex, qy * ey) * sd.z;
    const float t = fminf(fmaxf(tt, 0.f), 1.f);
    const float ddx = fmaf(-t, ex, qx), ddy = fmaf(-t, ey, qy);
    const float d2 = fmaf(ddx, ddx, ddy * ddy);
    const float rin = (float)HL_LANE_RIN - eps, rout = (float)HL_LANE_R + eps;
    if (d2 <= rin * rin) return 1;
    if (d2 > rout * rout) return 0;
    float g = f_sqrt(d2);
    if (tt < 0.f || tt > 1.f) {                                 // cap: chord fan of the GEOS buffer polygon
        const float il = sd.w;
        const float along = fabsf(fmaf(ddx, ex, ddy * ey)) * il, across = fabsf(fmaf(ddy, ex, -ddx * ey)) * il;
        const float phi = atan2f(across, along);
        const float q = 1.5707963267948966f - phi;
        const float step = 0.09817477042468103f;                // pi/32
        const float k = floorf(q / step);
        const float delta = fabsf(q - (k + 0.5f) * step);
        g = g * cosf(delta) * 1.0012061467251643f;              // / cos(pi/64)
    }
    if (g <= (float)HL_LANE_R - eps) return 1;
    if (g > (float)HL_LANE_R + eps) return 0;
    return 2;
}

// Stage E: the four corners against every capsule polygon, then the piecewise cover certificate.
// Returns HL_HIT, HL_FREE, or HL_AMBIG (cover not certified in float32).
__device__ __forceinline__ int k1_lane_corners(const K1Env& E, const K1Rect& R, float4 P) {
    const float eps = E.eps;
    const float px = P.x, py = P.y, c = P.z, s = P.w;
    const float rin = (float)HL_LANE_RIN - eps;
    float rx[4], ry[4];
    const float lx[4] = {R.x0, R.x0, R.x1, R.x1};
    const float ly[4] = {R.y1, R.y0, R.y0, R.y1};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        rx[j] = fmaf(c, lx[j], fmaf(-s, ly[j], px));
        ry[j] = fmaf(s, lx[j], fmaf(c, ly[j], py));
    }
    bool one_holds_all = false;
    unsigned maybe = 0;
    for (int i = 0; i < E.n_seg; ++i) {
        const float4 sg = lds4(E.seg_a + 4 * i), sd = lds4(E.segd_a + 4 * i);
        bool all_in = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int st = k1_corner_in_capsule(sg, sd, rx[j], ry[j], eps);
            all_in = all_in && (st == 1);
            if (st != 0) maybe |= 1u << j;
        }
        if (all_in) one_holds_all = true;
    }
    if (one_holds_all) return HL_FREE;
    if (maybe != 0xFu) return HL_HIT;
    unsigned prev = 0;
    bool covered = true;
#pragma unroll 1
    for (int q = 0; q <= 4 && covered; ++q) {
        const float lxq = R.x0 + 0.25f * (float)q * (R.x1 - R.x0);
        unsigned m = 0xFFFFu;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const float lyq = side ? R.y1 : R.y0;
            const float wx = fmaf(c, lxq, fmaf(-s, lyq, px)), wy = fmaf(s, lxq, fmaf(c, lyq, py));
            unsigned in = 0;
            for (int i = 0; i < E.n_seg; ++i) {
                const float4 sg = lds4(E.seg_a + 4 * i), sd = lds4(E.segd_a + 4 * i);
                const float qx = wx - sg.x, qy = wy - sg.y;
                const float t = fminf(fmaxf(fmaf(qx, sd.x, qy * sd.y) * sd.z, 0.f), 1.f);
                const float ddx = fmaf(-t, sd.x, qx), ddy = fmaf(-t, sd.y, qy);
                if (fmaf(ddx, ddx, ddy * ddy) <= rin * rin) in |= 1u << i;
            }
            m &= in;
        }
        if (q > 0 && (prev & m) == 0) covered = false;
        prev = m;
    }
    return covered ? HL_FREE : HL_AMBIG;
}

// Monolithic per-pose filter on global memory for slices the staged pipeline cannot take.
__device__ __noinline__ void k1_slow_slice(const EnvBatchDev& eb, const int32_t* env_id, const double* poses,
                                           const int32_t* pose_idx, long long n, long long base, unsigned flags,
                                           uint8_t* out, unsigned long long* n_exact, int lane) {
    for (int j = 0; j < K1_PER_LANE; ++j) {
        const long long i = base + j * 32 + lane;
        const bool active = i < n;
        const int e = active ? (env_id ? env_id[i] : 0) : 0;
        const EnvDesc& D = eb.desc[e];
        EnvSmem E;
        global_env(eb, D, E);
        double x = 0.0, y = 0.0, yaw = 0.0;
        if (active) { x = poses[3 * i]; y = poses[3 * i + 1]; yaw = poses[3 * i + 2]; }
        const bool with_aux = active && (pose_idx ? ((pose_idx[i] & 1) == 0) : true);
        const float px = (float)(x - D.origin[0]), py = (float)(y - D.origin[1]);
        const bool beyond = fabsf(px) > E.reach || fabsf(py) > E.reach;
        const bool far = !beyond && (!(fabs(yaw) < 1e6) || !(px == px) || !(py == py));
        float sf, cf;
        sincosf((float)yaw, &sf, &cf);
        bool bad = false;
        unsigned amb = flags;
        int r = HL_FREE;
        if (active) r = beyond ? far_status(flags, E.n_seg) : (far ? HL_AMBIG : filter_part(E, px, py, cf, sf, E.ext, flags, &amb));
        if (r == HL_HIT) bad = true;
        warp_resolve(r == HL_AMBIG, x, y, yaw, e, D.body_ext, amb, bad, eb, n_exact, lane);
        if (flags & HL_CHECK_AUX) {
            const unsigned aflags = flags & (HL_CHECK_OBSTACLES | HL_CHECK_BOUNDARY);
            const int na = (active && with_aux && !beyond) ? D.n_aux : 0;
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

__global__ void __launch_bounds__(K1_THREADS, K1_MIN_CTAS)
k_collision(EnvBatchDev eb, const int32_t* __restrict__ env_id, const double* __restrict__ poses,
            const int32_t* __restrict__ pose_idx, long long n, unsigned flags,
            uint8_t* __restrict__ out, unsigned long long* n_exact, int use_tma) {
    extern __shared__ __align__(16) unsigned char k1_smem[];
    K1Cta& S = *reinterpret_cast<K1Cta*>(k1_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    K1Warp& W = S.w[warp];
    const unsigned wbar = smem_u32(&W.bar), cbar = smem_u32(&S.bar);
    const long long n_tiles = (n + K1_CTA_TILE - 1) / K1_CTA_TILE;
    if (threadIdx.x == 0) mbar_init(cbar, 1);
    if (lane == 0) mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned wphase = 0, cphase = 0;
    int staged_env = -1;
    bool env_ok = false;                               // staged environment usable by the pipeline
    K1Env E;
    E.obs_a = E.field_a = E.seg_a = S.env; E.segd_a = reinterpret_cast<const float*>(S.segd);
    E.fld_a = reinterpret_cast<const float*>(S.fld);
    E.n_obs = E.n_field = E.n_seg = 0; E.eps = 0.f;

    // does this warp's slice of `tile` arrive by TMA?  (full slice, 16-byte aligned source)
    auto slice_tma = [&](long long tile) -> bool {
        const long long b = tile * K1_CTA_TILE + (long long)warp * K1_TILE;
        return use_tma && tile < n_tiles && b + K1_TILE <= n;
    };
    auto issue_slice = [&](long long tile) {
        if (lane == 0 && slice_tma(tile)) {
            const long long b = tile * K1_CTA_TILE + (long long)warp * K1_TILE;
            mbar_expect_tx(wbar, K1_TILE * 24);
            bulk_g2s(smem_u32(W.raw), poses + 3 * b, K1_TILE * 24, wbar);
        }
    };
    issue_slice(blockIdx.x);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long cta_base = tile * K1_CTA_TILE;
        const long long base = cta_base + (long long)warp * K1_TILE;
        // ---- environment of the tile (CTA-uniform): staged by bulk copies when it changes
        const int e0 = env_id ? env_id[cta_base] : 0;
        if (e0 != staged_env) {
            __syncthreads();                           // nobody still reads the previous environment
            const EnvDesc& D0 = eb.desc[e0];
            const int n_o = D0.n_obs * HL_OBS32_STRIDE, n_f = D0.n_field * HL_FIELD32_STRIDE, n_s = D0.n_seg * 4;
            env_ok = (n_o + n_f + n_s <= K1_ENV_FLOATS) && D0.all_rect && D0.n_seg <= HL_MAX_SEGS && D0.n_field <= 32 && D0.n_obs <= 32;
            if (env_ok) {
                if (threadIdx.x == 0) {
                    fence_proxy_async();
                    const unsigned bytes = 4u * (unsigned)(n_o + n_f + n_s);
                    if (bytes) {
                        mbar_expect_tx(cbar, bytes);
                        if (n_o) bulk_g2s(smem_u32(S.env), eb.obs32 + (size_t)HL_OBS32_STRIDE * D0.obs_off, 4u * n_o, cbar);
                        if (n_f) bulk_g2s(smem_u32(S.env + n_o), eb.field32 + (size_t)HL_FIELD32_STRIDE * D0.field_off, 4u * n_f, cbar);
                        if (n_s) bulk_g2s(smem_u32(S.env + n_o + n_f), eb.seg32 + 4 * (size_t)D0.seg_off, 4u * n_s, cbar);
                    }
                }
                if (n_o + n_f + n_s) { mbar_wait(cbar, cphase); cphase ^= 1u; }
                if (threadIdx.x < D0.n_seg) {          // derived per-segment record: ex, ey, 1/len^2, 1/len
                    const float* sg = S.env + n_o + n_f + 4 * threadIdx.x;
                    const float ex = sg[2] - sg[0], ey = sg[3] - sg[1];
                    const float len2 = fmaf(ex, ex, ey * ey);
                    S.segd[threadIdx.x] = make_float4(ex, ey, f_rcp(len2), rsqrtf(len2));
                }
                if (threadIdx.x >= 32 && threadIdx.x < 32 + D0.n_field) {     // first-pass field records, endpoints ordered by y
                    const int i = threadIdx.x - 32;
                    const float* e = S.env + n_o + HL_FIELD32_STRIDE * i;
                    const bool up = !(e[1] > e[7]);                    // Ay <= By
                    S.fld[i][0] = up ? make_float4(e[0], e[1], e[2], e[3]) : make_float4(e[0] + e[2], e[7], -e[2], -e[3]);
                    S.fld[i][1] = make_float4(e[4], e[5], e[6], up ? e[7] : e[1]);
                }
                E.obs_a = S.env; E.field_a = S.env + n_o; E.seg_a = S.env + n_o + n_f;
                E.n_obs = D0.n_obs; E.n_field = D0.n_field; E.n_seg = D0.n_seg; E.eps = D0.eps;
            }
            staged_env = e0;
            __syncthreads();
        }
        if (base >= n) { continue; }                   // warp without poses in this (last) tile
        const EnvDesc& D = eb.desc[e0];
        const bool tma = slice_tma(tile);
        if (tma) { mbar_wait(wbar, wphase); wphase ^= 1u; }
        // ---- stage 0: poses -> float32 frame; is the whole slice in the staged environment?
        bool same_env = env_ok;
#pragma unroll 1
        for (int j = 0; j < K1_PER_LANE; ++j) {
            const int q = j * 32 + lane;
            const long long i = base + q;
            const bool active = i < n;
            double x = 0.0, y = 0.0, yaw = 0.0;
            if (active) {
                if (tma) { x = W.raw[3 * q]; y = W.raw[3 * q + 1]; yaw = W.raw[3 * q + 2]; }
                else { x = poses[3 * i]; y = poses[3 * i + 1]; yaw = poses[3 * i + 2]; }
                if (env_id && env_id[i] != e0) same_env = false;
            }
            const float px = (float)(x - D.origin[0]), py = (float)(y - D.origin[1]);
            const bool beyond = fabsf(px) > D.reach || fabsf(py) > D.reach;
            const bool far = !beyond && (!(fabs(yaw) < 1e6) || !(px == px) || !(py == py));
            // float32 sine / cosine for the filter: the angle is reduced to [-pi, pi] in float64 (exact enough for any
            // |yaw| < 1e6), then the hardware approximations (abs. error 2^-21.4 there) -- an error of < 2e-6 m on
            // a footprint corner, far inside the band eps
            const double kk = rint(yaw * 0.15915494309189535);
            const float yr = (float)fma(-kk, 6.283185307179586, yaw);
            const float sf = __sinf(yr), cf = __cosf(yr);
            W.pose[q] = make_float4(px, py, cf, sf);
            unsigned char st = 0;
            if (!active) st = K1S_OFF | K1S_DONE;
            else if (beyond) st = K1S_DONE | (far_status(flags, D.n_seg) == HL_HIT ? K1S_HIT : 0) | K1S_OFF;   // no aux test either
            else if (far) st = K1S_FAR;
            W.st[q] = st;
            W.amb[q] = 0;
        }
        same_env = __all_sync(0xffffffffu, same_env);
        __syncwarp();
        if (tma || slice_tma(tile + gridDim.x)) {      // the landing zone is free again: fetch the next slice
            if (lane == 0) fence_proxy_async();
            issue_slice(tile + gridDim.x);
        }
        if (!same_env) {
            k1_slow_slice(eb, env_id, poses, pose_idx, n, base, flags, out, n_exact, lane);
            continue;
        }
        // ---- rectangles of the footprint: 0 = body (obstacles, field, lane), 1.. = implement rectangles (obstacles +
        // field polygon, never the lane, poses 0,2,4,.. of a path: orchard_geometry_environment.py:439-456,
        // car_model.py:58).  ONE instance of the stage code serves all of them.
        const int n_rect = ((flags & HL_CHECK_AUX) ? D.n_aux : 0) + 1;
#pragma unroll 1
        for (int rect = 0; rect < n_rect; ++rect) {
            const double* ext64 = rect == 0 ? D.body_ext : eb.aux64 + 4 * (size_t)(D.aux_off + rect - 1);
            float ext[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) ext[k] = (float)ext64[k];
            const K1Rect R = k1_rect(ext);
            const unsigned rflags = rect == 0 ? flags : (flags & (HL_CHECK_OBSTACLES | HL_CHECK_BOUNDARY));
            const bool do_lane = rect == 0 && (flags & HL_CHECK_LANE) && E.n_seg > 0;
            const float rho = sqrtf(fmaf(R.hx, R.hx, R.hy * R.hy));
            // ---- stage A: input list.  Body: every live pose, minus those the lane centre test rejects (cheapest
            // test, decides every pose far from the guide).  Implement: live poses at even path indices.
            int n_a = 0;
#pragma unroll 1
            for (int j = 0; j < K1_PER_LANE; ++j) {
                const int q = j * 32 + lane;
                unsigned char st = W.st[q];
                bool live = !(st & (K1S_DONE | K1S_FAR));
                if (rect > 0) {
                    const bool with_aux = !(st & (K1S_DONE | K1S_OFF)) && (pose_idx ? ((pose_idx[base + q] & 1) == 0) : true);
                    live = live && with_aux;
                    W.amb[q] = (with_aux && (st & K1S_FAR)) ? (unsigned char)rflags : (unsigned char)0;
                } else if (live && do_lane) {
                    const int r = k1_lane_centre(E, R, W.pose[q], rho);
                    if (r == 1) { st |= K1S_HIT; live = false; }
                    else if (r == 2) st |= K1S_CORNERS;
                    W.st[q] = st;
                }
                n_a = warp_append(W.la, n_a, live, q, lane);
            }
            __syncwarp();
            const float rho_eps = rho + E.eps;
            // ---- stage B: obstacles (la -> lb).  Pass 1 culls per pose to the obstacles within the circumradius, pass 2
            // runs the separating-axis test on dense (pose, obstacle) pairs.
            int n_b = 0;
            if (rflags & HL_CHECK_OBSTACLES) {
                int n_pairs = 0;
                auto process = [&](int np) {
                    __syncwarp();
#pragma unroll 1
                    for (int t = lane; t < np; t += 32) {
                        const unsigned pr = W.pairs[t];
                        const int q = pr & 0xFF;
                        bool hit, amb;
                        k1_obstacle_one(E, R, W.pose[q], (int)(pr >> 8), hit, amb);
                        if (hit) st_or(W.st, q, K1S_HIT);
                        else if (amb) st_or(W.amb, q, HL_CHECK_OBSTACLES);
                    }
                };
#pragma unroll 1
                for (int k = 0; k < n_a; k += 32) {
                    const int idx = k + lane;
                    const bool v = idx < n_a;
                    const int q = v ? W.la[idx] : 0;
                    const unsigned m = v ? k1_obstacle_mask(E, R, W.pose[q], rho_eps) : 0u;
                    if (!k1_push_pairs(W, n_pairs, m, q, lane, process)) {       // > K1_PAIRS pairs in one chunk: per pose
                        unsigned mm = m;
                        while (mm) {
                            const int ob = __ffs(mm) - 1;
                            mm &= mm - 1;
                            bool hit, amb;
                            k1_obstacle_one(E, R, W.pose[q], ob, hit, amb);
                            if (hit) W.st[q] |= K1S_HIT;
                            else if (amb) W.amb[q] |= HL_CHECK_OBSTACLES;
                        }
                    }
                }
                process(n_pairs);
                __syncwarp();
#pragma unroll 1
                for (int k = 0; k < n_a; k += 32) {
                    const int idx = k + lane;
                    const bool v = idx < n_a;
                    const int q = v ? W.la[idx] : 0;
                    n_b = warp_append(W.lb, n_b, v && !(W.st[q] & K1S_HIT), q, lane);
                }
            } else {
                for (int k = lane; k < n_a; k += 32) W.lb[k] = W.la[k];
                n_b = n_a;
            }
            __syncwarp();
            // ---- stages C + D: field polygon (lb -> survivors in lc; poses with nearby edges wait in la)
            int n_c = 0;
            if (rflags & HL_CHECK_BOUNDARY) {
                int n_d = 0, n_pairs = 0;
                auto process = [&](int np) {
                    __syncwarp();
#pragma unroll 1
                    for (int t = lane; t < np; t += 32) {
                        const unsigned pr = W.pairs[t];
                        const int q = pr & 0xFF;
                        const int r = k1_field2_one(E, R, W.pose[q], (int)(pr >> 8));
                        if (r == 2) st_or(W.st, q, K1S_HIT);
                        else if (r == 1) st_or(W.st, q, K1S_NOTCLEAR);
                    }
                };
#pragma unroll 1
                for (int k = 0; k < n_b; k += 32) {
                    const int idx = k + lane;
                    const bool v = idx < n_b;
                    const int q = v ? W.lb[idx] : 0;
                    bool keep = false, need2 = false;
                    unsigned nm = 0;
                    if (v) {
                        bool inside;
                        k1_field1(E, R, W.pose[q], rho_eps, nm, inside);
                        if (nm == 0) {                            // every edge clear: the parity of the centre decides
                            if (inside) keep = true; else W.st[q] |= K1S_HIT;
                        } else {
                            need2 = true;
                            W.st[q] = (unsigned char)((W.st[q] & ~(K1S_INSIDE | K1S_NOTCLEAR)) | (inside ? K1S_INSIDE : 0));
                        }
                    }
                    n_c = warp_append(W.lc, n_c, keep, q, lane);
                    n_d = warp_append(W.la, n_d, need2, q, lane);
                    __syncwarp();
                    if (!k1_push_pairs(W, n_pairs, nm, q, lane, process)) {
                        while (nm) {
                            const int ed = __ffs(nm) - 1;
                            nm &= nm - 1;
                            const int r = k1_field2_one(E, R, W.pose[q], ed);
                            if (r == 2) W.st[q] |= K1S_HIT;
                            else if (r == 1) W.st[q] |= K1S_NOTCLEAR;
                        }
                    }
                }
                process(n_pairs);
                __syncwarp();
#pragma unroll 1
                for (int k = 0; k < n_d; k += 32) {               // verdict of the poses that had nearby edges
                    const int idx = k + lane;
                    const bool v = idx < n_d;
                    const int q = v ? W.la[idx] : 0;
                    bool keep = false;
                    if (v) {
                        const unsigned char st = W.st[q];
                        if (st & K1S_HIT) {}
                        else if (st & K1S_NOTCLEAR) { keep = true; W.amb[q] |= HL_CHECK_BOUNDARY; }
                        else if (st & K1S_INSIDE) keep = true;
                        else W.st[q] = st | K1S_HIT;
                    }
                    n_c = warp_append(W.lc, n_c, keep, q, lane);
                }
            } else {
                for (int k = lane; k < n_b; k += 32) W.lc[k] = W.lb[k];
                n_c = n_b;
            }
            __syncwarp();
            // ---- stage E: lane corners of the body survivors that need them
            if (do_lane) {
#pragma unroll 1
                for (int k = 0; k < n_c; k += 32) {
                    const int idx = k + lane;
                    if (idx < n_c) {
                        const int q = W.lc[idx];
                        if (W.st[q] & K1S_CORNERS) {
                            const int r = k1_lane_corners(E, R, W.pose[q]);
                            if (r == HL_HIT) W.st[q] |= K1S_HIT;
                            else if (r == HL_AMBIG) W.amb[q] |= HL_CHECK_LANE;
                        }
                    }
                }
                __syncwarp();
            }
            // ---- float64 resolution of this rectangle; then "infeasible" becomes K1S_DONE | K1S_HIT
#pragma unroll 1
            for (int j = 0; j < K1_PER_LANE; ++j) {
                const int q = j * 32 + lane;
                unsigned char st = W.st[q];
                const unsigned amb = (rect == 0 && (st & K1S_FAR)) ? rflags : (unsigned)W.amb[q];
                const bool need = !(st & (K1S_HIT | K1S_DONE)) && amb != 0;
                if (__any_sync(0xffffffffu, need)) {        // rare: keep the call (and its spills) off the common path
                    const long long i = base + q;
                    double x = 0.0, y = 0.0, yaw = 0.0;
                    if (need) { x = poses[3 * i]; y = poses[3 * i + 1]; yaw = poses[3 * i + 2]; }
                    bool b2 = false;
                    warp_resolve(need, x, y, yaw, e0, ext64, amb, b2, eb, n_exact, lane);
                    if (b2) st |= K1S_HIT;
                }
                if (st & K1S_HIT) st |= K1S_DONE;
                W.st[q] = st;
                W.amb[q] = 0;
            }
            __syncwarp();
        }
        // ---- verdicts: 4 poses per lane, one 32-bit store per lane when the output is aligned
        if (base + K1_TILE <= n && ((reinterpret_cast<uintptr_t>(out) & 3) == 0)) {
            const uchar4 s4 = *reinterpret_cast<const uchar4*>(&W.st[4 * lane]);
            uchar4 o4;
            o4.x = s4.x & K1S_HIT; o4.y = s4.y & K1S_HIT; o4.z = s4.z & K1S_HIT; o4.w = s4.w & K1S_HIT;
            *reinterpret_cast<uchar4*>(out + base + 4 * lane) = o4;
        } else {
#pragma unroll 1
            for (int j = 0; j < K1_PER_LANE; ++j) {
                const long long i = base + j * 32 + lane;
                if (i < n) out[i] = (W.st[j * 32 + lane] & K1S_HIT) ? 1 : 0;
            }
        }
        __syncwarp();
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
    const int smem_bytes = (int)sizeof(K1Cta);
    static_assert(sizeof(K1Cta) * K1_MIN_CTAS + 1024 * K1_MIN_CTAS <= 227 * 1024, "K1 shared memory exceeds the SM");
    HL_CUDA_OK(cudaFuncSetAttribute(k_collision, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const long long tiles = (n + K1_CTA_TILE - 1) / K1_CTA_TILE;
    const long long cap = (long long)ctx->sm_count * K1_MIN_CTAS;
    const int grid = (int)(tiles < cap ? tiles : cap);
    const int use_tma = (((uintptr_t)d_poses) & 15) == 0 ? 1 : 0;     // cp.async.bulk needs a 16-byte aligned source
    k_collision<<<grid, K1_THREADS, smem_bytes, (cudaStream_t)stream>>>(
        envs->dev, d_env_id, d_poses, d_pose_idx, (long long)n, flags, d_out, d_n_exact, use_tma);
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
