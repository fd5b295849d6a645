// hl_dubins.cu -- K9: batched Dubins paths and their cubic-spline course.
//
// Replaces get_dubins_path (path_planner/utils/navigation_utils.py:206-215: dubins.shortest_path + sample_many) and
// get_dubins_path_full (path_planner/safety_forward_path_plan.py:286-297: + calc_spline_course, cubic_spline.py:92-112)
// for MANY start / end pose pairs at once -- the candidate generator of sample_start_end_pose_for_dubins,
// get_circle_back_path_full and of the Pawn goal extension (hybrid_a_star_search.py:289-304, which also appends the
// goal pose to the samples).  Two launches like K8: k_dubins_count (warp per pair: shortest word, number of samples,
// knots, chord length -> rows) and, after the caller's prefix sums, k_dubins_fill (warp per pair: lane 0 solves the
// tridiagonal systems, all lanes evaluate the rows x, y, yaw, curvature).  float64; pydubins parity is unpinned.
#include "hl_dubins.cuh"

#define DUB_WS 9                       // workspace doubles per sample slot: ts kx ky ks dx dy cp bx by

__global__ void __launch_bounds__(128) k_dubins_count(const double* __restrict__ pairs, long long n, double rho, double step,
                                                      double ds, int append_goal, long long* __restrict__ n_slots,
                                                      int* __restrict__ word, double* __restrict__ length) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= n) return;
    DubPath P;
    dub_shortest(pairs + 6 * p, pairs + 6 * p + 3, rho, P);
    word[p] = P.type;
    const double L = P.type >= 0 ? dub_length(P) : 0.0;
    length[p] = L;
    n_slots[p] = P.type >= 0 ? dub_n_samples(L, step) + 1 : 0;
}

__global__ void __launch_bounds__(128) k_dubins_knots(const double* __restrict__ pairs, long long n, double rho, double step,
                                                      double ds, int append_goal, const long long* __restrict__ slot_off,
                                                      double* __restrict__ ws, int* __restrict__ n_knots,
                                                      long long* __restrict__ n_rows, double* __restrict__ samples) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (warp >= n) return;
    const long long p = warp;
    const long long cap = slot_off[p + 1] - slot_off[p];
    int m = 0;
    long long rows = 0;
    if (cap > 0) {
        DubPath P;
        dub_shortest(pairs + 6 * p, pairs + 6 * p + 3, rho, P);
        double* ts = ws + DUB_WS * slot_off[p];
        double* kx = ts + cap; double* ky = kx + cap; double* ks = ky + cap;
        m = dub_course_knots(P, step, append_goal != 0, pairs + 6 * p + 3, cap - 1, ts, kx, ky, ks, lane,
                             samples ? samples + 3 * slot_off[p] : nullptr);
        if (m >= 2) rows = rp_count(ks[m - 1], ds);
    }
    if (lane == 0) { n_knots[p] = m; n_rows[p] = rows; }
}

__global__ void __launch_bounds__(128) k_dubins_fill(long long n, double ds, const long long* __restrict__ slot_off,
                                                     const long long* __restrict__ row_off, const int* __restrict__ n_knots,
                                                     double* __restrict__ ws, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (warp >= n) return;
    const long long p = warp;
    const int m = n_knots[p];
    if (m < 2) return;
    const long long cap = slot_off[p + 1] - slot_off[p];
    double* ts = ws + DUB_WS * slot_off[p];
    double* kx = ts + cap; double* ky = kx + cap; double* ks = ky + cap; double* dx = ks + cap; double* dy = dx + cap;
    double* cp = dy + cap; double* bx = cp + cap; double* by = bx + cap;
    if (lane == 0) rp_derivs(ks, kx, ky, m, dx, dy, cp, bx, by);
    __syncwarp();
    const long long cnt = row_off[p + 1] - row_off[p];
    double* o = out + 4 * row_off[p];
    for (long long i = lane; i < cnt; i += 32) {
        const double t = xmul((double)i, ds);                   // np.arange(0, s_end + ds, ds)[i]
        double x, x1, x2, y, y1, y2;
        rp_eval(ks, kx, dx, m, t, x, x1, x2);
        rp_eval(ks, ky, dy, m, t, y, y1, y2);
        const double q = x1 * x1 + y1 * y1;
        o[4 * i] = x; o[4 * i + 1] = y; o[4 * i + 2] = m_atan2(y1, x1);
        o[4 * i + 3] = (y2 * x1 - x2 * y1) / (q * sqrt(q));
    }
}

extern "C" int hl_dubins_count(hl_ctx* ctx, const double* d_pairs, int64_t n, double rho, double step, double ds,
                               int32_t append_goal, int64_t* d_slots, int32_t* d_word, double* d_length, void* stream) {
    if (!ctx || !d_pairs || !d_slots || !d_word || !d_length || n < 0 || !(rho > 0) || !(step > 0) || !(ds > 0)) {
        hl_set_error("hl_dubins_count: bad arguments"); return 1;
    }
    if (n == 0) return 0;
    if (hl_enter(ctx, nullptr, d_slots, "hl_dubins_count")) return 1;
    k_dubins_count<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_pairs, n, rho, step, ds, append_goal,
                                                                                 (long long*)d_slots, d_word, d_length);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int hl_dubins_knots(hl_ctx* ctx, const double* d_pairs, int64_t n, double rho, double step, double ds,
                               int32_t append_goal, const int64_t* d_slot_offsets, double* d_workspace,
                               int32_t* d_n_knots, int64_t* d_n_rows, double* d_samples, void* stream) {
    if (!ctx || !d_pairs || !d_slot_offsets || !d_workspace || !d_n_knots || !d_n_rows || n < 0) {
        hl_set_error("hl_dubins_knots: bad arguments"); return 1;
    }
    if (n == 0) return 0;
    if (hl_enter(ctx, nullptr, d_workspace, "hl_dubins_knots")) return 1;
    k_dubins_knots<<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        d_pairs, n, rho, step, ds, append_goal, (const long long*)d_slot_offsets, d_workspace, d_n_knots, (long long*)d_n_rows, d_samples);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int hl_dubins_fill(hl_ctx* ctx, int64_t n, double ds, const int64_t* d_slot_offsets, const int64_t* d_row_offsets,
                              const int32_t* d_n_knots, double* d_workspace, double* d_out, void* stream) {
    if (!ctx || !d_slot_offsets || !d_row_offsets || !d_n_knots || !d_workspace || !d_out || n < 0) {
        hl_set_error("hl_dubins_fill: bad arguments"); return 1;
    }
    if (n == 0) return 0;
    if (hl_enter(ctx, nullptr, d_out, "hl_dubins_fill")) return 1;
    k_dubins_fill<<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        n, ds, (const long long*)d_slot_offsets, (const long long*)d_row_offsets, d_n_knots, d_workspace, d_out);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- K10: get_min_distance_to_boundary (path_planner/orchard_geometry_environment.py:393-412) -------------------
// The reference unions the footprint rectangles of a path (shapely unary_union: body at every pose, each implement
// rectangle at every 2nd pose, car_model.py:39-73), walks the exterior vertices of every union and returns the
// smallest distance to the field polygon's ring, negative for vertices that are not inside the field.  The vertices
// of a union of rectangles are the rectangle corners that no other rectangle covers plus the crossings of two
// rectangle edges that no third rectangle covers -- computed here per path by one CTA (GEOS itself is not available:
// parity unpinned; vertices of holes of the union, which `.exterior` would skip, are not told apart).
#include "hl_geom.cuh"

struct MdRect { double x, y, c, s; };

__device__ __forceinline__ bool md_strictly_inside(const MdRect& r, const double* ext, double px, double py) {
    const double dx = px - r.x, dy = py - r.y;
    const double u = r.c * dx + r.s * dy, w = r.c * dy - r.s * dx;
    const double tol = 1e-12;                 // a vertex ON another rectangle's boundary is a boundary vertex
    return u > ext[0] + tol && u < ext[1] - tol && w > ext[2] + tol && w < ext[3] - tol;
}

__device__ double md_signed_distance(const double* poly, int n, double px, double py) {
    double best = INFINITY;
    bool inside = false, on_edge = false;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        const double ax = poly[2 * i], ay = poly[2 * i + 1], bx = poly[2 * j], by = poly[2 * j + 1];
        const double ex = bx - ax, ey = by - ay;
        double t = ((px - ax) * ex + (py - ay) * ey) / (ex * ex + ey * ey);
        t = fmin(fmax(t, 0.0), 1.0);
        const double d = hypot_cr(px - (ax + t * ex), py - (ay + t * ey));
        best = fmin(best, d);
        if (d == 0.0) on_edge = true;
        if ((ay > py) != (by > py)) {
            const double xint = (bx - ax) * (py - ay) / (by - ay) + ax;
            if (px < xint) inside = !inside;
        }
    }
    return (inside && !on_edge) ? best : -best;
}

__global__ void __launch_bounds__(256) k_min_boundary_distance(EnvBatchDev eb, const int32_t* __restrict__ env_id,
                                                               const double* __restrict__ poses,
                                                               const long long* __restrict__ path_start, long long n_paths,
                                                               int with_aux, double* __restrict__ out) {
    const long long p = blockIdx.x;
    if (p >= n_paths) return;
    const EnvDesc& D = eb.desc[env_id ? env_id[p] : 0];
    const double* field = eb.field64 + 2 * (size_t)D.field_off;
    const long long a = path_start[p], b = path_start[p + 1];
    const int P = (int)(b - a);
    __shared__ double s_min[256];
    double best = INFINITY;
    const int n_parts = 1 + (with_aux ? D.n_aux : 0);
    for (int part = 0; part < n_parts; ++part) {
        const int stride = part == 0 ? 1 : 2;
        const double* ext = part == 0 ? D.body_ext : eb.aux64 + 4 * (size_t)(D.aux_off + part - 1);
        const int R = (P + stride - 1) / stride;
        const double reach2 = 4.0 * (fmax(fabs(ext[0]), fabs(ext[1])) * fmax(fabs(ext[0]), fabs(ext[1])) +
                                     fmax(fabs(ext[2]), fabs(ext[3])) * fmax(fabs(ext[2]), fabs(ext[3])));
        auto rect = [&](int r) {
            const double* q = poses + 3 * (a + (long long)r * stride);
            MdRect m; m.x = q[0]; m.y = q[1]; m.c = cos(q[2]); m.s = sin(q[2]);
            return m;
        };
        auto corner = [&](const MdRect& m, int k, double& cx, double& cy) {
            const double lx = (k == 0 || k == 1) ? ext[0] : ext[1];
            const double ly = (k == 0 || k == 3) ? ext[3] : ext[2];
            cx = m.c * lx - m.s * ly + m.x;
            cy = m.s * lx + m.c * ly + m.y;
        };
        // (a) corners that no other rectangle of the part covers
        for (int idx = threadIdx.x; idx < 4 * R; idx += blockDim.x) {
            const int r = idx >> 2, k = idx & 3;
            const MdRect m = rect(r);
            double cx, cy;
            corner(m, k, cx, cy);
            bool covered = false;
            for (int j = 0; j < R && !covered; ++j) {
                if (j == r) continue;
                const MdRect o = rect(j);
                if ((o.x - cx) * (o.x - cx) + (o.y - cy) * (o.y - cy) > reach2) continue;
                covered = md_strictly_inside(o, ext, cx, cy);
            }
            if (!covered) best = fmin(best, md_signed_distance(field, D.n_field, cx, cy));
        }
        // (b) crossings of two rectangles' edges that no third rectangle covers
        const long long n_pairs = (long long)R * (R - 1) / 2;
        for (long long pr = threadIdx.x; pr < n_pairs; pr += blockDim.x) {
            // pair index -> (i, j), i < j
            int j = (int)((1.0 + sqrt(1.0 + 8.0 * (double)pr)) * 0.5);
            while ((long long)j * (j - 1) / 2 > pr) --j;
            while ((long long)(j + 1) * j / 2 <= pr) ++j;
            const int i = (int)(pr - (long long)j * (j - 1) / 2);
            const MdRect A = rect(i), B = rect(j);
            if ((A.x - B.x) * (A.x - B.x) + (A.y - B.y) * (A.y - B.y) > reach2) continue;
            double ax[4], ay[4], bx[4], by[4];
            for (int k = 0; k < 4; ++k) { corner(A, k, ax[k], ay[k]); corner(B, k, bx[k], by[k]); }
            for (int ea = 0; ea < 4; ++ea) {
                const double p0x = ax[ea], p0y = ay[ea], rx = ax[(ea + 1) & 3] - p0x, ry = ay[(ea + 1) & 3] - p0y;
                for (int eb2 = 0; eb2 < 4; ++eb2) {
                    const double q0x = bx[eb2], q0y = by[eb2], sx = bx[(eb2 + 1) & 3] - q0x, sy = by[(eb2 + 1) & 3] - q0y;
                    const double den = rx * sy - ry * sx;
                    if (fabs(den) < 1e-14) continue;                          // parallel (collinear overlaps add no vertex)
                    const double t = ((q0x - p0x) * sy - (q0y - p0y) * sx) / den;
                    const double u = ((q0x - p0x) * ry - (q0y - p0y) * rx) / den;
                    if (t < 0.0 || t > 1.0 || u < 0.0 || u > 1.0) continue;
                    const double ix = p0x + t * rx, iy = p0y + t * ry;
                    bool covered = false;
                    for (int k = 0; k < R && !covered; ++k) {
                        if (k == i || k == j) continue;
                        const MdRect o = rect(k);
                        if ((o.x - ix) * (o.x - ix) + (o.y - iy) * (o.y - iy) > reach2) continue;
                        covered = md_strictly_inside(o, ext, ix, iy);
                    }
                    if (!covered) best = fmin(best, md_signed_distance(field, D.n_field, ix, iy));
                }
            }
        }
    }
    s_min[threadIdx.x] = best;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) s_min[threadIdx.x] = fmin(s_min[threadIdx.x], s_min[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[p] = s_min[0];
}

extern "C" int hl_min_boundary_distance(hl_ctx* ctx, const hl_env_batch* envs, const int32_t* d_env_id, const double* d_poses,
                                        const int64_t* d_path_start, int64_t n_paths, int32_t with_aux, double* d_out,
                                        void* stream) {
    if (!ctx || !envs || !d_poses || !d_path_start || !d_out || n_paths < 0) { hl_set_error("hl_min_boundary_distance: bad arguments"); return 1; }
    if (n_paths == 0) return 0;
    if (hl_enter(ctx, envs, d_out, "hl_min_boundary_distance")) return 1;
    k_min_boundary_distance<<<(unsigned)n_paths, 256, 0, (cudaStream_t)stream>>>(envs->dev, d_env_id, d_poses,
                                                                               (const long long*)d_path_start, n_paths, with_aux, d_out);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- K11: corridor test of classic_circle_back_turning_path (path_planner/safety_forward_path_plan.py:811-822) -----
// `LineString(points).buffer(0.3, cap_style=flat, join_style=round)` intersects the nearest obstacle polygon  <=>  some
// segment's flat-capped rectangle (half width r) or some interior vertex's disc of radius r meets some obstacle
// (closed sets).  One thread per polyline point; float64; GEOS approximates the join discs by 32-gons (parity unpinned).
__global__ void k_corridor_hits(EnvBatchDev eb, const int32_t* __restrict__ env_id, const double* __restrict__ pts,
                                const long long* __restrict__ line_start, long long n_lines, double r,
                                uint8_t* __restrict__ out) {
    const long long line = blockIdx.x;
    if (line >= n_lines) return;
    const EnvDesc& D = eb.desc[env_id ? env_id[line] : 0];
    const double* V = eb.obs64 + 8 * (size_t)D.obs_off;
    const long long a = line_start[line], b = line_start[line + 1];
    int hit = 0;
    for (long long k = a + threadIdx.x; k < b && !hit; k += blockDim.x) {
        const double x0 = pts[2 * k], y0 = pts[2 * k + 1];
        if (k + 1 < b) {                                     // rectangle of segment k -> k+1
            const double dx = pts[2 * k + 2] - x0, dy = pts[2 * k + 3] - y0;
            const double len = sqrt(dx * dx + dy * dy);
            if (len > 0.0) {
                Pose64 p; p.x = x0; p.y = y0; p.c = dx / len; p.s = dy / len;
                const double ext[4] = {0.0, len, -r, r};
                double cx[4], cy[4];
                exact_corners(p, ext, cx, cy);
                for (int o = 0; o < D.n_obs && !hit; ++o) hit = exact_rect_hits_quad(p, ext, cx, cy, V + 8 * o) ? 1 : 0;
            }
        }
        if (k > a && k + 1 < b && !hit) {                    // round join: disc about an interior vertex
            for (int o = 0; o < D.n_obs && !hit; ++o) {
                const double* Q = V + 8 * o;
                bool inside = true;
                double best = INFINITY;
                for (int i = 0; i < 4; ++i) {
                    const int j = (i + 1) & 3;
                    const double ex = Q[2 * j] - Q[2 * i], ey = Q[2 * j + 1] - Q[2 * i + 1];
                    if (ex * (y0 - Q[2 * i + 1]) - ey * (x0 - Q[2 * i]) < 0.0) inside = false;     // CCW quad: left of every edge
                    double t = ((x0 - Q[2 * i]) * ex + (y0 - Q[2 * i + 1]) * ey) / (ex * ex + ey * ey);
                    t = fmin(fmax(t, 0.0), 1.0);
                    best = fmin(best, hypot_cr(x0 - (Q[2 * i] + t * ex), y0 - (Q[2 * i + 1] + t * ey)));
                }
                if (inside || best <= r) hit = 1;
            }
        }
    }
    hit = __syncthreads_or(hit);
    if (threadIdx.x == 0) out[line] = hit ? 1 : 0;
}

extern "C" int hl_corridor_hits(hl_ctx* ctx, const hl_env_batch* envs, const int32_t* d_env_id, const double* d_points,
                                const int64_t* d_line_start, int64_t n_lines, double radius, uint8_t* d_out, void* stream) {
    if (!ctx || !envs || !d_points || !d_line_start || !d_out || n_lines < 0 || !(radius >= 0)) { hl_set_error("hl_corridor_hits: bad arguments"); return 1; }
    if (n_lines == 0) return 0;
    if (hl_enter(ctx, envs, d_out, "hl_corridor_hits")) return 1;
    k_corridor_hits<<<(unsigned)n_lines, 128, 0, (cudaStream_t)stream>>>(envs->dev, d_env_id, d_points, (const long long*)d_line_start,
                                                                       n_lines, radius, d_out);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}
