// hl_dubins.cuh -- Dubins shortest path + sampling (device), the `dubins` package of the reference.
//
// The reference imports pydubins (requirements.txt:14, a Cython wrapper of Andrew Walker's dubins.c; un-vendored and
// un-pinned, so PARITY IS UNPINNED here): hybrid_a_star_search.py:294-295 (Pawn goal extension),
// utils/navigation_utils.py:206-215 (get_dubins_path -> get_dubins_path_full, circle-back and Dubins turns).
// These functions restate dubins.c 1.x in its operation order (oracle/dubins_port.py is the CPU twin):
//   shortest_path(q0, q1, rho): words LSL LSR RSL RSR RLR LRL evaluated in that order, first minimum of t+p+q wins
//   sample_many(step): configurations at t = 0, step, 2 step, ... < length (t accumulated by repeated addition)
#pragma once
#include "hl_spline.cuh"

struct DubPath {
    double qi[3];
    double param[3];
    double rho;
    int type;             // 0..5 = LSL LSR RSL RSR RLR LRL, -1 = no path
};

static __device__ __forceinline__ double dub_fmodr(double x, double y) { return xsub(x, xmul(y, floor(xdiv(x, y)))); }
static __device__ __forceinline__ double dub_mod2pi(double t) { return dub_fmodr(t, 2.0 * HL_PI); }

static __device__ __noinline__ void dub_shortest(const double* q0, const double* q1, double rho, DubPath& P) {
    const double dx = xsub(q1[0], q0[0]), dy = xsub(q1[1], q0[1]);
    const double D = sqrt(xadd(xmul(dx, dx), xmul(dy, dy)));
    const double d = xdiv(D, rho);
    double theta = 0.0;
    if (d > 0) theta = dub_mod2pi(m_atan2(dy, dx));
    const double alpha = dub_mod2pi(xsub(q0[2], theta)), beta = dub_mod2pi(xsub(q1[2], theta));
    const double sa = m_sin(alpha), sb = m_sin(beta), ca = m_cos(alpha), cb = m_cos(beta);
    const double c_ab = m_cos(xsub(alpha, beta)), d_sq = xmul(d, d);
    P.qi[0] = q0[0]; P.qi[1] = q0[1]; P.qi[2] = q0[2]; P.rho = rho; P.type = -1;
    double best = INFINITY;
#pragma unroll 1
    for (int w = 0; w < 6; ++w) {
        double o0 = 0.0, o1 = 0.0, o2 = 0.0;
        bool ok = false;
        if (w == 0) {                                        // LSL
            const double tmp0 = xsub(xadd(d, sa), sb);
            const double p_sq = xadd(xsub(xadd(2.0, d_sq), xmul(2.0, c_ab)), xmul(xmul(2.0, d), xsub(sa, sb)));
            if (p_sq >= 0) {
                const double tmp1 = m_atan2(xsub(cb, ca), tmp0);
                o0 = dub_mod2pi(xsub(tmp1, alpha)); o1 = sqrt(p_sq); o2 = dub_mod2pi(xsub(beta, tmp1)); ok = true;
            }
        } else if (w == 1) {                                 // LSR
            const double p_sq = xadd(xadd(xadd(-2.0, d_sq), xmul(2.0, c_ab)), xmul(xmul(2.0, d), xadd(sa, sb)));
            if (p_sq >= 0) {
                const double p = sqrt(p_sq);
                const double tmp0 = xsub(m_atan2(xsub(-ca, cb), xadd(xadd(d, sa), sb)), m_atan2(-2.0, p));
                o0 = dub_mod2pi(xsub(tmp0, alpha)); o1 = p; o2 = dub_mod2pi(xsub(tmp0, dub_mod2pi(beta))); ok = true;
            }
        } else if (w == 2) {                                 // RSL
            const double p_sq = xsub(xadd(xadd(-2.0, d_sq), xmul(2.0, c_ab)), xmul(xmul(2.0, d), xadd(sa, sb)));
            if (p_sq >= 0) {
                const double p = sqrt(p_sq);
                const double tmp0 = xsub(m_atan2(xadd(ca, cb), xsub(xsub(d, sa), sb)), m_atan2(2.0, p));
                o0 = dub_mod2pi(xsub(alpha, tmp0)); o1 = p; o2 = dub_mod2pi(xsub(beta, tmp0)); ok = true;
            }
        } else if (w == 3) {                                 // RSR
            const double tmp0 = xadd(xsub(d, sa), sb);
            const double p_sq = xadd(xsub(xadd(2.0, d_sq), xmul(2.0, c_ab)), xmul(xmul(2.0, d), xsub(sb, sa)));
            if (p_sq >= 0) {
                const double tmp1 = m_atan2(xsub(ca, cb), tmp0);
                o0 = dub_mod2pi(xsub(alpha, tmp1)); o1 = sqrt(p_sq); o2 = dub_mod2pi(xsub(tmp1, beta)); ok = true;
            }
        } else if (w == 4) {                                 // RLR
            const double tmp0 = xdiv(xadd(xadd(xsub(6.0, d_sq), xmul(2.0, c_ab)), xmul(xmul(2.0, d), xsub(sa, sb))), 8.0);
            const double phi = m_atan2(xsub(ca, cb), xadd(xsub(d, sa), sb));
            if (fabs(tmp0) <= 1) {
                const double p = dub_mod2pi(xsub(2.0 * HL_PI, m_acos(tmp0)));
                const double t = dub_mod2pi(xadd(xsub(alpha, phi), dub_mod2pi(xdiv(p, 2.0))));
                o0 = t; o1 = p; o2 = dub_mod2pi(xadd(xsub(xsub(alpha, beta), t), dub_mod2pi(p))); ok = true;
            }
        } else {                                             // LRL
            const double tmp0 = xdiv(xadd(xadd(xsub(6.0, d_sq), xmul(2.0, c_ab)), xmul(xmul(2.0, d), xsub(sb, sa))), 8.0);
            const double phi = m_atan2(xsub(ca, cb), xsub(xadd(d, sa), sb));
            if (fabs(tmp0) <= 1) {
                const double p = dub_mod2pi(xsub(2.0 * HL_PI, m_acos(tmp0)));
                const double t = dub_mod2pi(xadd(xsub(-alpha, phi), xdiv(p, 2.0)));
                o0 = t; o1 = p; o2 = dub_mod2pi(xadd(xsub(xsub(dub_mod2pi(beta), alpha), t), dub_mod2pi(p))); ok = true;
            }
        }
        if (ok) {
            const double cost = xadd(xadd(o0, o1), o2);
            if (cost < best) { best = cost; P.type = w; P.param[0] = o0; P.param[1] = o1; P.param[2] = o2; }
        }
    }
}

static __device__ __forceinline__ double dub_length(const DubPath& P) {
    return xmul(xadd(xadd(P.param[0], P.param[1]), P.param[2]), P.rho);
}

// dubins_segment: kind 0 = L, 1 = S, 2 = R
static __device__ __forceinline__ void dub_segment(double t, const double* qi, int kind, double* qt) {
    const double st = m_sin(qi[2]), ct = m_cos(qi[2]);
    if (kind == 0) { qt[0] = xsub(m_sin(xadd(qi[2], t)), st); qt[1] = xadd(-m_cos(xadd(qi[2], t)), ct); qt[2] = t; }
    else if (kind == 2) { qt[0] = xadd(-m_sin(xsub(qi[2], t)), st); qt[1] = xsub(m_cos(xsub(qi[2], t)), ct); qt[2] = -t; }
    else { qt[0] = xmul(ct, t); qt[1] = xmul(st, t); qt[2] = 0.0; }
    qt[0] = xadd(qt[0], qi[0]); qt[1] = xadd(qt[1], qi[1]); qt[2] = xadd(qt[2], qi[2]);
}

static __device__ __noinline__ void dub_sample(const DubPath& P, double t, double* q) {
    const int kinds[6][3] = {{0, 1, 0}, {0, 1, 2}, {2, 1, 0}, {2, 1, 2}, {2, 0, 2}, {0, 2, 0}};
    const int* ty = kinds[P.type];
    const double tprime = xdiv(t, P.rho);
    const double qi[3] = {0.0, 0.0, P.qi[2]};
    double q1[3], q2[3];
    const double p1 = P.param[0], p2 = P.param[1];
    dub_segment(p1, qi, ty[0], q1);
    dub_segment(p2, q1, ty[1], q2);
    if (tprime < p1) dub_segment(tprime, qi, ty[0], q);
    else if (tprime < xadd(p1, p2)) dub_segment(xsub(tprime, p1), q1, ty[1], q);
    else dub_segment(xsub(xsub(tprime, p1), p2), q2, ty[2], q);
    q[0] = xadd(xmul(q[0], P.rho), P.qi[0]);
    q[1] = xadd(xmul(q[1], P.rho), P.qi[1]);
    q[2] = dub_mod2pi(q[2]);
}

// number of configurations sample_many(step) returns: x = 0; while (x < length) { ...; x += step; }
static __device__ __forceinline__ long long dub_n_samples(double length, double step) {
    long long n = 0;
    for (double x = 0.0; x < length; x = xadd(x, step)) ++n;
    return n;
}

// Knots of the spline course of one Dubins path, by a whole warp: the sample parameters are accumulated by lane 0
// (repeated addition, like sample_many), the configurations are evaluated lane-parallel, then lane 0 drops every
// point that the next one repeats (cubic_spline.py:94-99) and accumulates the chord lengths.  `ts`, `kx`, `ky`, `ks`
// hold at least n_samples + 1 doubles.  Returns the number of knots (uniform), 0 when a spline cannot be built.
static __device__ __noinline__ int dub_course_knots(const DubPath& P, double step, bool append_goal, const double* goal,
                                                    long long n_s, double* ts, double* kx, double* ky, double* ks, int lane,
                                                    double* samples = nullptr) {
    if (lane == 0) {
        double x = 0.0;
        for (long long i = 0; i < n_s; ++i) { ts[i] = x; x = xadd(x, step); }
    }
    __syncwarp();
    for (long long i = lane; i < n_s; i += 32) {
        double q[3];
        dub_sample(P, ts[i], q);
        kx[i] = q[0]; ky[i] = q[1];
        if (samples) { samples[3 * i] = q[0]; samples[3 * i + 1] = q[1]; samples[3 * i + 2] = q[2]; }
    }
    long long n_pts = n_s;
    if (append_goal) { if (lane == 0) { kx[n_s] = goal[0]; ky[n_s] = goal[1]; } n_pts = n_s + 1; }
    __syncwarp();
    int m = 0;
    if (lane == 0) {
        for (long long i = 0; i < n_pts; ++i) {
            if (i + 1 < n_pts && kx[i + 1] == kx[i] && ky[i + 1] == ky[i]) continue;
            const double x = kx[i], y = ky[i];
            kx[m] = x; ky[m] = y;
            ks[m] = (m == 0) ? 0.0 : xadd(ks[m - 1], hypot_cr(xsub(x, kx[m - 1]), xsub(y, ky[m - 1])));
            ++m;
        }
        if (m < 2) m = 0;
    }
    m = __shfl_sync(0xffffffffu, m, 0);
    __syncwarp();
    return m;
}
