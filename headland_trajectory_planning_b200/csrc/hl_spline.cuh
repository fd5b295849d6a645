// hl_spline.cuh -- scipy's not-a-knot CubicSpline over the chord length of a polyline, as the reference uses it
// (path_planner/utils/cubic_spline.py:19-112: Spline2D + calc_spline_course).  Shared by K8 (hl_refpath.cu: warm-start
// path -> OBCA initial guess) and K9 (hl_dubins.cu / the Pawn goal extension: Dubins samples -> spline course).
// float64 throughout; agrees with scipy's pivoted banded solve to ~1e-9.
#pragma once
#include "hl_common.cuh"

// knots of one piece [a, b) of the path: drop pose i when pose i+1 repeats it (cubic_spline.py:94-99),
// chord length s (np.hypot + np.cumsum).  Returns the number of knots.
static __device__ int rp_knots(const double* px, const double* py, long long a, long long b, double* kx, double* ky, double* ks) {
    int m = 0;
    for (long long i = a; i < b; ++i) {
        if (i + 1 < b && px[i + 1] == px[i] && py[i + 1] == py[i]) continue;
        kx[m] = px[i]; ky[m] = py[i];
        ks[m] = (m == 0) ? 0.0 : xadd(ks[m - 1], hypot_cr(xsub(kx[m], kx[m - 1]), xsub(ky[m], ky[m - 1])));
        ++m;
    }
    return m;
}

// len(np.arange(0, s_end + ds, ds))
static __device__ __forceinline__ long long rp_count(double s_end, double ds) {
    return (long long)ceil(xdiv(xadd(s_end, ds), ds));
}

// nodal first derivatives of scipy's CubicSpline(s, y) with bc_type='not-a-knot' for two ordinates at once
static __device__ void rp_derivs(const double* s, const double* x, const double* y, int m, double* dx, double* dy, double* cp,
                          double* bx, double* by) {
    if (m == 2) {
        const double h = xsub(s[1], s[0]);
        dx[0] = dx[1] = xdiv(xsub(x[1], x[0]), h);
        dy[0] = dy[1] = xdiv(xsub(y[1], y[0]), h);
        return;
    }
    if (m == 3) {
        // scipy's special case: the parabola through the three points,
        //   [1 1 0; h1 2(h0+h1) h0; 0 1 1] d = [2 m0; 3 (h0 m1 + h1 m0); 2 m1]
        const double h0 = s[1] - s[0], h1 = s[2] - s[1];
        for (int c = 0; c < 2; ++c) {
            const double* v = c ? y : x;
            double* d = c ? dy : dx;
            const double m0 = (v[1] - v[0]) / h0, m1 = (v[2] - v[1]) / h1;
            const double r0 = 2.0 * m0, r1 = 3.0 * (h0 * m1 + h1 * m0), r2 = 2.0 * m1;
            // eliminate: d0 = r0 - d1, d2 = r2 - d1  ->  d1 (2(h0+h1) - h1 - h0) = r1 - h1 r0 - h0 r2
            const double d1 = (r1 - h1 * r0 - h0 * r2) / (h0 + h1);
            d[0] = r0 - d1; d[1] = d1; d[2] = r2 - d1;
        }
        return;
    }
    // general case: tridiagonal system (scipy/interpolate/_cubic.py), Thomas algorithm
    const int n = m;
    auto h = [&](int i) { return s[i + 1] - s[i]; };
    auto sl = [&](const double* v, int i) { return (v[i + 1] - v[i]) / (s[i + 1] - s[i]); };
    // row 0 (not-a-knot): A00 = h1, A01 = s2 - s0
    {
        const double d = s[2] - s[0];
        const double diag = h(1), up = d;
        bx[0] = ((h(0) + 2.0 * d) * h(1) * sl(x, 0) + h(0) * h(0) * sl(x, 1)) / d;
        by[0] = ((h(0) + 2.0 * d) * h(1) * sl(y, 0) + h(0) * h(0) * sl(y, 1)) / d;
        cp[0] = up / diag; bx[0] /= diag; by[0] /= diag;
    }
    for (int i = 1; i < n - 1; ++i) {
        const double lo = h(i), diag = 2.0 * (h(i - 1) + h(i)), up = h(i - 1);
        const double rx = 3.0 * (h(i) * sl(x, i - 1) + h(i - 1) * sl(x, i));
        const double ry = 3.0 * (h(i) * sl(y, i - 1) + h(i - 1) * sl(y, i));
        const double den = diag - lo * cp[i - 1];
        cp[i] = up / den;
        bx[i] = (rx - lo * bx[i - 1]) / den;
        by[i] = (ry - lo * by[i - 1]) / den;
    }
    {
        // last row (not-a-knot): A[n-1][n-2] = s[n-1] - s[n-3], A[n-1][n-1] = h[n-3]
        const int i = n - 1;
        const double d = s[n - 1] - s[n - 3];
        const double lo = d, diag = h(n - 3);
        const double rx = (h(n - 2) * h(n - 2) * sl(x, n - 3) + (2.0 * d + h(n - 2)) * h(n - 3) * sl(x, n - 2)) / d;
        const double ry = (h(n - 2) * h(n - 2) * sl(y, n - 3) + (2.0 * d + h(n - 2)) * h(n - 3) * sl(y, n - 2)) / d;
        const double den = diag - lo * cp[i - 1];
        bx[i] = (rx - lo * bx[i - 1]) / den;
        by[i] = (ry - lo * by[i - 1]) / den;
    }
    dx[n - 1] = bx[n - 1]; dy[n - 1] = by[n - 1];
    for (int i = n - 2; i >= 0; --i) { dx[i] = bx[i] - cp[i] * dx[i + 1]; dy[i] = by[i] - cp[i] * dy[i + 1]; }
}

// value, first and second derivative of the piecewise cubic at t (PPoly: last interval extrapolates)
static __device__ __forceinline__ void rp_eval(const double* s, const double* v, const double* d, int m, double t, double& f,
                                        double& f1, double& f2) {
    int lo = 0, hi = m - 1;                       // largest j with s[j] <= t, clamped to [0, m-2]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s[mid] <= t) lo = mid; else hi = mid; }
    const int j = lo;
    const double h = s[j + 1] - s[j];
    const double slope = (v[j + 1] - v[j]) / h;
    const double tt = (d[j] + d[j + 1] - 2.0 * slope) / h;
    const double c0 = tt / h, c1 = (slope - d[j]) / h - tt, c2 = d[j], c3 = v[j];
    const double u = t - s[j];
    f = c3 + c2 * u + c1 * u * u + c0 * u * u * u;
    f1 = c2 + 2.0 * c1 * u + 3.0 * c0 * u * u;
    f2 = 2.0 * c1 + 6.0 * c0 * u;
}

