// hl_rs.cuh -- Reeds-Shepp analytic expansion (device).
//
// Replaces path_planner/utils/reeds_shepp.py: the 46 candidate words
// (SCS 2, CSC 8, CCC 8, CCCC 8, CCSC 16, CCSCC 4; :565-582) are one table, one
// thread per row; validity flags, lengths, the signed-sum dedup of set_path
// (:68-87) and the sample COUNT of generate_local_course (:471-530) are float64 in
// the reference's operation order, because they feed discrete decisions.
#pragma once
#include "hl_common.cuh"
#ifdef HL_SHARED_CODE
#define HL_CODE2 __noinline__
#else
#define HL_CODE2
#endif

enum { RS_SLS = 0, RS_LSL, RS_LSR, RS_LRL, RS_LRLRN, RS_LRLRP, RS_LRSR, RS_LRSL, RS_LRSLR };
enum { RP_TUV = 0, RP_VUT, RP_TUnUV, RP_TUUV, RP_THUV, RP_VUHT, RP_THUHV };
enum { RS_S = 0, RS_L = 1, RS_R = 2 };

// row = solver | sx<0 (bit 4) | sy<0 (bit 5) | backwards (bit 6) | pattern << 8 | nseg << 12 | letters << 16 (2 bits each)
struct RsRow { unsigned char solver, neg_x, neg_y, backwards, pattern, nseg; unsigned short letters; };

#define RS_LET3(a, b, c) ((a) | ((b) << 2) | ((c) << 4))
#define RS_LET4(a, b, c, d) (RS_LET3(a, b, c) | ((d) << 6))
#define RS_LET5(a, b, c, d, e) (RS_LET4(a, b, c, d) | ((e) << 8))
#define RS_QUAD(solver, back, pat, n, pos, neg)                                     \
    {solver, 0, 0, back, pat, n, pos}, {solver, 1, 0, back, pat, n, pos},           \
    {solver, 0, 1, back, pat, n, neg}, {solver, 1, 1, back, pat, n, neg}

static __constant__ RsRow c_rs_rows[HL_RS_CANDIDATES] = {
    {RS_SLS, 0, 0, 0, RP_TUV, 3, RS_LET3(RS_S, RS_L, RS_S)},
    {RS_SLS, 0, 1, 0, RP_TUV, 3, RS_LET3(RS_S, RS_R, RS_S)},
    RS_QUAD(RS_LSL, 0, RP_TUV, 3, RS_LET3(RS_L, RS_S, RS_L), RS_LET3(RS_R, RS_S, RS_R)),
    RS_QUAD(RS_LSR, 0, RP_TUV, 3, RS_LET3(RS_L, RS_S, RS_R), RS_LET3(RS_R, RS_S, RS_L)),
    RS_QUAD(RS_LRL, 0, RP_TUV, 3, RS_LET3(RS_L, RS_R, RS_L), RS_LET3(RS_R, RS_L, RS_R)),
    RS_QUAD(RS_LRL, 1, RP_VUT, 3, RS_LET3(RS_L, RS_R, RS_L), RS_LET3(RS_R, RS_L, RS_R)),
    RS_QUAD(RS_LRLRN, 0, RP_TUnUV, 4, RS_LET4(RS_L, RS_R, RS_L, RS_R), RS_LET4(RS_R, RS_L, RS_R, RS_L)),
    RS_QUAD(RS_LRLRP, 0, RP_TUUV, 4, RS_LET4(RS_L, RS_R, RS_L, RS_R), RS_LET4(RS_R, RS_L, RS_R, RS_L)),
    RS_QUAD(RS_LRSL, 0, RP_THUV, 4, RS_LET4(RS_L, RS_R, RS_S, RS_L), RS_LET4(RS_R, RS_L, RS_S, RS_R)),
    RS_QUAD(RS_LRSR, 0, RP_THUV, 4, RS_LET4(RS_L, RS_R, RS_S, RS_R), RS_LET4(RS_R, RS_L, RS_S, RS_L)),
    RS_QUAD(RS_LRSL, 1, RP_VUHT, 4, RS_LET4(RS_L, RS_S, RS_R, RS_L), RS_LET4(RS_R, RS_S, RS_L, RS_R)),
    RS_QUAD(RS_LRSR, 1, RP_VUHT, 4, RS_LET4(RS_R, RS_S, RS_R, RS_L), RS_LET4(RS_L, RS_S, RS_L, RS_R)),
    RS_QUAD(RS_LRSLR, 0, RP_THUHV, 5, RS_LET5(RS_L, RS_R, RS_S, RS_L, RS_R), RS_LET5(RS_R, RS_L, RS_S, RS_R, RS_L)),
};

// candidate row of (pass, lane) when a warp evaluates the 46 rows in two passes with DISJOINT solver sets
// (pass 0 = SLS, LSL, LRL, LRSL, LRSR rows; pass 1 = LSR, LRLRN, LRLRP, LRSLR rows): the lanes of a pass diverge
// over the solver switch, so each formula's code then runs once per pair instead of once per pass
static __constant__ signed char c_rs_pass_cand[2][32] = {
    {0, 1, 2, 3, 4, 5, 10, 11, 12, 13, 14, 15, 16, 17, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, -1, -1},
    {6, 7, 8, 9, 18, 19, 20, 21, 22, 23, 24, 25, 42, 43, 44, 45, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1},
};

__device__ __forceinline__ int rs_letter(unsigned short letters, int i) { return (letters >> (2 * i)) & 3; }

// ---- word solvers: float64, reference operation order -------------------------
// Each returns validity and (t,u,v).  Products/sums that the reference evaluates
// as separate Python float operations use xmul/xadd so no FMA is formed.
__device__ __forceinline__ void rs_polar(double x, double y, double& r, double& th) {
    r = hypot_cr(x, y);
    th = m_atan2(y, x);
}

// sphi, cphi = sine and cosine of the (already mirrored) phi: sin is odd and cos even bit-for-bit,
// so the caller derives them from one sincos per pose pair.
static __device__ HL_CODE2 bool rs_solve(int solver, double x, double y, double phi, double sphi, double cphi,
                                double& t, double& u, double& v) {
    const double PI = HL_PI;
    switch (solver) {
    case RS_SLS: {                                              // reeds_shepp.py:144-160
        phi = rs_mod2pi(phi);
        if ((y > 0.0 || y < 0.0) && 0.0 < phi && phi < xmul(PI, 0.99)) {
            double tn = m_tan(phi);
            double xd = xadd(xdiv(-y, tn), x);
            double th = m_tan(xdiv(phi, 2.0));
            t = xsub(xd, th);
            u = phi;
            double dx = xsub(x, xd);
            double rt = sqrt(xadd(xmul(dx, dx), xmul(y, y)));
            v = (y > 0.0) ? xsub(rt, th) : xsub(-rt, th);
            return true;
        }
        return false;
    }
    case RS_LSL: {                                              // :90-98
        double r, th;
        rs_polar(xsub(x, sphi), xadd(xsub(y, 1.0), cphi), r, th);
        u = r; t = th;
        if (t >= 0.0) {
            v = rs_mod2pi(xsub(phi, t));
            if (v >= 0.0) return true;
        }
        return false;
    }
    case RS_LSR: {                                              // :101-114
        double u1, t1;
        rs_polar(xadd(x, sphi), xsub(xsub(y, 1.0), cphi), u1, t1);
        u1 = xmul(u1, u1);
        if (u1 >= 4.0) {
            u = sqrt(xsub(u1, 4.0));
            double theta = m_atan2(2.0, u);
            t = rs_mod2pi(xadd(t1, theta));
            v = rs_mod2pi(xsub(t, phi));
            if (t >= 0.0 && v >= 0.0) return true;
        }
        return false;
    }
    case RS_LRL: {                                              // :117-128
        double u1, t1;
        rs_polar(xsub(x, sphi), xadd(xsub(y, 1.0), cphi), u1, t1);
        if (u1 <= 4.0) {
            u = xmul(-2.0, m_asin(xmul(0.25, u1)));
            t = rs_mod2pi(xadd(xadd(t1, xmul(0.5, u)), PI));
            v = rs_mod2pi(xadd(xsub(phi, t), u));
            if (t >= 0.0 && u <= 0.0) return true;
        }
        return false;
    }
    case RS_LRLRN:
    case RS_LRLRP: {                                            // :239-283
        double xi = xadd(x, sphi);
        double eta = xsub(xsub(y, 1.0), cphi);
        double uu, vv;
        if (solver == RS_LRLRN) {
            double rho = xmul(0.25, xadd(2.0, sqrt(xadd(xmul(xi, xi), xmul(eta, eta)))));
            if (!(rho <= 1.0)) return false;
            uu = m_acos(rho);
            vv = -uu;
        } else {
            double rho = xdiv(xsub(xsub(20.0, xmul(xi, xi)), xmul(eta, eta)), 16.0);
            if (!(0.0 <= rho && rho <= 1.0)) return false;
            uu = -m_acos(rho);
            if (!(uu >= xmul(-0.5, PI))) return false;
            vv = uu;
        }
        // calc_tauOmega(u, v, xi, eta, phi)
        double delta = rs_mod2pi(xsub(uu, vv));
        double A = xsub(m_sin(uu), m_sin(delta));
        double B = xsub(xsub(m_cos(uu), m_cos(delta)), 1.0);
        double t1 = m_atan2(xsub(xmul(eta, A), xmul(xi, B)), xadd(xmul(xi, A), xmul(eta, B)));
        double t2 = xadd(xmul(2.0, xsub(xsub(m_cos(delta), m_cos(vv)), m_cos(uu))), 3.0);
        double tau = (t2 < 0) ? rs_mod2pi(xadd(t1, PI)) : rs_mod2pi(t1);
        double omega = rs_mod2pi(xsub(xadd(xsub(tau, uu), vv), phi));
        t = tau; u = uu; v = omega;
        if (solver == RS_LRLRN) return t >= 0.0 && v <= 0.0;
        return t >= 0.0 && v >= 0.0;
    }
    case RS_LRSR: {                                             // :322-334
        double xi = xadd(x, sphi);
        double eta = xsub(xsub(y, 1.0), cphi);
        double rho, theta;
        rs_polar(-eta, xi, rho, theta);
        if (rho >= 2.0) {
            t = theta;
            u = xsub(2.0, rho);
            v = rs_mod2pi(xsub(xadd(t, xmul(0.5, PI)), phi));
            if (t >= 0.0 && u <= 0.0 && v <= 0.0) return true;
        }
        return false;
    }
    case RS_LRSL: {                                             // :337-350
        double xi = xsub(x, sphi);
        double eta = xadd(xsub(y, 1.0), cphi);
        double rho, theta;
        rs_polar(xi, eta, rho, theta);
        if (rho >= 2.0) {
            double r = sqrt(xsub(xmul(rho, rho), 4.0));
            u = xsub(2.0, r);
            t = rs_mod2pi(xadd(theta, m_atan2(r, -2.0)));
            v = rs_mod2pi(xsub(xsub(phi, xmul(0.5, PI)), t));
            if (t >= 0.0 && u <= 0.0 && v <= 0.0) return true;
        }
        return false;
    }
    default: {                                                  // RS_LRSLR :425-440
        double xi = xadd(x, sphi);
        double eta = xsub(xsub(y, 1.0), cphi);
        double rho, theta;
        rs_polar(xi, eta, rho, theta);
        if (rho >= 2.0) {
            u = xsub(4.0, sqrt(xsub(xmul(rho, rho), 4.0)));
            if (u <= 0.0) {
                double num = xsub(xmul(xsub(4.0, u), xi), xmul(2.0, eta));
                double den = xadd(xmul(-2.0, xi), xmul(xsub(u, 4.0), eta));
                t = rs_mod2pi(m_atan2(num, den));
                v = rs_mod2pi(xsub(t, phi));
                if (t >= 0.0 && v >= 0.0) return true;
            }
        }
        return false;
    }
    }
}

// Normalised problem of one pose pair (generate_path, :565-572)
struct RsProblem { double x, y, phi, xb, yb, sp, cp; };

static __device__ HL_CODE RsProblem rs_normalise(const double* q0, const double* q1, double maxc) {
    RsProblem P;
    double dx = xsub(q1[0], q0[0]), dy = xsub(q1[1], q0[1]);
    P.phi = xsub(q1[2], q0[2]);
    double c = m_cos(q0[2]), s = m_sin(q0[2]);
    P.x = xmul(xadd(xmul(c, dx), xmul(s, dy)), maxc);
    P.y = xmul(xadd(xmul(-s, dx), xmul(c, dy)), maxc);
    double cp = m_cos(P.phi), sp = m_sin(P.phi);                    // :217-218 / :387-388
    P.sp = sp; P.cp = cp;
    P.xb = xadd(xmul(P.x, cp), xmul(P.y, sp));
    P.yb = xsub(xmul(P.x, sp), xmul(P.y, cp));
    return P;
}

// Evaluate candidate row `cand`: returns validity, writes nseg normalised lengths.
static __device__ HL_CODE2 bool rs_candidate(int cand, const RsProblem& P, double* lens) {
    const RsRow row = c_rs_rows[cand];
    double ax = row.backwards ? P.xb : P.x, ay = row.backwards ? P.yb : P.y;
    if (row.neg_x) ax = -ax;
    if (row.neg_y) ay = -ay;
    const bool negphi = row.neg_x != row.neg_y;
    double aphi = negphi ? -P.phi : P.phi;
    double t, u, v;
    if (!rs_solve(row.solver, ax, ay, aphi, negphi ? -P.sp : P.sp, P.cp, t, u, v)) return false;
    const double H = xmul(-0.5, HL_PI);
    switch (row.pattern) {
    case RP_TUV:   lens[0] = t; lens[1] = u; lens[2] = v; break;
    case RP_VUT:   lens[0] = v; lens[1] = u; lens[2] = t; break;
    case RP_TUnUV: lens[0] = t; lens[1] = u; lens[2] = -u; lens[3] = v; break;
    case RP_TUUV:  lens[0] = t; lens[1] = u; lens[2] = u; lens[3] = v; break;
    case RP_THUV:  lens[0] = t; lens[1] = H; lens[2] = u; lens[3] = v; break;
    case RP_VUHT:  lens[0] = v; lens[1] = u; lens[2] = H; lens[3] = t; break;
    default:       lens[0] = t; lens[1] = H; lens[2] = u; lens[3] = H; lens[4] = v; break;
    }
    if (row.neg_x)
        HL_LOOP
        for (int i = 0; i < row.nseg; ++i) lens[i] = -lens[i];
    return true;
}

// The 46 candidate rows of generate_path (reeds_shepp.py:565-582) in two warp-wide passes built around UNIFORM calls of
// the float64 transcendentals.  rs_solve's switch makes the lanes of a pass run every formula one after the other
// (~3700 dependent instructions per shot); here a pass is a fixed sequence of slots -- polar, atan2 / asin, the two
// M() folds; in pass B also acos, sincos, tan -- that all lanes enter together with their own arguments, and the
// solver-specific arithmetic between the slots is a few selects.  Pass A: LSL, LSR, LRL, LRSL, LRSR rows (32 lanes, all
// start with R(x -+ sin phi, y - 1 +- cos phi)).  Pass B: LRLRn / LRLRp on lanes 0-7, LRSLR on 8-11, SLS on 12-13; lanes
// 16-23 and 28-29 are HELPERS of lanes 0-7 / 12-13: they repeat the cheap prefix and take the second transcendental of
// the same kind (sincos(delta) next to sincos(u); tan(phi/2) next to tan(phi)).  Every value is produced by the same
// float64 operations in the same order as rs_solve / rs_candidate (cos(v) = cos(u) for v = +-u: cos is even bit for bit).
static __constant__ signed char c_rs_uni_pass_a[32] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17,
                                                   26, 27, 28, 29, 34, 35, 36, 37, 30, 31, 32, 33, 38, 39, 40, 41};
static __constant__ signed char c_rs_uni_pass_b[16] = {18, 19, 20, 21, 22, 23, 24, 25, 42, 43, 44, 45, 0, 1, -1, -1};

__device__ __forceinline__ void rs_store_candidate(unsigned char* valid, double (*lens)[HL_RS_MAX_SEGS], int c, const RsRow& row, bool ok,
                                                   double t, double u, double v) {
    double l[HL_RS_MAX_SEGS] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (ok) {
        const double H = xmul(-0.5, HL_PI);
        switch (row.pattern) {
        case RP_TUV:   l[0] = t; l[1] = u; l[2] = v; break;
        case RP_VUT:   l[0] = v; l[1] = u; l[2] = t; break;
        case RP_TUnUV: l[0] = t; l[1] = u; l[2] = -u; l[3] = v; break;
        case RP_TUUV:  l[0] = t; l[1] = u; l[2] = u; l[3] = v; break;
        case RP_THUV:  l[0] = t; l[1] = H; l[2] = u; l[3] = v; break;
        case RP_VUHT:  l[0] = v; l[1] = u; l[2] = H; l[3] = t; break;
        default:       l[0] = t; l[1] = H; l[2] = u; l[3] = H; l[4] = v; break;
        }
        if (row.neg_x)
#pragma unroll
            for (int k = 0; k < HL_RS_MAX_SEGS; ++k) l[k] = (k < row.nseg) ? -l[k] : l[k];
    }
    valid[c] = ok ? 1 : 0;
#pragma unroll
    for (int k = 0; k < HL_RS_MAX_SEGS; ++k) lens[c][k] = l[k];
}

static __device__ __noinline__ void rs_candidates_warp(const RsProblem& Pin, unsigned char* valid, double (*lens)[HL_RS_MAX_SEGS], int lane) {
    const double PI = HL_PI;
    const RsProblem Pb = Pin;
    const unsigned FULL_ = 0xffffffffu;
    {   // ------------------------------------------------ pass A
        const int c = c_rs_uni_pass_a[lane];
        const RsRow row = c_rs_rows[c];
        double x = row.backwards ? Pb.xb : Pb.x, y = row.backwards ? Pb.yb : Pb.y;
        if (row.neg_x) x = -x;
        if (row.neg_y) y = -y;
        const bool negphi = row.neg_x != row.neg_y;
        const double phi = negphi ? -Pb.phi : Pb.phi, sphi = negphi ? -Pb.sp : Pb.sp, cphi = Pb.cp;
        const int solver = row.solver;
        const bool plus = solver == RS_LSR || solver == RS_LRSR;
        const double xi = plus ? xadd(x, sphi) : xsub(x, sphi);
        const double eta = plus ? xsub(xsub(y, 1.0), cphi) : xadd(xsub(y, 1.0), cphi);
        const double X = (solver == RS_LRSR) ? -eta : xi, Y = (solver == RS_LRSR) ? xi : eta;
        const double r = hypot_cr(X, Y);                     // slot: R(X, Y)
        const double th = m_atan2(Y, X);
        bool ok = true;
        double u = r, r2 = 0.0;
        if (solver == RS_LSR) { const double u1 = xmul(r, r); ok = u1 >= 4.0; u = sqrt(xsub(u1, 4.0)); }
        else if (solver == RS_LRSL) { ok = r >= 2.0; r2 = sqrt(xsub(xmul(r, r), 4.0)); u = xsub(2.0, r2); }
        else if (solver == RS_LRSR) { ok = r >= 2.0; u = xsub(2.0, r); }
        else if (solver == RS_LRL) ok = r <= 4.0;
        double aux = 0.0;
        if (solver == RS_LSR || solver == RS_LRSL)           // slot: atan2(2, u) / atan2(r, -2)
            aux = m_atan2(solver == RS_LSR ? 2.0 : r2, solver == RS_LSR ? u : -2.0);
        if (solver == RS_LRL) u = xmul(-2.0, m_asin(xmul(0.25, r)));       // slot: asin
        const bool tmod = solver == RS_LSR || solver == RS_LRL || solver == RS_LRSL;
        const double targ = (solver == RS_LRL) ? xadd(xadd(th, xmul(0.5, u)), PI) : xadd(th, aux);
        const double tm = rs_mod2pi(tmod ? targ : 0.0);      // slot: M(t)
        const double t = tmod ? tm : th;
        double varg;
        if (solver == RS_LSL) varg = xsub(phi, t);
        else if (solver == RS_LSR) varg = xsub(t, phi);
        else if (solver == RS_LRL) varg = xadd(xsub(phi, t), u);
        else if (solver == RS_LRSL) varg = xsub(xsub(phi, xmul(0.5, PI)), t);
        else varg = xsub(xadd(t, xmul(0.5, PI)), phi);
        const double v = rs_mod2pi(varg);                    // slot: M(v)
        if (solver == RS_LSL || solver == RS_LSR) ok = ok && t >= 0.0 && v >= 0.0;
        else if (solver == RS_LRL) ok = ok && t >= 0.0 && u <= 0.0;
        else ok = ok && t >= 0.0 && u <= 0.0 && v <= 0.0;
        rs_store_candidate(valid, lens, c, row, ok, t, u, v);
    }
    __syncwarp();
    {   // ------------------------------------------------ pass B
        const bool helper = lane >= 16;
        const int c = c_rs_uni_pass_b[lane & 15];
        const bool have = c >= 0;
        const RsRow row = c_rs_rows[have ? c : 0];
        double x = row.backwards ? Pb.xb : Pb.x, y = row.backwards ? Pb.yb : Pb.y;
        if (row.neg_x) x = -x;
        if (row.neg_y) y = -y;
        const bool negphi = row.neg_x != row.neg_y;
        const double phi = negphi ? -Pb.phi : Pb.phi, sphi = negphi ? -Pb.sp : Pb.sp, cphi = Pb.cp;
        const int solver = have ? (int)row.solver : -1;
        const bool lrlr = solver == RS_LRLRN || solver == RS_LRLRP, isn = solver == RS_LRLRN;
        const bool sls = solver == RS_SLS, lrslr = solver == RS_LRSLR;
        const double xi = xadd(x, sphi);
        const double eta = xsub(xsub(y, 1.0), cphi);
        const double r = hypot_cr(xi, eta);                  // slot: rho of R(xi, eta) (LRSLR; its theta is not used)
        bool ok = have;
        double rho = 0.5;
        if (isn) { rho = xmul(0.25, xadd(2.0, sqrt(xadd(xmul(xi, xi), xmul(eta, eta))))); ok = rho <= 1.0; }
        else if (lrlr) { rho = xdiv(xsub(xsub(20.0, xmul(xi, xi)), xmul(eta, eta)), 16.0); ok = 0.0 <= rho && rho <= 1.0; }
        const double ac = m_acos(lrlr ? rho : 0.5);          // slot: acos
        const double uu = isn ? ac : -ac, vv = isn ? -uu : uu;
        if (lrlr && !isn) ok = ok && uu >= xmul(-0.5, PI);
        const double m1 = rs_mod2pi(sls ? phi : xsub(uu, vv));   // slot: M(phi) (SLS) / delta = M(u - v)
        double sn, cs;
        m_sincos(lrlr ? (helper ? m1 : uu) : 0.0, &sn, &cs);    // slot: sincos(u) | sincos(delta)
        const double osn = __shfl_xor_sync(FULL_, sn, 16), ocs = __shfl_xor_sync(FULL_, cs, 16);
        const double s_uu = helper ? osn : sn, c_uu = helper ? ocs : cs, s_de = helper ? sn : osn, c_de = helper ? cs : ocs;
        const double tn = m_tan(sls ? (helper ? xdiv(m1, 2.0) : m1) : 0.0);   // slot: tan(phi) | tan(phi / 2)
        const double otn = __shfl_xor_sync(FULL_, tn, 16);
        const double tan_phi = helper ? otn : tn, tan_half = helper ? tn : otn;
        double a_y = 0.0, a_x = 1.0, u = 0.0;
        if (lrlr) {
            const double A = xsub(s_uu, s_de);
            const double B = xsub(xsub(c_uu, c_de), 1.0);
            a_y = xsub(xmul(eta, A), xmul(xi, B)); a_x = xadd(xmul(xi, A), xmul(eta, B));
            u = uu;
        } else if (lrslr) {
            ok = ok && r >= 2.0;
            u = xsub(4.0, sqrt(xsub(xmul(r, r), 4.0)));
            ok = ok && u <= 0.0;
            a_y = xsub(xmul(xsub(4.0, u), xi), xmul(2.0, eta)); a_x = xadd(xmul(-2.0, xi), xmul(xsub(u, 4.0), eta));
        }
        const double at = m_atan2(a_y, a_x);                 // slot: atan2
        double targ = at;
        if (lrlr) {
            const double t2 = xadd(xmul(2.0, xsub(xsub(c_de, c_uu), c_uu)), 3.0);     // cos(v) = cos(u)
            if (t2 < 0) targ = xadd(at, PI);
        }
        double t = rs_mod2pi(targ);                          // slot: M(t)
        const double varg = lrlr ? xsub(xadd(xsub(t, uu), vv), phi) : xsub(t, phi);
        double v = rs_mod2pi(varg);                          // slot: M(v)
        if (lrlr) ok = ok && t >= 0.0 && (isn ? v <= 0.0 : v >= 0.0);
        else if (lrslr) ok = ok && t >= 0.0 && v >= 0.0;
        else if (sls) {
            ok = (y > 0.0 || y < 0.0) && 0.0 < m1 && m1 < xmul(PI, 0.99);
            const double xd = xadd(xdiv(-y, tan_phi), x);
            t = xsub(xd, tan_half);
            u = m1;
            const double dx = xsub(x, xd);
            const double rt = sqrt(xadd(xmul(dx, dx), xmul(y, y)));
            v = (y > 0.0) ? xsub(rt, tan_half) : xsub(-rt, tan_half);
        }
        if (have && !helper) rs_store_candidate(valid, lens, c, row, ok, t, u, v);
    }
    __syncwarp();
}

// set_path (:68-87).  A candidate is only ever compared with EARLIER accepted candidates of the same
// letters, so the 46 rows split into 20 independent letter groups (<= 4 rows each, listed in evaluation
// order).  rs_select_group decides one group; rs_select_compact then lists the accepted rows in
// evaluation order.  accept[c]: 0 rejected / invalid, 1 accepted, 2 = the reference's
// `assert path.L >= 0.01` would fire.
#define RS_N_GROUPS 20
static __constant__ signed char c_rs_groups[RS_N_GROUPS][4] = {
    {0, -1, -1, -1}, {1, -1, -1, -1},                       // SLS, SRS
    {2, 3, -1, -1}, {4, 5, -1, -1},                         // LSL, RSR
    {6, 7, -1, -1}, {8, 9, -1, -1},                         // LSR, RSL
    {10, 11, 14, 15}, {12, 13, 16, 17},                     // LRL, RLR
    {18, 19, 22, 23}, {20, 21, 24, 25},                     // LRLR, RLRL
    {26, 27, -1, -1}, {28, 29, -1, -1},                     // LRSL, RLSR
    {30, 31, -1, -1}, {32, 33, -1, -1},                     // LRSR, RLSL
    {34, 35, -1, -1}, {36, 37, -1, -1},                     // LSRL, RSLR
    {38, 39, -1, -1}, {40, 41, -1, -1},                     // RSRL, LSLR
    {42, 43, -1, -1}, {44, 45, -1, -1},                     // LRSLR, RLSRL
};

static __device__ HL_CODE2 void rs_select_group(int g, const unsigned char* valid, const double (*lens)[HL_RS_MAX_SEGS],
                                       unsigned char* accept, double* Lc) {
    int kept[4];
    int nk = 0;
    HL_LOOP
    for (int q = 0; q < 4; ++q) {
        const int c = c_rs_groups[g][q];
        if (c < 0) break;
        accept[c] = 0;
        if (!valid[c]) continue;
        const int nseg = c_rs_rows[c].nseg;
        bool dup = false;
        HL_LOOP
        for (int k = 0; k < nk && !dup; ++k) {
            double s = 0.0;                                  // Python sum(): 0 + d0 + d1 + ...
            HL_LOOP
            for (int i = 0; i < nseg; ++i) s = xadd(s, xsub(lens[kept[k]][i], lens[c][i]));
            if (s <= 0.01) dup = true;
        }
        if (dup) continue;
        double tot = 0.0;
        HL_LOOP
        for (int i = 0; i < nseg; ++i) tot = xadd(tot, fabs(lens[c][i]));
        if (tot >= 1000.0) continue;                         // MAX_LENGTH
        Lc[c] = tot;
        if (!(tot >= 0.01)) { accept[c] = 2; continue; }
        accept[c] = 1;
        kept[nk++] = c;
    }
}

// accepted rows in evaluation order; returns their number or -1 when the assert would fire
static __device__ HL_CODE2 int rs_select_compact(const unsigned char* accept, const double* Lc, int* acc, double* L) {
    int n = 0;
    for (int c = 0; c < HL_RS_CANDIDATES; ++c) {
        if (accept[c] == 2) return -1;
        if (accept[c] == 1) { acc[n] = c; L[n] = Lc[c]; ++n; }
    }
    return n;
}

static __device__ int rs_select(const unsigned char* valid, const double (*lens)[HL_RS_MAX_SEGS], int* acc, double* L) {
    unsigned char accept[HL_RS_CANDIDATES];
    double Lc[HL_RS_CANDIDATES];
    for (int g = 0; g < RS_N_GROUPS; ++g) rs_select_group(g, valid, lens, accept, Lc);
    return rs_select_compact(accept, Lc, acc, L);
}

// calculate_reeds_shepp_path_cost (hybrid_a_star_search.py:129-160) with the quirks:
// +DIRECTION_CHANGE_COST and +MAX_STEER always (len(np.where(..)) == 1); 'L' arcs steer 0.
// Lengths here may be normalised or metric: only signs matter.
static __device__ HL_CODE2 double rs_path_cost(double node_cost, int cand, const double* lens, double max_steer,
                               double reverse_cost, double dir_change_cost, double steer_cost) {
    const RsRow row = c_rs_rows[cand];
    int nneg = 0;
    HL_LOOP
    for (int i = 0; i < row.nseg; ++i) nneg += (lens[i] < 0.0) ? 1 : 0;
    double cost = node_cost;
    cost = xadd(cost, xadd(xmul(reverse_cost, (double)nneg), (double)(row.nseg - nneg)));
    cost = xadd(cost, xmul(1.0, dir_change_cost));
    cost = xadd(cost, xmul(xmul(max_steer, steer_cost), 1.0));
    double prev = 0.0, sum = 0.0;
    HL_LOOP
    for (int i = 0; i < row.nseg; ++i) {
        double st = (rs_letter(row.letters, i) == RS_R) ? -max_steer : 0.0;
        if (i > 0) sum = xadd(sum, fabs(xsub(st, prev)));        // np.sum of <= 4 elements: sequential
        prev = st;
    }
    return xadd(cost, sum);
}

// heapdict pop order of n entries inserted in index order with the given priorities
// (hybrid_a_star_search.py:265-271).  order[] receives the indices in pop order.
static __device__ HL_CODE2 void heapdict_order(const double* prio, int n, int* order) {
    int heap[HL_RS_CANDIDATES];
    int m = 0;
    HL_LOOP
    for (int k = 0; k < n; ++k) {                 // __setitem__: append + _decrease_key
        int i = m++;
        heap[i] = k;
        HL_LOOP
        while (i) {
            int parent = (i - 1) >> 1;
            if (prio[heap[parent]] < prio[heap[i]]) break;
            int tmp = heap[i]; heap[i] = heap[parent]; heap[parent] = tmp;
            i = parent;
        }
    }
    HL_LOOP
    for (int k = 0; k < n; ++k) {                 // popitem: move last to root + _min_heapify
        order[k] = heap[0];
        --m;
        if (m > 0) {
            heap[0] = heap[m];
            int i = 0;
            HL_LOOP
            while (true) {
                int l = (i << 1) + 1, r = (i + 1) << 1, low = i;
                if (l < m && prio[heap[l]] < prio[heap[i]]) low = l;
                if (r < m && prio[heap[r]] < prio[heap[low]]) low = r;
                if (low == i) break;
                int tmp = heap[i]; heap[i] = heap[low]; heap[low] = tmp;
                i = low;
            }
        }
    }
}

// ---- sampler: generate_local_course + interpolate (:471-562) -----------------
struct RsSegPlan {
    double ox, oy, oyaw;     // origin pose of the segment (local frame of the start pose)
    double pd0, d, l;        // first offset, step, signed normalised length
    int first, count, letter;  // first emitted index, loop samples, letter
    // float32 view for the collision filter (rs_plan_world32): segment origin in the environment's
    // float32 frame and cos/sin of the world heading there
    float fox, foy, fc0, fs0;
};
struct RsPlan {
    RsSegPlan seg[HL_RS_MAX_SEGS];
    int nseg;
    int npts;                // after the trailing `px == 0.0` pops
    int dir0;
};

static __device__ HL_CODE void rs_interp(double l, int letter, double maxc, double ox, double oy, double oyaw,
                                          double& px, double& py, double& pyaw) {
    if (letter == RS_S) {
        double lm = xdiv(l, maxc);
        px = xadd(ox, xmul(lm, m_cos(oyaw)));
        py = xadd(oy, xmul(lm, m_sin(oyaw)));
        pyaw = oyaw;
    } else {
        double ldx = xdiv(m_sin(l), maxc);
        double ldy = (letter == RS_L) ? xdiv(xsub(1.0, m_cos(l)), maxc) : xdiv(xsub(1.0, m_cos(l)), -maxc);
        double cn = m_cos(-oyaw), sn = m_sin(-oyaw);
        double gdx = xadd(xmul(cn, ldx), xmul(sn, ldy));
        double gdy = xadd(xmul(-sn, ldx), xmul(cn, ldy));
        px = xadd(ox, gdx);
        py = xadd(oy, gdy);
        pyaw = (letter == RS_L) ? xadd(oyaw, l) : xsub(oyaw, l);
    }
}

// Build the per-segment plan of one word.  `lens` normalised, step = step_size*maxc.
// The loop offsets are accumulated by repeated addition exactly like the reference, so
// the sample count is bit-exact.
static __device__ HL_CODE2 void rs_make_plan(int cand, const double* lens, double maxc, double step, RsPlan& P) {
    const RsRow row = c_rs_rows[cand];
    P.nseg = row.nseg;
    P.dir0 = (lens[0] > 0.0) ? 1 : -1;
    double ox = 0.0, oy = 0.0, oyaw = 0.0;        // px[1] of the zero-initialised arrays
    double ll = 0.0;
    int ind = 1;
    HL_LOOP
    for (int i = 0; i < row.nseg; ++i) {
        double l = lens[i];
        double d = (l > 0.0) ? step : -step;
        RsSegPlan& S = P.seg[i];
        S.ox = ox; S.oy = oy; S.oyaw = oyaw; S.l = l; S.d = d; S.letter = rs_letter(row.letters, i);
        ind -= 1;
        double pd;
        if (i >= 1 && xmul(lens[i - 1], lens[i]) > 0.0) pd = xsub(-d, ll);
        else pd = xsub(d, ll);
        S.pd0 = pd;
        S.first = ind + 1;
        // loop count of `while abs(pd) <= abs(l): pd += d`.  Along the step direction the offsets are
        // a + k*|d| with a = +-pd, so the count is floor((|l| - a)/|d|) + 1 unless the boundary is within
        // 1e-7 of a sample, where the reference's repeated addition is replayed exactly.
        int cnt = 0;
        const double al = fabs(l);
        if (fabs(pd) <= al) {
            const double a = (d > 0.0) ? pd : -pd;
            const double r = (al - a) / step;
            const double kf = floor(r);
            if (r - kf > 1e-7 && kf + 1.0 - r > 1e-7 && r < 1e7) {
                cnt = (int)kf + 1;
                pd = xadd(pd, xmul((double)cnt, d));
            } else {
                HL_LOOP
                while (fabs(pd) <= al) { ++cnt; pd = xadd(pd, d); }
            }
        }
        S.count = cnt;
        ind += cnt;
        ll = xsub(xsub(l, pd), d);
        ind += 1;                                 // segment end point at `ind`
        rs_interp(l, S.letter, maxc, S.ox, S.oy, S.oyaw, ox, oy, oyaw);
    }
    // `ind` is the index of the final end point; points 0..ind are written.
    int npts = ind + 1;
    // trailing pops: while px[-1] == 0.0 (:523-528).  The allocated tail beyond `ind`
    // is zero and always popped; written samples are popped only if their x is exactly 0.
    // last written point = end of the last segment = (ox, oy, oyaw) computed above
    if (ox == 0.0) {
        npts -= 1;
        // walk back through loop samples while their x is exactly 0.0
        HL_LOOP
        while (npts > 1) {
            int j = npts - 1;
            int si = P.nseg - 1;
            HL_LOOP
            while (si > 0 && j < P.seg[si].first) --si;
            const RsSegPlan& S = P.seg[si];
            double pd = S.pd0;
            HL_LOOP
            for (int k = 0; k < j - S.first; ++k) pd = xadd(pd, S.d);
            double px, py, pyaw;
            rs_interp(pd, S.letter, maxc, S.ox, S.oy, S.oyaw, px, py, pyaw);
            if (px != 0.0) break;
            npts -= 1;
        }
        // index 0 is (0,0,0): popping it too would raise IndexError in the reference
    }
    P.npts = npts;
}

// Pose j (0 <= j < npts) of a planned word in the LOCAL frame, plus curvature sign and
// direction tag.  Loop offsets use pd0 + k*d (differs from the repeated sum by <1e-12).
static __device__ HL_CODE2 void rs_sample_local(const RsPlan& P, int j, double maxc, double& px, double& py, double& pyaw,
                                int& cs_sign, int& dir) {
    if (j == 0) { px = 0.0; py = 0.0; pyaw = 0.0; cs_sign = 0; dir = P.dir0; return; }
    int si = P.nseg - 1;
    HL_LOOP
    while (si > 0 && j < P.seg[si].first) --si;
    const RsSegPlan& S = P.seg[si];
    int k = j - S.first;
    double off = (k < S.count) ? (S.pd0 + (double)k * S.d) : S.l;
    rs_interp(off, S.letter, maxc, S.ox, S.oy, S.oyaw, px, py, pyaw);
    cs_sign = (S.letter == RS_S) ? 0 : (S.letter == RS_L ? 1 : -1);
    dir = (off > 0.0) ? 1 : -1;
}

// float32 view of a plan for the collision filter: per segment the world position of its origin
// relative to the environment origin and cos/sin of the world heading there.  One call per word.
static __device__ HL_CODE2 void rs_plan_world32(RsPlan& P, const double* q0, double cq, double sq, const double* env_origin) {
    HL_LOOP
    for (int i = 0; i < P.nseg; ++i) {
        RsSegPlan& S = P.seg[i];
        double wx = xadd(xadd(xmul(cq, S.ox), xmul(sq, S.oy)), q0[0]);
        double wy = xadd(xadd(xmul(-sq, S.ox), xmul(cq, S.oy)), q0[1]);
        S.fox = (float)(wx - env_origin[0]);
        S.foy = (float)(wy - env_origin[1]);
        double sn, cs;
        m_sincos(S.oyaw + q0[2], &sn, &cs);
        S.fc0 = (float)cs; S.fs0 = (float)sn;
    }
}

// Pose j in the environment's float32 frame: position and cos/sin of the world yaw (1 sincosf).
static __device__ HL_CODE void rs_sample_world32(const RsPlan& P, int j, float inv_maxc, float& wx, float& wy,
                                                  float& c, float& s) {
    int si = P.nseg - 1;
    HL_LOOP
    while (si > 0 && j < P.seg[si].first) --si;
    const RsSegPlan& S = P.seg[si];
    if (j == 0) { wx = P.seg[0].fox; wy = P.seg[0].foy; c = P.seg[0].fc0; s = P.seg[0].fs0; return; }
    const int k = j - S.first;
    const float off = (float)((k < S.count) ? (S.pd0 + (double)k * S.d) : S.l);
    float dx, dy, cy, sy;
    if (S.letter == RS_S) {
        dx = off * inv_maxc; dy = 0.f; cy = 1.f; sy = 0.f;
    } else {
        float sn, cs;
        sincosf(off, &sn, &cs);
        dx = sn * inv_maxc;
        const float one_m = (1.0f - cs) * inv_maxc;
        if (S.letter == RS_L) { dy = one_m; cy = cs; sy = sn; }
        else { dy = -one_m; cy = cs; sy = -sn; }
    }
    wx = fmaf(S.fc0, dx, fmaf(-S.fs0, dy, S.fox));
    wy = fmaf(S.fs0, dx, fmaf(S.fc0, dy, S.foy));
    c = fmaf(S.fc0, cy, -S.fs0 * sy);
    s = fmaf(S.fs0, cy, S.fc0 * sy);
}

// Local -> world (calc_all_paths, :50-59)
static __device__ HL_CODE void rs_to_world(const double* q0, double cq, double sq, double lx, double ly, double lyaw,
                                            double& wx, double& wy, double& wyaw) {
    // cq = m_cos(-q0yaw), sq = m_sin(-q0yaw)
    wx = xadd(xadd(xmul(cq, lx), xmul(sq, ly)), q0[0]);
    wy = xadd(xadd(xmul(-sq, lx), xmul(cq, ly)), q0[1]);
    wyaw = rs_pi_2_pi(xadd(lyaw, q0[2]));
}

// `while px[-1] == 0.0: pop` of generate_local_course (reeds_shepp.py:520-528) for a word whose end point has a local
// x of exactly 0: the end point goes, then loop samples as long as their x is exactly 0.0.  Rare; one lane.
static __device__ __noinline__ void rs_plan_trailing_pops(RsPlan& P, double maxc) {
    int npts = P.npts - 1;
#pragma unroll 1
    while (npts > 1) {
        const int j = npts - 1;
        int si = P.nseg - 1;
#pragma unroll 1
        while (si > 0 && j < P.seg[si].first) --si;
        const RsSegPlan& S = P.seg[si];
        double pd = S.pd0;
#pragma unroll 1
        for (int k = 0; k < j - S.first; ++k) pd = xadd(pd, S.d);
        double px, py, pyaw;
        rs_interp(pd, S.letter, maxc, S.ox, S.oy, S.oyaw, px, py, pyaw);
        if (px != 0.0) break;
        npts -= 1;
    }
    P.npts = npts;
}

// Plans of up to 6 words at once, one warp: rs_make_plan + rs_plan_world32 re-arranged so that the
// expensive part -- the sine / cosine of every segment's length and start heading, 4 float64 transcendentals per
// segment in rs_interp -- runs on one lane per (word, segment) as TWO sincos calls in lock-step instead of 20 serial
// calls per word.  Lane 5w + i holds segment i of word k0 + w.  Bit-identical to rs_make_plan in everything that feeds
// a decision (sample counts, segment origins and headings: same operations in the same order; sincos == sin, cos and
// sin odd / cos even bit for bit, tools/sincos_check.cu).  The float32 view (fox, foy, fc0, fs0) takes the cosine /
// sine of the world heading from the angle-addition formula in float64 instead of another sincos: it only feeds the
// conservative float32 filter.
static __device__ __noinline__ void rs_make_plans_warp(const int* acc, const double (*lens_all)[HL_RS_MAX_SEGS], int* npts_out, int k0, int nw,
                                                RsPlan* dst, const double* q0, double cq, double sq, const double* origin,
                                                double maxc, double step, int lane) {
    const unsigned FULL_ = 0xffffffffu;
    const int w = lane / 5, i = lane - 5 * w;
    const bool wact = w < nw;
    const int c = wact ? acc[k0 + w] : 0;
    const RsRow row = c_rs_rows[c];
    const int nseg = row.nseg;
    const bool act = wact && i < nseg;
    const double* lens = lens_all[c];
    if (wact && i == 0) {
        // sample bookkeeping of the whole word (generate_local_course's pd / ll / ind chain): sequential, cheap
        RsPlan& P = dst[w];
        P.nseg = nseg;
        P.dir0 = (lens[0] > 0.0) ? 1 : -1;
        double ll = 0.0;
        int ind = 1;
#pragma unroll 1
        for (int j = 0; j < nseg; ++j) {
            const double l = lens[j];
            const double d = (l > 0.0) ? step : -step;
            RsSegPlan& S = P.seg[j];
            S.l = l; S.d = d; S.letter = rs_letter(row.letters, j);
            ind -= 1;
            double pd = (j >= 1 && xmul(lens[j - 1], lens[j]) > 0.0) ? xsub(-d, ll) : xsub(d, ll);
            S.pd0 = pd;
            S.first = ind + 1;
            int cnt = 0;
            const double al = fabs(l);
            if (fabs(pd) <= al) {
                const double a = (d > 0.0) ? pd : -pd;
                const double r = (al - a) / step;
                const double kf = floor(r);
                if (r - kf > 1e-7 && kf + 1.0 - r > 1e-7 && r < 1e7) {
                    cnt = (int)kf + 1;
                    pd = xadd(pd, xmul((double)cnt, d));
                } else {
#pragma unroll 1
                    while (fabs(pd) <= al) { ++cnt; pd = xadd(pd, d); }
                }
            }
            S.count = cnt;
            ind += cnt;
            ll = xsub(xsub(l, pd), d);
            ind += 1;
        }
        P.npts = ind + 1;
    }
    // heading at the start of segment i: oyaw accumulates +-l over the arcs before it, in the reference's order
    double oyaw = 0.0;
#pragma unroll 1
    for (int j = 0; j < HL_RS_MAX_SEGS - 1; ++j) {
        if (act && j < i) {
            const int lt = rs_letter(row.letters, j);
            if (lt == RS_L) oyaw = xadd(oyaw, lens[j]);
            else if (lt == RS_R) oyaw = xsub(oyaw, lens[j]);
        }
    }
    const double l = act ? lens[i] : 0.0;
    const int letter = rs_letter(row.letters, i);
    double sl, cl, so, co;
    m_sincos(l, &sl, &cl);
    m_sincos(oyaw, &so, &co);
    // displacement of segment i in the local frame (rs_interp with cos(-oyaw) = co, sin(-oyaw) = -so)
    double gdx, gdy;
    if (letter == RS_S) {
        const double lm = xdiv(l, maxc);
        gdx = xmul(lm, co); gdy = xmul(lm, so);
    } else {
        const double ldx = xdiv(sl, maxc);
        const double ldy = xdiv(xsub(1.0, cl), letter == RS_L ? maxc : -maxc);
        gdx = xadd(xmul(co, ldx), xmul(-so, ldy));
        gdy = xadd(xmul(so, ldx), xmul(co, ldy));
    }
    // origin of segment i: ((0 + g_0) + g_1) + ... in order
    double ox = 0.0, oy = 0.0;
#pragma unroll
    for (int j = 0; j < HL_RS_MAX_SEGS - 1; ++j) {
        const int src = min(5 * w + j, 31);
        const double gx = __shfl_sync(FULL_, gdx, src), gy = __shfl_sync(FULL_, gdy, src);
        if (j < i) { ox = xadd(ox, gx); oy = xadd(oy, gy); }
    }
    const double ex = xadd(ox, gdx);                             // end point of segment i (x only: the pop test)
    const double end_x = __shfl_sync(FULL_, ex, min(5 * w + nseg - 1, 31));
    if (act) {
        RsSegPlan& S = dst[w].seg[i];
        S.ox = ox; S.oy = oy; S.oyaw = oyaw;
        if (origin) {                                            // float32 view for the collision filter
            const double wx = xadd(xadd(xmul(cq, ox), xmul(sq, oy)), q0[0]);
            const double wy = xadd(xadd(xmul(-sq, ox), xmul(cq, oy)), q0[1]);
            S.fox = (float)(wx - origin[0]);
            S.foy = (float)(wy - origin[1]);
            S.fc0 = (float)(co * cq + so * sq);                  // cos(oyaw + yaw0), sin(oyaw + yaw0): cq = cos yaw0, sq = -sin yaw0
            S.fs0 = (float)(so * cq - co * sq);
        }
    }
    __syncwarp();
    if (wact && i == 0) {
        if (end_x == 0.0) rs_plan_trailing_pops(dst[w], maxc);
        if (npts_out) npts_out[k0 + w] = dst[w].npts;
    }
    __syncwarp();
}

