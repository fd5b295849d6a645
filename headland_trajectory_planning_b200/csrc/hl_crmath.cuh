// hl_crmath.cuh -- sin / cos that round like a correctly rounded libm.
//
// The reference's sin / cos are glibc's / numpy's (< 0.55 ulp, i.e. the correctly rounded value in all but a
// vanishing fraction of calls) while CUDA's are "<= 2 ulp": they differ in the last bit in a few percent of calls.
// cr_sincos evaluates sin and cos in double-double arithmetic (relative error < 2^-68) and rounds once, so it
// returns the correctly rounded double unless the exact value lies within 2^-68 of a rounding boundary.  It costs
// ~4x CUDA's sincos, so it is used where the rollout is cheap and bit-equality with numpy is worth it (the
// candidate-path generators of hl_ypark.cu); the search kernels keep CUDA's functions (measured: +70 % K4 time
// with cr_sincos everywhere) -- see DESIGN.md section 3 for what a last-bit difference can and cannot change.
//
// Valid for |x| <= 2^20 (arguments here are headings and arc lengths of a few radians); larger / non-finite
// arguments fall back to the CUDA / libm functions.  Compiles for host and device (tests/test_crmath.py checks the
// host build against a 60-digit reference).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define CR_HD __host__ __device__ __forceinline__
#else
#define CR_HD static inline
#endif

struct crdd { double hi, lo; };

CR_HD crdd cr_two_sum(double a, double b) {                 // exact a + b
    double s = a + b;
    double bb = s - a;
    crdd r; r.hi = s; r.lo = (a - (s - bb)) + (b - bb);
    return r;
}
CR_HD crdd cr_fast_two_sum(double a, double b) {            // |a| >= |b|
    double s = a + b;
    crdd r; r.hi = s; r.lo = b - (s - a);
    return r;
}
CR_HD crdd cr_two_prod(double a, double b) {                // exact a * b
    double p = a * b;
    crdd r; r.hi = p; r.lo = fma(a, b, -p);
    return r;
}
CR_HD crdd cr_add(crdd a, crdd b) {
    crdd s = cr_two_sum(a.hi, b.hi);
    crdd t = cr_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = cr_fast_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return cr_fast_two_sum(s.hi, s.lo);
}
CR_HD crdd cr_add_d(crdd a, double b) {
    crdd s = cr_two_sum(a.hi, b);
    s.lo += a.lo;
    return cr_fast_two_sum(s.hi, s.lo);
}
CR_HD crdd cr_mul(crdd a, crdd b) {
    crdd p = cr_two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return cr_fast_two_sum(p.hi, p.lo);
}
CR_HD crdd cr_mul_d(crdd a, double b) {
    crdd p = cr_two_prod(a.hi, b);
    p.lo += a.lo * b;
    return cr_fast_two_sum(p.hi, p.lo);
}

// sin(r), cos(r) for a double-double |r| <= pi/4 + eps: the terms that are not below 2^-70 relative are summed in
// double-double (Horner on r^2), the tail of the series in plain double.
CR_HD void cr_kernel(crdd r, crdd* s, crdd* c) {
    const crdd r2 = cr_mul(r, r);
    const double z = r2.hi;
    // tails in double:  sin: r^13/13! ...   cos: r^14/14! ...   (relative size < 2e-13 and < 1e-14 of the result)
    const double st = z * (1.0 / 6227020800.0 + z * (-1.0 / 1307674368000.0 + z * (1.0 / 355687428096000.0)));
    const double ct = z * (-1.0 / 87178291200.0 + z * (1.0 / 20922789888000.0 + z * (-1.0 / 6402373705728000.0)));
    // double-double coefficients 1/k! (hi, lo)
    const crdd S3 = {-0.16666666666666666, -9.25185853854297e-18};     // -1/3!
    const crdd S5 = {0.008333333333333333, 1.1564823173178714e-19};       //  1/5!
    const crdd S7 = {-0.0001984126984126984, -1.7209558293420705e-22};     // -1/7!
    const crdd S9 = {2.7557319223985893e-06, -1.858393274046472e-22};      //  1/9!
    const crdd S11 = {-2.505210838544172e-08, 1.448814070935912e-24};     // -1/11!
    const crdd C2 = {-0.5, 0.0};
    const crdd C4 = {0.041666666666666664, 2.3129646346357427e-18};       //  1/4!
    const crdd C6 = {-0.001388888888888889, 5.300543954373577e-20};      // -1/6!
    const crdd C8 = {2.48015873015873e-05, 2.1511947866775882e-23};       //  1/8!
    const crdd C10 = {-2.755731922398589e-07, -2.3767714622250297e-23};    // -1/10!
    const crdd C12 = {2.08767569878681e-09, -1.20734505911326e-25};     //  1/12!
    // sin = r * (1 + z*(S3 + z*(S5 + z*(S7 + z*(S9 + z*(S11 + st))))))
    crdd p = cr_add_d(S11, st);
    p = cr_add(cr_mul(p, r2), S9);
    p = cr_add(cr_mul(p, r2), S7);
    p = cr_add(cr_mul(p, r2), S5);
    p = cr_add(cr_mul(p, r2), S3);
    p = cr_mul(p, r2);
    p = cr_mul(p, r);
    *s = cr_add(r, p);
    // cos = 1 + z*(C2 + z*(C4 + z*(C6 + z*(C8 + z*(C10 + z*(C12 + ct))))))
    crdd q = cr_add_d(C12, ct);
    q = cr_add(cr_mul(q, r2), C10);
    q = cr_add(cr_mul(q, r2), C8);
    q = cr_add(cr_mul(q, r2), C6);
    q = cr_add(cr_mul(q, r2), C4);
    q = cr_add(cr_mul(q, r2), C2);
    q = cr_mul(q, r2);
    *c = cr_add_d(q, 1.0);
}

// x = k * pi/2 + r, |r| <= pi/4 (Cody-Waite with 33-bit pieces: every product k * piece is exact for |k| < 2^20)
CR_HD crdd cr_reduce(double x, int* quadrant) {
    const double INV_PIO2 = 0.6366197723675814;
    const double P1 = 1.5707963267341256;      // first 33 bits of pi/2
    const double P2 = 6.077100506303966e-11;      // next 33 bits
    const double P3 = 2.0222662487111665e-21;      // next 33 bits
    const double P4 = 8.4784276603689e-32;      // pi/2 - (P1 + P2 + P3), to double precision
    const double k = rint(x * INV_PIO2);
    *quadrant = (int)((long long)k & 3);
    crdd r = cr_two_sum(x, -k * P1);                    // k*P1 exact
    r = cr_add_d(r, -k * P2);
    r = cr_add_d(r, -k * P3);
    r = cr_add_d(r, -k * P4);
    return r;
}

CR_HD void cr_sincos(double x, double* sn, double* cs) {
    if (!(fabs(x) <= 1048576.0)) { *sn = sin(x); *cs = cos(x); return; }      // out of the supported range / NaN
    if (x == 0.0) { *sn = x; *cs = 1.0; return; }
    int q;
    const crdd r = cr_reduce(x, &q);
    crdd s, c;
    cr_kernel(r, &s, &c);
    double sv, cv;
    switch (q) {
    case 0: sv = s.hi; cv = c.hi; break;
    case 1: sv = c.hi; cv = -s.hi; break;
    case 2: sv = -s.hi; cv = -c.hi; break;
    default: sv = -c.hi; cv = s.hi; break;
    }
    *sn = sv; *cs = cv;
}
CR_HD double cr_sin(double x) { double s, c; cr_sincos(x, &s, &c); return s; }
CR_HD double cr_cos(double x) { double s, c; cr_sincos(x, &s, &c); return c; }
