// hl_common.cuh -- shared host/device declarations of the sm_100a warm-start library.
//
// Data layout in HBM (see DESIGN.md "Data layout"): one EnvDesc per environment
// (offsets into pooled SoA arrays) + pools.  Every geometric quantity exists
// twice: float64 exactly as the host built it (decisions that must be bit-exact)
// and a float32 copy relative to a per-environment origin (bulk filter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/headland_b200.h"

#define HL_MAX_SEGS 16           // lane capsules per environment (device-side arrays)
#define HL_OBS32_STRIDE 28       // floats per obstacle: 4 vertices (8) + 4 x (nx, ny, c) + box form
                                 // [20] is_rect, [21..22] centre, [23..24] unit axis a, [25..26] half extents
#define HL_FIELD32_STRIDE 12     // per field edge: Ax, Ay, Ex, Ey | nx, ny (unit, outward), c = n.A, By | t.A, t.B, pad x2
#define HL_PI 3.141592653589793  // == math.pi

struct EnvDesc {
    int n_obs, obs_off;
    int n_field, field_off;
    int n_seg, seg_off;
    int n_crit, crit_off;
    int n_guide, guide_off;
    int n_aux, aux_off;
    int all_rect, pad0;     // every obstacle quad is a rectangle (box form usable)
    float eps;              // float32 filter band (metres), scaled to the environment extent
    float reach;            // poses farther than this from origin skip the float32 filter
    double origin[2];
    double default_len;
    double body_ext[4];
};

struct EnvBatchDev {
    const EnvDesc* desc;
    const float* obs32;       // [n][HL_OBS32_STRIDE]
    const double* obs64;      // [n][4][2]
    const float* field32;     // [n][HL_FIELD32_STRIDE] relative to origin
    const double* field64;    // [n][2]
    const float* seg32;       // [n][4]  ax, ay, bx, by relative to origin
    const double* seg64;      // [n][4]
    const double* seg_len;    // [n]
    const double* seg_poly;   // [n][66][2]
    const double* crit64;     // [n][2]
    const double* guide_x;    // SoA guide polyline
    const double* guide_y;
    const double* guide_yaw;
    const double* guide_s;
    const double* aux64;      // [n][4]
    int n_env;
};

struct hl_ctx;
struct hl_env_batch {
    EnvBatchDev dev;
    void* allocs[20];
    int n_allocs;
    int device;
    hl_ctx* owner;
    size_t block_bytes;
};

struct hl_ctx {
    int device;
    int sm_count;
    int max_smem_optin;
    void* astar_ws;           // cached workspace of the search kernel
    size_t astar_ws_bytes;
    unsigned int* d_counters; // small device scratch (work-queue counter etc.)
    void* stage;              // pinned staging buffer of hl_env_upload (grow-only)
    size_t stage_bytes;
    void* env_cache[2];       // device blocks of freed environment batches, reused by later uploads (two: double buffering)
    size_t env_cache_bytes[2];
    void* copy_stream;        // non-blocking stream of hl_env_upload's H2D copy (does not wait for running kernels)
    void* mu;                 // std::recursive_mutex*: one upload / cache hand-over at a time (uploads may come from a prefetch thread)
    void* ws_mu;              // std::recursive_mutex*: the search workspace / variant (NOT `mu`: a search must be launchable
                              // while a prefetch thread spends milliseconds inside hl_env_upload)
    void* ls_state;           // level-synchronous search: graph, streams, pools (hl_astar.cu)
    void (*ls_free)(void*);
    int* astar_defer;         // two-phase sweeps: scenarios left over by the first-shot launch (hl_astar.cu)
    size_t astar_defer_cap;
    int astar_variant;        // HL_ASTAR_SPEC / _WARP / _LEVEL (HL_ASTAR_VARIANT read ONCE at hl_ctx_create; hl_ctx_set_astar_variant)
    void* astar_done;         // cudaEvent_t recorded after the last search launch: the workspace / work counter are
                              // per context, so a search on another stream waits for it (searches serialise per ctx)
};
#define HL_ASTAR_SPEC 0
#define HL_ASTAR_WARP 1
#define HL_ASTAR_LEVEL 2

void hl_set_error(const char* fmt, ...);
// Every entry point calls this first: selects the context's device and rejects an environment batch or a
// device pointer that lives on ANOTHER device (or is not device memory at all) with an hl_last_error text
// instead of launching on the wrong GPU.  `probe` may be NULL.
int hl_enter(const hl_ctx* ctx, const hl_env_batch* envs, const void* probe, const char* what);
#define HL_CUDA_OK(call)                                                            \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess) {                                                    \
            hl_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),    \
                         __FILE__, __LINE__);                                       \
            return 1;                                                               \
        }                                                                           \
    } while (0)

// ------------------------------------------------------------------ device math
#ifdef __CUDACC__
// HL_CODE: helpers are force-inlined by default (K1/K2 want that); a translation unit that defines
// HL_SHARED_CODE gets them out-of-line instead (one copy, instruction-cache friendly).
// HL_LOOP: in the search kernels (HL_SHARED_CODE) code bytes are the bottleneck (32 KB instruction cache), so
// runtime loops are kept rolled there; K1/K2 let the compiler unroll them for throughput.
#ifdef HL_SHARED_CODE
#define HL_LOOP _Pragma("unroll 1")
#else
#define HL_LOOP
#endif
#ifndef HL_TINY
#define HL_TINY __forceinline__          // tiny helpers: a call frame costs as much as their body
#endif
#ifdef HL_SHARED_CODE
#define HL_CODE __noinline__
#else
#define HL_CODE __forceinline__
#endif
static __device__ HL_CODE double m_sin(double x) { return sin(x); }
static __device__ HL_CODE double m_cos(double x) { return cos(x); }
static __device__ HL_CODE double m_tan(double x) { return tan(x); }
static __device__ HL_CODE double m_atan2(double y, double x) { return atan2(y, x); }
static __device__ HL_CODE double m_asin(double x) { return asin(x); }
static __device__ HL_CODE double m_acos(double x) { return acos(x); }
static __device__ HL_CODE double m_fmod(double a, double b) { return fmod(a, b); }
static __device__ HL_CODE void m_sincos(double x, double* s, double* c) { sincos(x, s, c); }
// fmod(a, m) for m > 0.  fmod is EXACT, so the two ranges every angle on this path falls into need no libdevice loop:
// |a| < m leaves a as it is, m <= |a| < 2m takes m off once (exact by Sterbenz's lemma); anything else (NaN / inf
// included) goes to fmod.  Same bits as fmod(a, m) for every input (a == -m gives +0 instead of -0; the only caller,
// py_mod_pos, maps both to +0).
static __device__ HL_CODE double m_fmod_pos(double a, double m) {
    const double aa = fabs(a);
    if (aa < m) return a;
    if (aa < 2.0 * m) return (a > 0.0) ? __dadd_rn(a, -m) : __dadd_rn(a, m);
    return fmod(a, m);
}
// float64 arithmetic that must not be contracted into FMAs: the oracle evaluates
// the same expressions with one rounding per operation.
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dadd_rn(a, -b); }
static __device__ HL_TINY double xdiv(double a, double b) { return __ddiv_rn(a, b); }

// Python / numpy floored modulo for a positive modulus (CPython float_rem,
// numpy npy_divmod): fmod, then shift negative remainders up.
static __device__ HL_TINY double py_mod_pos(double a, double m) {
    double r = m_fmod_pos(a, m);     // exact
    if (r != 0.0) {
        if (r < 0.0) r = xadd(r, m);
    } else {
        r = 0.0;                     // copysign(0, m) with m > 0
    }
    return r;
}

// path_utils.angle_wrap: (a + pi) % (2*pi) - pi
__device__ __forceinline__ double angle_wrap(double a) {
    return xsub(py_mod_pos(xadd(a, HL_PI), 2.0 * HL_PI), HL_PI);
}

// reeds_shepp.M: theta % 2pi folded to (-pi, pi]
__device__ __forceinline__ double rs_mod2pi(double theta) {
    double phi = py_mod_pos(theta, 2.0 * HL_PI);
    if (phi < -HL_PI) phi = xadd(phi, 2.0 * HL_PI);
    if (phi > HL_PI) phi = xsub(phi, 2.0 * HL_PI);
    return phi;
}

// reeds_shepp.pi_2_pi
__device__ __forceinline__ double rs_pi_2_pi(double t) {
    while (t > HL_PI) t = xsub(t, 2.0 * HL_PI);
    while (t < -HL_PI) t = xadd(t, 2.0 * HL_PI);
    return t;
}

// Correctly rounded hypot for the magnitudes on this path (no over/underflow
// handling needed: |x|,|y| are metres or unit-circle offsets).  glibc >= 2.35's
// hypot is correctly rounded, CUDA's is 1-2 ulp; this one follows Borges'
// "fused" algorithm: h = sqrt(x^2+y^2) with an FMA-computed residual correction.
static __device__ HL_CODE double hypot_cr(double x, double y) {
    double ax = fabs(x), ay = fabs(y);
    if (ax < ay) { double t = ax; ax = ay; ay = t; }
    if (ay == 0.0) return ax;
    double h = sqrt(fma(ax, ax, ay * ay));
    double h_sq = h * h;
    double ax_sq = ax * ax;
    double corr = fma(-ay, ay, h_sq - ax_sq) + fma(h, h, -h_sq) - fma(ax, ax, -ax_sq);
    return h - corr / (2.0 * h);
}
#endif
