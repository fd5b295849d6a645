// hl_ctx.cu -- context, error reporting, environment upload (host -> HBM).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cstdlib>
#include <thread>
#include <mutex>
#include <cmath>
#include "hl_common.cuh"

static thread_local char g_err[512] = "";

void hl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* hl_last_error(void) { return g_err; }
extern "C" int hl_abi_version(void) { return HL_ABI_VERSION; }

extern "C" int hl_ctx_create(hl_ctx** out, int device) {
    if (!out) { hl_set_error("hl_ctx_create: out is NULL"); return 1; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        hl_set_error("hl_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return 1;
    }
    if (device < 0 || device >= n) { hl_set_error("hl_ctx_create: bad device %d of %d", device, n); return 1; }
    HL_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    HL_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {                 // only sm_100a SASS is in the library (no PTX): sm_12x has no kernel image either
        hl_set_error("hl_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                     device, prop.major, prop.minor);
        return 1;
    }
    hl_ctx* c = new hl_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    c->astar_ws = nullptr;
    c->astar_ws_bytes = 0;
    c->astar_defer = nullptr;
    c->astar_defer_cap = 0;
    c->d_counters = nullptr;
    c->stage = nullptr;
    c->stage_bytes = 0;
    c->env_cache[0] = c->env_cache[1] = nullptr;
    c->env_cache_bytes[0] = c->env_cache_bytes[1] = 0;
    c->ls_state = nullptr;
    c->ls_free = nullptr;
    c->copy_stream = nullptr;
    c->mu = new std::recursive_mutex();
    c->ws_mu = new std::recursive_mutex();
    {
        cudaStream_t cs = nullptr;
        if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) == cudaSuccess) c->copy_stream = cs;
    }
    c->astar_variant = HL_ASTAR_SPEC;
    if (const char* var = getenv("HL_ASTAR_VARIANT")) {       // read once; hl_ctx_set_astar_variant changes it later
        if (strcmp(var, "warp") == 0) c->astar_variant = HL_ASTAR_WARP;
        else if (strcmp(var, "level") == 0) c->astar_variant = HL_ASTAR_LEVEL;
    }
    c->astar_done = nullptr;
    {
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) c->astar_done = ev;
    }
    if (!c->astar_done || cudaMalloc(&c->d_counters, 256 * sizeof(unsigned int)) != cudaSuccess) {
        hl_set_error("hl_ctx_create: cudaMalloc / cudaEventCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (c->astar_done) cudaEventDestroy((cudaEvent_t)c->astar_done);
        if (c->copy_stream) cudaStreamDestroy((cudaStream_t)c->copy_stream);
        delete (std::recursive_mutex*)c->mu;
        delete (std::recursive_mutex*)c->ws_mu;
        delete c;
        return 1;
    }
    cudaMemset(c->d_counters, 0, 256 * sizeof(unsigned int));
    *out = c;
    return 0;
}

extern "C" int hl_ctx_device(const hl_ctx* ctx) { return ctx ? ctx->device : -1; }
extern "C" int hl_env_device(const hl_env_batch* envs) { return envs ? envs->device : -1; }

extern "C" int hl_ctx_set_astar_variant(hl_ctx* ctx, int variant) {
    if (!ctx || variant < HL_ASTAR_SPEC || variant > HL_ASTAR_LEVEL) { hl_set_error("hl_ctx_set_astar_variant: bad arguments"); return 1; }
    std::lock_guard<std::recursive_mutex> lock(*(std::recursive_mutex*)ctx->ws_mu);
    ctx->astar_variant = variant;
    return 0;
}

int hl_enter(const hl_ctx* ctx, const hl_env_batch* envs, const void* probe, const char* what) {
    if (!ctx) { hl_set_error("%s: ctx is NULL", what); return 1; }
    if (envs && envs->device != ctx->device) {
        hl_set_error("%s: the environment batch lives on device %d but the context is for device %d "
                     "(upload it through the context of the device that runs the kernels)", what, envs->device, ctx->device);
        return 1;
    }
    if (envs && envs->owner != ctx) { hl_set_error("%s: the environment batch belongs to another context", what); return 1; }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) { hl_set_error("%s: cudaSetDevice(%d) failed: %s", what, ctx->device, cudaGetErrorString(e)); return 1; }
    if (probe) {
        cudaPointerAttributes a;
        e = cudaPointerGetAttributes(&a, probe);
        if (e != cudaSuccess) { cudaGetLastError(); hl_set_error("%s: cudaPointerGetAttributes failed: %s", what, cudaGetErrorString(e)); return 1; }
        if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) {
            hl_set_error("%s: a d_* argument is not device memory (host pointer passed?)", what); return 1;
        }
        if (a.type == cudaMemoryTypeDevice && a.device != ctx->device) {
            hl_set_error("%s: a d_* argument lives on device %d but the context is for device %d", what, a.device, ctx->device);
            return 1;
        }
    }
    return 0;
}

extern "C" void hl_ctx_destroy(hl_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->astar_ws) cudaFree(ctx->astar_ws);
    if (ctx->astar_defer) cudaFree(ctx->astar_defer);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    for (int k = 0; k < 2; ++k) if (ctx->env_cache[k]) cudaFree(ctx->env_cache[k]);
    if (ctx->ls_state && ctx->ls_free) ctx->ls_free(ctx->ls_state);
    if (ctx->copy_stream) cudaStreamDestroy((cudaStream_t)ctx->copy_stream);
    if (ctx->astar_done) cudaEventDestroy((cudaEvent_t)ctx->astar_done);
    delete (std::recursive_mutex*)ctx->mu;
    delete (std::recursive_mutex*)ctx->ws_mu;
    delete ctx;
}

extern "C" int hl_ctx_sm_count(const hl_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" void hl_env_free(hl_env_batch* envs) {
    if (!envs) return;
    cudaSetDevice(envs->device);
    hl_ctx* c = envs->owner;
    std::unique_lock<std::recursive_mutex> lock;
    if (c && c->mu) lock = std::unique_lock<std::recursive_mutex>(*(std::recursive_mutex*)c->mu);
    for (int i = 0; i < envs->n_allocs; ++i) {
        // keep up to two blocks per context for later uploads (cudaFree synchronises the device): the smallest
        // cached block gives way to a bigger one
        int slot = -1;
        if (i == 0 && c) {
            if (!c->env_cache[0]) slot = 0;
            else if (!c->env_cache[1]) slot = 1;
            else {
                const int small = c->env_cache_bytes[0] <= c->env_cache_bytes[1] ? 0 : 1;
                if (envs->block_bytes > c->env_cache_bytes[small]) slot = small;
            }
        }
        if (slot >= 0) {
            cudaDeviceSynchronize();               // like cudaFree: no kernel on any stream may still read the block
            if (c->env_cache[slot]) cudaFree(c->env_cache[slot]);
            c->env_cache[slot] = envs->allocs[0]; c->env_cache_bytes[slot] = envs->block_bytes;
        } else cudaFree(envs->allocs[i]);
    }
    delete envs;
}

extern "C" int32_t hl_env_count(const hl_env_batch* envs) { return envs ? envs->dev.n_env : 0; }

// One device block + one pinned staging block per upload.  The host only lays the caller's float64 arrays end to
// end in the (context-cached, grow-only) pinned buffer -- memcpy, no arithmetic -- and ONE cudaMemcpyAsync moves them;
// everything DERIVED (float32 frame origin, filter band, float32 obstacle / field-edge / lane records, the SoA
// transposition of the guide polyline) is computed on the device by k_env_derive, one CTA per environment.
// History: 15 std::vectors + 15 pageable copies 37 ms per 4096 environments; host-side derivation into one pinned
// block 3.5 ms on 8 threads (but 8 ranks share the host cores of a box); device-side derivation: see DESIGN.md.
enum { P_DESC = 0, P_OBS64, P_FIELD64, P_SEG64, P_SEGLEN, P_SEGPOLY, P_CRIT64, P_AUX64, P_GUIDE_AOS,   // host-filled
       P_OBS32, P_FIELD32, P_SEG32, P_GX, P_GY, P_GYAW, P_GS, P_COUNT };                               // device-derived
#define P_HOST_POOLS P_OBS32
struct PoolLayout {
    size_t off[P_COUNT];
    size_t host_bytes;      // the first P_HOST_POOLS pools, copied from the staging block
    size_t total;
};
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct EnvDerive {
    EnvDesc* desc;
    const double* obs64; const double* field64; const double* seg64; const double* aux64; const double* guide_aos;
    float* obs32; float* field32; float* seg32;
    double* gx; double* gy; double* gyaw; double* gs;
};

// float64 in the host's operation order (one rounding per operation, no FMA contraction): the records are bit for
// bit the ones the round-1 host code produced.
__global__ void __launch_bounds__(128) k_env_derive(EnvDerive A, int n_env) {
    const int e = blockIdx.x;
    if (e >= n_env) return;
    EnvDesc& D = A.desc[e];
    const int t = threadIdx.x, nt = blockDim.x;
    __shared__ double s_lo[2][128], s_hi[2][128];
    __shared__ double s_origin[2], s_orient;
    __shared__ int s_all_rect;
    const double* obs = A.obs64 + 8 * (size_t)D.obs_off;
    const double* fld = A.field64 + 2 * (size_t)D.field_off;
    const double* seg = A.seg64 + 4 * (size_t)D.seg_off;
    double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
    for (int i = t; i < 4 * D.n_obs; i += nt)
        for (int c = 0; c < 2; ++c) { lo[c] = fmin(lo[c], obs[2 * i + c]); hi[c] = fmax(hi[c], obs[2 * i + c]); }
    for (int i = t; i < D.n_field; i += nt)
        for (int c = 0; c < 2; ++c) { lo[c] = fmin(lo[c], fld[2 * i + c]); hi[c] = fmax(hi[c], fld[2 * i + c]); }
    for (int i = t; i < 2 * D.n_seg; i += nt)
        for (int c = 0; c < 2; ++c) { lo[c] = fmin(lo[c], seg[2 * i + c]); hi[c] = fmax(hi[c], seg[2 * i + c]); }
    for (int c = 0; c < 2; ++c) { s_lo[c][t] = lo[c]; s_hi[c][t] = hi[c]; }
    if (t == 0) s_all_rect = 1;
    __syncthreads();
    for (int w = nt / 2; w > 0; w >>= 1) {
        if (t < w)
            for (int c = 0; c < 2; ++c) { s_lo[c][t] = fmin(s_lo[c][t], s_lo[c][t + w]); s_hi[c][t] = fmax(s_hi[c][t], s_hi[c][t + w]); }
        __syncthreads();
    }
    if (t == 0) {
        double l0 = s_lo[0][0], l1 = s_lo[1][0], h0 = s_hi[0][0], h1 = s_hi[1][0];
        if (l0 > h0) { l0 = l1 = h0 = h1 = 0.0; }
        const double ox = xmul(0.5, xadd(l0, h0)), oy = xmul(0.5, xadd(l1, h1));
        s_origin[0] = ox; s_origin[1] = oy;
        D.origin[0] = ox; D.origin[1] = oy;
        // origin of the float32 frame = centre of the bounding box of all geometry; band scaled to the extent
        const double extent = xmul(xmul(0.5, fmax(xsub(h0, l0), xsub(h1, l1))), 1.4142135623730951);
        double foot = hypot_cr(fmax(fabs(D.body_ext[0]), fabs(D.body_ext[1])), fmax(fabs(D.body_ext[2]), fabs(D.body_ext[3])));
        for (int a = 0; a < D.n_aux; ++a) {
            const double* x = A.aux64 + 4 * (size_t)(D.aux_off + a);
            foot = fmax(foot, hypot_cr(fmax(fabs(x[0]), fabs(x[1])), fmax(fabs(x[2]), fabs(x[3]))));
        }
        const double reach = xadd(xadd(xadd(extent, 6.0), foot), 4.0);     // lane radius 6 m (reference_line_heuristic.py:66) + margin
        D.reach = (float)reach;
        D.eps = (float)xmul(xmul(32.0, 1.1920929e-07), fmax(reach, 8.0));
        double area2 = 0.0;
        for (int i = 0; i < D.n_field; ++i) {
            const int j = (i + 1 == D.n_field) ? 0 : i + 1;
            area2 = xadd(area2, xsub(xmul(fld[2 * i], fld[2 * j + 1]), xmul(fld[2 * j], fld[2 * i + 1])));
        }
        s_orient = area2 < 0 ? -1.0 : 1.0;
    }
    __syncthreads();
    const double ox = s_origin[0], oy = s_origin[1];
    for (int k = t; k < D.n_obs; k += nt) {
        const double* V = obs + 8 * k;
        float* o = A.obs32 + HL_OBS32_STRIDE * (size_t)(D.obs_off + k);
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = (float)xsub(V[2 * i], ox);
            o[2 * i + 1] = (float)xsub(V[2 * i + 1], oy);
        }
        for (int i = 0; i < 4; ++i) {
            const int j = (i + 1) & 3;
            const double ex = xsub(V[2 * j], V[2 * i]), ey = xsub(V[2 * j + 1], V[2 * i + 1]);
            const double ln = sqrt(xadd(xmul(ex, ex), xmul(ey, ey)));
            const double nx = ln > 0 ? xdiv(ey, ln) : 0.0, ny = ln > 0 ? xdiv(-ex, ln) : 0.0;
            const double cc = xadd(xmul(nx, xsub(V[2 * i], ox)), xmul(ny, xsub(V[2 * i + 1], oy)));
            o[8 + 3 * i] = (float)nx; o[9 + 3 * i] = (float)ny; o[10 + 3 * i] = (float)cc;
        }
        // box form when the quad is a rectangle (tree rows, obstacle squares): centre, unit axis, half extents
        const double e0x = xsub(V[2], V[0]), e0y = xsub(V[3], V[1]), e1x = xsub(V[4], V[2]), e1y = xsub(V[5], V[3]);
        const double e2x = xsub(V[6], V[4]), e2y = xsub(V[7], V[5]), e3x = xsub(V[0], V[6]), e3y = xsub(V[1], V[7]);
        const double l0 = sqrt(xadd(xmul(e0x, e0x), xmul(e0y, e0y))), l1 = sqrt(xadd(xmul(e1x, e1x), xmul(e1y, e1y)));
        const double scale = fmax(l0, l1);
        const bool para = xadd(xadd(xadd(fabs(xadd(e0x, e2x)), fabs(xadd(e0y, e2y))), fabs(xadd(e1x, e3x))), fabs(xadd(e1y, e3y))) <= xmul(1e-9, scale);
        const bool perp = l0 > 0 && l1 > 0 && fabs(xadd(xmul(e0x, e1x), xmul(e0y, e1y))) <= xmul(xmul(1e-9, l0), l1);
        const bool is_rect = para && perp;
        if (!is_rect) s_all_rect = 0;
        o[20] = is_rect ? 1.0f : 0.0f;
        o[21] = (float)xsub(xmul(0.25, xadd(xadd(xadd(V[0], V[2]), V[4]), V[6])), ox);
        o[22] = (float)xsub(xmul(0.25, xadd(xadd(xadd(V[1], V[3]), V[5]), V[7])), oy);
        o[23] = (float)(is_rect ? xdiv(e0x, l0) : 1.0); o[24] = (float)(is_rect ? xdiv(e0y, l0) : 0.0);
        o[25] = (float)xmul(0.5, l0); o[26] = (float)xmul(0.5, l1); o[27] = 0.0f;
    }
    {
        const double orient = s_orient;
        for (int i = t; i < D.n_field; i += nt) {
            const int j = (i + 1 == D.n_field) ? 0 : i + 1;
            const double Ax = xsub(fld[2 * i], ox), Ay = xsub(fld[2 * i + 1], oy);
            const double Bx = xsub(fld[2 * j], ox), By = xsub(fld[2 * j + 1], oy);
            const double ex = xsub(Bx, Ax), ey = xsub(By, Ay), ln = sqrt(xadd(xmul(ex, ex), xmul(ey, ey)));
            const double nx = ln > 0 ? xdiv(xmul(orient, ey), ln) : 0.0, ny = ln > 0 ? xdiv(xmul(-orient, ex), ln) : 0.0;
            float* f = A.field32 + HL_FIELD32_STRIDE * (size_t)(D.field_off + i);
            f[0] = (float)Ax; f[1] = (float)Ay; f[2] = (float)xsub(Bx, Ax); f[3] = (float)xsub(By, Ay);
            f[4] = (float)nx; f[5] = (float)ny; f[6] = (float)xadd(xmul(nx, Ax), xmul(ny, Ay));
            f[7] = (float)By;                                   // == the next record's (float)Ay, bit for bit
            f[8] = (float)xadd(xmul(-ny, Ax), xmul(nx, Ay)); f[9] = (float)xadd(xmul(-ny, Bx), xmul(nx, By));
            f[10] = 0.0f; f[11] = 0.0f;
        }
    }
    for (int i = t; i < 4 * D.n_seg; i += nt)
        A.seg32[4 * (size_t)D.seg_off + i] = (float)xsub(seg[i], (i & 1) ? oy : ox);
    {
        const double* g = A.guide_aos + 4 * (size_t)D.guide_off;
        const size_t go = (size_t)D.guide_off;
        for (int i = t; i < D.n_guide; i += nt) {
            A.gx[go + i] = g[4 * i]; A.gy[go + i] = g[4 * i + 1]; A.gyaw[go + i] = g[4 * i + 2]; A.gs[go + i] = g[4 * i + 3];
        }
    }
    __syncthreads();
    if (t == 0) { D.all_rect = s_all_rect; D.pad0 = 0; }
}

extern "C" int hl_env_upload(hl_ctx* ctx, const HlEnvHost* h, int32_t n_env, hl_env_batch** out) {
    if (!ctx || !h || !out || n_env <= 0) { hl_set_error("hl_env_upload: bad arguments"); return 1; }
    if (hl_enter(ctx, nullptr, nullptr, "hl_env_upload")) return 1;
    std::lock_guard<std::recursive_mutex> lock(*(std::recursive_mutex*)ctx->mu);      // the staging buffer and the block cache are shared
    size_t n_obs = 0, n_field = 0, n_seg = 0, n_crit = 0, n_guide = 0, n_aux = 0;
    std::vector<size_t> offs(6 * (size_t)n_env);
    for (int e = 0; e < n_env; ++e) {
        const HlEnvHost& E = h[e];
        size_t* o6 = &offs[6 * (size_t)e];
        o6[0] = n_obs; o6[1] = n_field; o6[2] = n_seg; o6[3] = n_crit; o6[4] = n_guide; o6[5] = n_aux;
        if (E.n_seg > HL_MAX_SEGS) { hl_set_error("hl_env_upload: env %d has %d lane segments (max %d)", e, E.n_seg, HL_MAX_SEGS); return 1; }
        if (E.n_obs < 0 || E.n_field < 0 || E.n_seg < 0 || E.n_guide < 0 || E.n_aux < 0 || E.n_crit < 0) {
            hl_set_error("hl_env_upload: env %d has a negative count", e); return 1;
        }
        n_obs += E.n_obs; n_field += E.n_field; n_seg += E.n_seg; n_crit += E.n_crit; n_guide += E.n_guide; n_aux += E.n_aux;
    }
    size_t bytes[P_COUNT];
    bytes[P_DESC] = sizeof(EnvDesc) * (size_t)n_env;
    bytes[P_OBS64] = sizeof(double) * 8 * n_obs;       bytes[P_FIELD64] = sizeof(double) * 2 * n_field;
    bytes[P_SEG64] = sizeof(double) * 4 * n_seg;       bytes[P_SEGLEN] = sizeof(double) * n_seg;
    bytes[P_SEGPOLY] = sizeof(double) * 2 * HL_CAPSULE_VERTS * n_seg;
    bytes[P_CRIT64] = sizeof(double) * 2 * n_crit;     bytes[P_AUX64] = sizeof(double) * 4 * n_aux;
    bytes[P_GUIDE_AOS] = sizeof(double) * 4 * n_guide;
    bytes[P_OBS32] = sizeof(float) * HL_OBS32_STRIDE * n_obs;
    bytes[P_FIELD32] = sizeof(float) * HL_FIELD32_STRIDE * n_field;
    bytes[P_SEG32] = sizeof(float) * 4 * n_seg;
    bytes[P_GX] = bytes[P_GY] = bytes[P_GYAW] = bytes[P_GS] = sizeof(double) * n_guide;
    PoolLayout L;
    L.total = 0;
    for (int k = 0; k < P_COUNT; ++k) {
        if (k == P_HOST_POOLS) L.host_bytes = L.total;
        L.off[k] = L.total; L.total += al256(bytes[k] ? bytes[k] : 1);
    }
    if (L.host_bytes > ctx->stage_bytes) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr; ctx->stage_bytes = 0;
        HL_CUDA_OK(cudaMallocHost(&ctx->stage, L.host_bytes));
        ctx->stage_bytes = L.host_bytes;
    }
    char* S = (char*)ctx->stage;
    EnvDesc* desc = (EnvDesc*)(S + L.off[P_DESC]);
    double* obs64 = (double*)(S + L.off[P_OBS64]);     double* field64 = (double*)(S + L.off[P_FIELD64]);
    double* seg64 = (double*)(S + L.off[P_SEG64]);     double* seg_len = (double*)(S + L.off[P_SEGLEN]);
    double* seg_poly = (double*)(S + L.off[P_SEGPOLY]); double* crit64 = (double*)(S + L.off[P_CRIT64]);
    double* aux64 = (double*)(S + L.off[P_AUX64]);     double* guide = (double*)(S + L.off[P_GUIDE_AOS]);
    auto fill = [&](int e_lo, int e_hi) {
    for (int e = e_lo; e < e_hi; ++e) {
        const HlEnvHost& E = h[e];
        EnvDesc& D = desc[e];
        const size_t* o6 = &offs[6 * (size_t)e];
        const size_t c_obs = o6[0], c_field = o6[1], c_seg = o6[2], c_crit = o6[3], c_guide = o6[4], c_aux = o6[5];
        D.n_obs = E.n_obs;     D.obs_off = (int)c_obs;
        D.n_field = E.n_field; D.field_off = (int)c_field;
        D.n_seg = E.n_seg;     D.seg_off = (int)c_seg;
        D.n_crit = E.n_crit;   D.crit_off = (int)c_crit;
        D.n_guide = E.n_guide; D.guide_off = (int)c_guide;
        D.n_aux = E.n_aux;     D.aux_off = (int)c_aux;
        D.default_len = E.default_search_length;
        for (int k = 0; k < 4; ++k) D.body_ext[k] = E.body_ext[k];
        D.origin[0] = D.origin[1] = 0.0; D.reach = 0.f; D.eps = 0.f; D.all_rect = 1; D.pad0 = 0;   // k_env_derive fills these
        if (E.n_obs) memcpy(obs64 + 8 * c_obs, E.obs_xy, sizeof(double) * 8 * E.n_obs);
        if (E.n_field) memcpy(field64 + 2 * c_field, E.field_xy, sizeof(double) * 2 * E.n_field);
        if (E.n_seg) {
            memcpy(seg64 + 4 * c_seg, E.seg_xy, sizeof(double) * 4 * E.n_seg);
            memcpy(seg_len + c_seg, E.seg_len, sizeof(double) * E.n_seg);
            memcpy(seg_poly + 2 * HL_CAPSULE_VERTS * c_seg, E.seg_poly, sizeof(double) * 2 * HL_CAPSULE_VERTS * E.n_seg);
        }
        if (E.n_crit) memcpy(crit64 + 2 * c_crit, E.crit_xy, sizeof(double) * 2 * E.n_crit);
        if (E.n_guide) memcpy(guide + 4 * c_guide, E.guide, sizeof(double) * 4 * E.n_guide);
        if (E.n_aux) memcpy(aux64 + 4 * c_aux, E.aux_ext, sizeof(double) * 4 * E.n_aux);
    }
    };
    {
        int nt = (int)std::thread::hardware_concurrency();
        // one process per GPU: share the host cores between the ranks of this node (torchrun exports LOCAL_WORLD_SIZE)
        const char* lws = getenv("LOCAL_WORLD_SIZE");
        const int ranks = lws ? atoi(lws) : 1;
        if (ranks > 1) nt /= ranks;
        nt = nt < 1 ? 1 : (nt > 4 ? 4 : nt);
        if (n_env < 1024) nt = 1;
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(fill, (int)((long long)n_env * t / nt), (int)((long long)n_env * (t + 1) / nt));
        fill(0, (int)((long long)n_env / nt));
        for (auto& x : th) x.join();
    }
    hl_env_batch* b = new hl_env_batch();
    b->n_allocs = 0;
    b->device = ctx->device;
    b->dev.n_env = n_env;
    char* d = nullptr;
    b->owner = ctx;
    b->block_bytes = L.total;
    int pick = -1;                                                    // smallest cached block that is big enough
    for (int k = 0; k < 2; ++k)
        if (ctx->env_cache[k] && ctx->env_cache_bytes[k] >= L.total &&
            (pick < 0 || ctx->env_cache_bytes[k] < ctx->env_cache_bytes[pick])) pick = k;
    if (pick >= 0) {                                                  // reuse the block of a batch freed earlier
        d = (char*)ctx->env_cache[pick]; b->block_bytes = ctx->env_cache_bytes[pick];
        ctx->env_cache[pick] = nullptr; ctx->env_cache_bytes[pick] = 0;
    } else if (cudaMalloc(&d, L.total) != cudaSuccess) { delete b; hl_set_error("hl_env_upload: cudaMalloc(%zu) failed", L.total); return 1; }
    b->allocs[b->n_allocs++] = d;
    cudaStream_t cs = (cudaStream_t)ctx->copy_stream;       // nullptr = legacy default stream
    EnvDerive A;
    A.desc = (EnvDesc*)(d + L.off[P_DESC]);
    A.obs64 = (const double*)(d + L.off[P_OBS64]);     A.field64 = (const double*)(d + L.off[P_FIELD64]);
    A.seg64 = (const double*)(d + L.off[P_SEG64]);     A.aux64 = (const double*)(d + L.off[P_AUX64]);
    A.guide_aos = (const double*)(d + L.off[P_GUIDE_AOS]);
    A.obs32 = (float*)(d + L.off[P_OBS32]); A.field32 = (float*)(d + L.off[P_FIELD32]); A.seg32 = (float*)(d + L.off[P_SEG32]);
    A.gx = (double*)(d + L.off[P_GX]); A.gy = (double*)(d + L.off[P_GY]); A.gyaw = (double*)(d + L.off[P_GYAW]); A.gs = (double*)(d + L.off[P_GS]);
    cudaError_t ce = cudaMemcpyAsync(d, S, L.host_bytes, cudaMemcpyHostToDevice, cs);
    if (ce == cudaSuccess) {
        k_env_derive<<<n_env, 128, 0, cs>>>(A, n_env);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(cs);
    if (ce != cudaSuccess) {
        hl_env_free(b); hl_set_error("hl_env_upload: copy / k_env_derive failed: %s", cudaGetErrorString(ce)); return 1;
    }
    b->dev.desc = A.desc;
    b->dev.obs32 = A.obs32;     b->dev.obs64 = A.obs64;
    b->dev.field32 = A.field32; b->dev.field64 = A.field64;
    b->dev.seg32 = A.seg32;     b->dev.seg64 = A.seg64;
    b->dev.seg_len = (const double*)(d + L.off[P_SEGLEN]); b->dev.seg_poly = (const double*)(d + L.off[P_SEGPOLY]);
    b->dev.crit64 = (const double*)(d + L.off[P_CRIT64]);
    b->dev.guide_x = A.gx; b->dev.guide_y = A.gy; b->dev.guide_yaw = A.gyaw; b->dev.guide_s = A.gs;
    b->dev.aux64 = A.aux64;
    *out = b;
    return 0;
}
