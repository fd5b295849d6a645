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
    if (prop.major < 10) {
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
    c->d_counters = nullptr;
    c->stage = nullptr;
    c->stage_bytes = 0;
    c->env_cache[0] = c->env_cache[1] = nullptr;
    c->env_cache_bytes[0] = c->env_cache_bytes[1] = 0;
    c->ls_state = nullptr;
    c->ls_free = nullptr;
    c->copy_stream = nullptr;
    c->mu = new std::recursive_mutex();
    {
        cudaStream_t cs = nullptr;
        if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) == cudaSuccess) c->copy_stream = cs;
    }
    if (cudaMalloc(&c->d_counters, 256 * sizeof(unsigned int)) != cudaSuccess) {
        hl_set_error("hl_ctx_create: cudaMalloc failed");
        delete c;
        return 1;
    }
    cudaMemset(c->d_counters, 0, 256 * sizeof(unsigned int));
    *out = c;
    return 0;
}

extern "C" void hl_ctx_destroy(hl_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->astar_ws) cudaFree(ctx->astar_ws);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    for (int k = 0; k < 2; ++k) if (ctx->env_cache[k]) cudaFree(ctx->env_cache[k]);
    if (ctx->ls_state && ctx->ls_free) ctx->ls_free(ctx->ls_state);
    if (ctx->copy_stream) cudaStreamDestroy((cudaStream_t)ctx->copy_stream);
    delete (std::recursive_mutex*)ctx->mu;
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

// One device block + one pinned staging block per upload: sizes are computed in a first pass, the second
// pass writes every pool straight into the (context-cached, grow-only) pinned buffer, and ONE
// cudaMemcpyAsync moves it.  (The first version used 15 std::vectors with push_back and 15 pageable
// copies: 37 ms for 4096 environments, as long as the search itself.)
struct PoolLayout {
    size_t off[15];
    size_t total;
};
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int hl_env_upload(hl_ctx* ctx, const HlEnvHost* h, int32_t n_env, hl_env_batch** out) {
    if (!ctx || !h || !out || n_env <= 0) { hl_set_error("hl_env_upload: bad arguments"); return 1; }
    HL_CUDA_OK(cudaSetDevice(ctx->device));
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
    const size_t bytes[15] = {
        sizeof(EnvDesc) * (size_t)n_env,
        sizeof(float) * HL_OBS32_STRIDE * n_obs, sizeof(double) * 8 * n_obs,
        sizeof(float) * HL_FIELD32_STRIDE * n_field, sizeof(double) * 2 * n_field,
        sizeof(float) * 4 * n_seg, sizeof(double) * 4 * n_seg, sizeof(double) * n_seg,
        sizeof(double) * 2 * HL_CAPSULE_VERTS * n_seg, sizeof(double) * 2 * n_crit,
        sizeof(double) * n_guide, sizeof(double) * n_guide, sizeof(double) * n_guide, sizeof(double) * n_guide,
        sizeof(double) * 4 * n_aux};
    PoolLayout L;
    L.total = 0;
    for (int k = 0; k < 15; ++k) { L.off[k] = L.total; L.total += al256(bytes[k] ? bytes[k] : 1); }
    if (L.total > ctx->stage_bytes) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr; ctx->stage_bytes = 0;
        HL_CUDA_OK(cudaMallocHost(&ctx->stage, L.total));
        ctx->stage_bytes = L.total;
    }
    char* S = (char*)ctx->stage;
    EnvDesc* desc = (EnvDesc*)(S + L.off[0]);
    float* obs32 = (float*)(S + L.off[1]);   double* obs64 = (double*)(S + L.off[2]);
    float* field32 = (float*)(S + L.off[3]); double* field64 = (double*)(S + L.off[4]);
    float* seg32 = (float*)(S + L.off[5]);   double* seg64 = (double*)(S + L.off[6]);
    double* seg_len = (double*)(S + L.off[7]); double* seg_poly = (double*)(S + L.off[8]);
    double* crit64 = (double*)(S + L.off[9]);
    double* gx = (double*)(S + L.off[10]); double* gy = (double*)(S + L.off[11]);
    double* gyaw = (double*)(S + L.off[12]); double* gs = (double*)(S + L.off[13]);
    double* aux64 = (double*)(S + L.off[14]);
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
        // origin of the float32 frame: centre of the bounding box of all geometry
        double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
        auto acc = [&](const double* p, int npts) {
            for (int i = 0; i < npts; ++i)
                for (int c = 0; c < 2; ++c) { lo[c] = fmin(lo[c], p[2 * i + c]); hi[c] = fmax(hi[c], p[2 * i + c]); }
        };
        acc(E.obs_xy, 4 * E.n_obs);
        acc(E.field_xy, E.n_field);
        acc(E.seg_xy, 2 * E.n_seg);
        if (lo[0] > hi[0]) { lo[0] = lo[1] = hi[0] = hi[1] = 0.0; }
        D.origin[0] = 0.5 * (lo[0] + hi[0]);
        D.origin[1] = 0.5 * (lo[1] + hi[1]);
        double extent = 0.5 * fmax(hi[0] - lo[0], hi[1] - lo[1]) * 1.4142135623730951;
        double foot = hypot(fmax(fabs(E.body_ext[0]), fabs(E.body_ext[1])), fmax(fabs(E.body_ext[2]), fabs(E.body_ext[3])));
        for (int a = 0; a < E.n_aux; ++a) {
            const double* x = E.aux_ext + 4 * a;
            foot = fmax(foot, hypot(fmax(fabs(x[0]), fabs(x[1])), fmax(fabs(x[2]), fabs(x[3]))));
        }
        double reach = extent + 6.0 + foot + 4.0;            // lane radius 6 m (reference_line_heuristic.py:66) + margin
        D.reach = (float)reach;
        D.eps = (float)(32.0 * 1.1920929e-07 * fmax(reach, 8.0));
        D.all_rect = 1; D.pad0 = 0;
        if (E.n_obs) memcpy(obs64 + 8 * c_obs, E.obs_xy, sizeof(double) * 8 * E.n_obs);
        for (int k = 0; k < E.n_obs; ++k) {
            const double* V = E.obs_xy + 8 * k;
            float* o = obs32 + HL_OBS32_STRIDE * (c_obs + k);
            for (int i = 0; i < 4; ++i) {
                o[2 * i] = (float)(V[2 * i] - D.origin[0]);
                o[2 * i + 1] = (float)(V[2 * i + 1] - D.origin[1]);
            }
            for (int i = 0; i < 4; ++i) {
                int j = (i + 1) & 3;
                double ex = V[2 * j] - V[2 * i], ey = V[2 * j + 1] - V[2 * i + 1];
                double ln = sqrt(ex * ex + ey * ey);
                double nx = ln > 0 ? ey / ln : 0.0, ny = ln > 0 ? -ex / ln : 0.0;
                double cc = nx * (V[2 * i] - D.origin[0]) + ny * (V[2 * i + 1] - D.origin[1]);
                o[8 + 3 * i] = (float)nx; o[9 + 3 * i] = (float)ny; o[10 + 3 * i] = (float)cc;
            }
            // box form when the quad is a rectangle (tree rows, obstacle squares): centre, unit axis, half extents
            double e0x = V[2] - V[0], e0y = V[3] - V[1], e1x = V[4] - V[2], e1y = V[5] - V[3];
            double e2x = V[6] - V[4], e2y = V[7] - V[5], e3x = V[0] - V[6], e3y = V[1] - V[7];
            double l0 = sqrt(e0x * e0x + e0y * e0y), l1 = sqrt(e1x * e1x + e1y * e1y);
            double scale = fmax(l0, l1);
            bool para = fabs(e0x + e2x) + fabs(e0y + e2y) + fabs(e1x + e3x) + fabs(e1y + e3y) <= 1e-9 * scale;
            bool perp = l0 > 0 && l1 > 0 && fabs(e0x * e1x + e0y * e1y) <= 1e-9 * l0 * l1;
            bool is_rect = para && perp;
            if (!is_rect) D.all_rect = 0;
            o[20] = is_rect ? 1.0f : 0.0f;
            o[21] = (float)(0.25 * (V[0] + V[2] + V[4] + V[6]) - D.origin[0]);
            o[22] = (float)(0.25 * (V[1] + V[3] + V[5] + V[7]) - D.origin[1]);
            o[23] = (float)(is_rect ? e0x / l0 : 1.0); o[24] = (float)(is_rect ? e0y / l0 : 0.0);
            o[25] = (float)(0.5 * l0); o[26] = (float)(0.5 * l1); o[27] = 0.0f;
        }
        if (E.n_field) memcpy(field64 + 2 * c_field, E.field_xy, sizeof(double) * 2 * E.n_field);
        {
            double area2 = 0.0;
            for (int i = 0; i < E.n_field; ++i) {
                int j = (i + 1 == E.n_field) ? 0 : i + 1;
                area2 += E.field_xy[2 * i] * E.field_xy[2 * j + 1] - E.field_xy[2 * j] * E.field_xy[2 * i + 1];
            }
            const double orient = area2 < 0 ? -1.0 : 1.0;
            for (int i = 0; i < E.n_field; ++i) {
                int j = (i + 1 == E.n_field) ? 0 : i + 1;
                double Ax = E.field_xy[2 * i] - D.origin[0], Ay = E.field_xy[2 * i + 1] - D.origin[1];
                double Bx = E.field_xy[2 * j] - D.origin[0], By = E.field_xy[2 * j + 1] - D.origin[1];
                double ex = Bx - Ax, ey = By - Ay, ln = sqrt(ex * ex + ey * ey);
                double nx = ln > 0 ? orient * ey / ln : 0.0, ny = ln > 0 ? -orient * ex / ln : 0.0;
                float* f = field32 + HL_FIELD32_STRIDE * (c_field + i);
                f[0] = (float)Ax; f[1] = (float)Ay; f[2] = (float)(Bx - Ax); f[3] = (float)(By - Ay);
                f[4] = (float)nx; f[5] = (float)ny; f[6] = (float)(nx * Ax + ny * Ay);
                f[7] = (float)By;                                   // == the next record's (float)Ay, bit for bit
                f[8] = (float)(-ny * Ax + nx * Ay); f[9] = (float)(-ny * Bx + nx * By);
                f[10] = 0.0f; f[11] = 0.0f;
            }
        }
        if (E.n_seg) {
            memcpy(seg64 + 4 * c_seg, E.seg_xy, sizeof(double) * 4 * E.n_seg);
            memcpy(seg_len + c_seg, E.seg_len, sizeof(double) * E.n_seg);
            memcpy(seg_poly + 2 * HL_CAPSULE_VERTS * c_seg, E.seg_poly, sizeof(double) * 2 * HL_CAPSULE_VERTS * E.n_seg);
            for (int i = 0; i < 4 * E.n_seg; ++i) seg32[4 * c_seg + i] = (float)(E.seg_xy[i] - D.origin[i & 1]);
        }
        if (E.n_crit) memcpy(crit64 + 2 * c_crit, E.crit_xy, sizeof(double) * 2 * E.n_crit);
        for (int i = 0; i < E.n_guide; ++i) {
            gx[c_guide + i] = E.guide[4 * i]; gy[c_guide + i] = E.guide[4 * i + 1];
            gyaw[c_guide + i] = E.guide[4 * i + 2]; gs[c_guide + i] = E.guide[4 * i + 3];
        }
        if (E.n_aux) memcpy(aux64 + 4 * c_aux, E.aux_ext, sizeof(double) * 4 * E.n_aux);
    }
    };
    {
        int nt = (int)std::thread::hardware_concurrency();
        // one process per GPU: share the host cores between the ranks of this node (torchrun exports LOCAL_WORLD_SIZE)
        const char* lws = getenv("LOCAL_WORLD_SIZE");
        const int ranks = lws ? atoi(lws) : 1;
        if (ranks > 1) nt /= ranks;
        nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
        if (n_env < 256) nt = 1;
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
    if (cudaMemcpyAsync(d, S, L.total, cudaMemcpyHostToDevice, cs) != cudaSuccess || cudaStreamSynchronize(cs) != cudaSuccess) {
        hl_env_free(b); hl_set_error("hl_env_upload: copy failed"); return 1;
    }
    b->dev.desc = (const EnvDesc*)(d + L.off[0]);
    b->dev.obs32 = (const float*)(d + L.off[1]);   b->dev.obs64 = (const double*)(d + L.off[2]);
    b->dev.field32 = (const float*)(d + L.off[3]); b->dev.field64 = (const double*)(d + L.off[4]);
    b->dev.seg32 = (const float*)(d + L.off[5]);   b->dev.seg64 = (const double*)(d + L.off[6]);
    b->dev.seg_len = (const double*)(d + L.off[7]); b->dev.seg_poly = (const double*)(d + L.off[8]);
    b->dev.crit64 = (const double*)(d + L.off[9]);
    b->dev.guide_x = (const double*)(d + L.off[10]); b->dev.guide_y = (const double*)(d + L.off[11]);
    b->dev.guide_yaw = (const double*)(d + L.off[12]); b->dev.guide_s = (const double*)(d + L.off[13]);
    b->dev.aux64 = (const double*)(d + L.off[14]);
    *out = b;
    return 0;
}
