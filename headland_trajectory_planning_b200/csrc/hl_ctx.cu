// hl_ctx.cu -- context, error reporting, environment upload (host -> HBM).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
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
    delete ctx;
}

extern "C" int hl_ctx_sm_count(const hl_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

template <typename T>
static int upload(hl_env_batch* b, const std::vector<T>& v, const T** dst) {
    void* d = nullptr;
    size_t bytes = (v.size() ? v.size() : 1) * sizeof(T);
    HL_CUDA_OK(cudaMalloc(&d, bytes));
    b->allocs[b->n_allocs++] = d;
    if (v.size()) HL_CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T*)d;
    return 0;
}

extern "C" void hl_env_free(hl_env_batch* envs) {
    if (!envs) return;
    cudaSetDevice(envs->device);
    for (int i = 0; i < envs->n_allocs; ++i) cudaFree(envs->allocs[i]);
    delete envs;
}

extern "C" int32_t hl_env_count(const hl_env_batch* envs) { return envs ? envs->dev.n_env : 0; }

extern "C" int hl_env_upload(hl_ctx* ctx, const HlEnvHost* h, int32_t n_env, hl_env_batch** out) {
    if (!ctx || !h || !out || n_env <= 0) { hl_set_error("hl_env_upload: bad arguments"); return 1; }
    HL_CUDA_OK(cudaSetDevice(ctx->device));
    std::vector<EnvDesc> desc(n_env);
    std::vector<float> obs32, field32, seg32;
    std::vector<double> obs64, field64, seg64, seg_len, seg_poly, crit64, gx, gy, gyaw, gs, aux64;
    for (int e = 0; e < n_env; ++e) {
        const HlEnvHost& E = h[e];
        EnvDesc& D = desc[e];
        if (E.n_seg > HL_MAX_SEGS) { hl_set_error("hl_env_upload: env %d has %d lane segments (max %d)", e, E.n_seg, HL_MAX_SEGS); return 1; }
        if (E.n_obs < 0 || E.n_field < 0 || E.n_seg < 0 || E.n_guide < 0 || E.n_aux < 0 || E.n_crit < 0) {
            hl_set_error("hl_env_upload: env %d has a negative count", e); return 1;
        }
        D.n_obs = E.n_obs;     D.obs_off = (int)(obs64.size() / 8);
        D.n_field = E.n_field; D.field_off = (int)(field64.size() / 2);
        D.n_seg = E.n_seg;     D.seg_off = (int)(seg64.size() / 4);
        D.n_crit = E.n_crit;   D.crit_off = (int)(crit64.size() / 2);
        D.n_guide = E.n_guide; D.guide_off = (int)gx.size();
        D.n_aux = E.n_aux;     D.aux_off = (int)(aux64.size() / 4);
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
        double foot = 0.0;                                    // farthest footprint corner from the base link
        for (int k = 0; k < 4; k += 2) foot = fmax(foot, hypot(fmax(fabs(E.body_ext[0]), fabs(E.body_ext[1])), fmax(fabs(E.body_ext[2]), fabs(E.body_ext[3]))));
        for (int a = 0; a < E.n_aux; ++a) {
            const double* x = E.aux_ext + 4 * a;
            foot = fmax(foot, hypot(fmax(fabs(x[0]), fabs(x[1])), fmax(fabs(x[2]), fabs(x[3]))));
        }
        double reach = extent + 6.0 + foot + 4.0;            // lane radius 6 m (reference_line_heuristic.py:66) + margin
        D.reach = (float)reach;
        D.eps = (float)(32.0 * 1.1920929e-07 * fmax(reach, 8.0));
        D.all_rect = 1; D.pad0 = 0;
        for (int k = 0; k < E.n_obs; ++k) {
            const double* V = E.obs_xy + 8 * k;
            for (int i = 0; i < 8; ++i) obs64.push_back(V[i]);
            float v32[8];
            for (int i = 0; i < 4; ++i) {
                v32[2 * i] = (float)(V[2 * i] - D.origin[0]);
                v32[2 * i + 1] = (float)(V[2 * i + 1] - D.origin[1]);
            }
            for (int i = 0; i < 8; ++i) obs32.push_back(v32[i]);
            for (int i = 0; i < 4; ++i) {
                int j = (i + 1) & 3;
                double ex = V[2 * j] - V[2 * i], ey = V[2 * j + 1] - V[2 * i + 1];
                double ln = sqrt(ex * ex + ey * ey);
                double nx = ln > 0 ? ey / ln : 0.0, ny = ln > 0 ? -ex / ln : 0.0;
                double cc = nx * (V[2 * i] - D.origin[0]) + ny * (V[2 * i + 1] - D.origin[1]);
                obs32.push_back((float)nx); obs32.push_back((float)ny); obs32.push_back((float)cc);
            }
            // box form when the quad is a rectangle (tree rows, obstacle squares): centre, unit axis, half extents
            {
                double e0x = V[2] - V[0], e0y = V[3] - V[1], e1x = V[4] - V[2], e1y = V[5] - V[3];
                double e2x = V[6] - V[4], e2y = V[7] - V[5], e3x = V[0] - V[6], e3y = V[1] - V[7];
                double l0 = sqrt(e0x * e0x + e0y * e0y), l1 = sqrt(e1x * e1x + e1y * e1y);
                double scale = fmax(l0, l1);
                bool para = fabs(e0x + e2x) + fabs(e0y + e2y) + fabs(e1x + e3x) + fabs(e1y + e3y) <= 1e-9 * scale;
                bool perp = l0 > 0 && l1 > 0 && fabs(e0x * e1x + e0y * e1y) <= 1e-9 * l0 * l1;
                bool is_rect = para && perp;
                double cx = 0.25 * (V[0] + V[2] + V[4] + V[6]) - D.origin[0];
                double cy = 0.25 * (V[1] + V[3] + V[5] + V[7]) - D.origin[1];
                double ax = is_rect ? e0x / l0 : 1.0, ay = is_rect ? e0y / l0 : 0.0;
                if (!is_rect) D.all_rect = 0;
                obs32.push_back(is_rect ? 1.0f : 0.0f);
                obs32.push_back((float)cx); obs32.push_back((float)cy);
                obs32.push_back((float)ax); obs32.push_back((float)ay);
                obs32.push_back((float)(0.5 * l0)); obs32.push_back((float)(0.5 * l1));
                obs32.push_back(0.0f);
            }
        }
        {
            double area2 = 0.0;
            for (int i = 0; i < E.n_field; ++i) {
                int j = (i + 1 == E.n_field) ? 0 : i + 1;
                area2 += E.field_xy[2 * i] * E.field_xy[2 * j + 1] - E.field_xy[2 * j] * E.field_xy[2 * i + 1];
            }
            const double orient = area2 < 0 ? -1.0 : 1.0;
            for (int i = 0; i < E.n_field; ++i) {
                int j = (i + 1 == E.n_field) ? 0 : i + 1;
                field64.push_back(E.field_xy[2 * i]); field64.push_back(E.field_xy[2 * i + 1]);
                double Ax = E.field_xy[2 * i] - D.origin[0], Ay = E.field_xy[2 * i + 1] - D.origin[1];
                double Bx = E.field_xy[2 * j] - D.origin[0], By = E.field_xy[2 * j + 1] - D.origin[1];
                double ex = Bx - Ax, ey = By - Ay, ln = sqrt(ex * ex + ey * ey);
                double nx = ln > 0 ? orient * ey / ln : 0.0, ny = ln > 0 ? -orient * ex / ln : 0.0;
                field32.push_back((float)Ax); field32.push_back((float)Ay);
                field32.push_back((float)(Bx - Ax)); field32.push_back((float)(By - Ay));
                field32.push_back((float)nx); field32.push_back((float)ny);
                field32.push_back((float)(nx * Ax + ny * Ay));
                field32.push_back((float)By);                       // == the next record's (float)Ay, bit for bit
                field32.push_back((float)(-ny * Ax + nx * Ay)); field32.push_back((float)(-ny * Bx + nx * By));
                field32.push_back(0.0f); field32.push_back(0.0f);
            }
        }
        for (int i = 0; i < E.n_seg; ++i) {
            for (int c = 0; c < 4; ++c) {
                seg64.push_back(E.seg_xy[4 * i + c]);
                seg32.push_back((float)(E.seg_xy[4 * i + c] - D.origin[c & 1]));
            }
            seg_len.push_back(E.seg_len[i]);
            for (int c = 0; c < 2 * HL_CAPSULE_VERTS; ++c) seg_poly.push_back(E.seg_poly[2 * HL_CAPSULE_VERTS * i + c]);
        }
        for (int i = 0; i < 2 * E.n_crit; ++i) crit64.push_back(E.crit_xy[i]);
        for (int i = 0; i < E.n_guide; ++i) {
            gx.push_back(E.guide[4 * i]); gy.push_back(E.guide[4 * i + 1]);
            gyaw.push_back(E.guide[4 * i + 2]); gs.push_back(E.guide[4 * i + 3]);
        }
        for (int i = 0; i < 4 * E.n_aux; ++i) aux64.push_back(E.aux_ext[i]);
    }
    hl_env_batch* b = new hl_env_batch();
    b->n_allocs = 0;
    b->device = ctx->device;
    b->dev.n_env = n_env;
    int rc = 0;
    rc |= upload(b, desc, &b->dev.desc);
    rc |= upload(b, obs32, &b->dev.obs32);
    rc |= upload(b, obs64, &b->dev.obs64);
    rc |= upload(b, field32, &b->dev.field32);
    rc |= upload(b, field64, &b->dev.field64);
    rc |= upload(b, seg32, &b->dev.seg32);
    rc |= upload(b, seg64, &b->dev.seg64);
    rc |= upload(b, seg_len, &b->dev.seg_len);
    rc |= upload(b, seg_poly, &b->dev.seg_poly);
    rc |= upload(b, crit64, &b->dev.crit64);
    rc |= upload(b, gx, &b->dev.guide_x);
    rc |= upload(b, gy, &b->dev.guide_y);
    rc |= upload(b, gyaw, &b->dev.guide_yaw);
    rc |= upload(b, gs, &b->dev.guide_s);
    rc |= upload(b, aux64, &b->dev.aux64);
    if (rc) { hl_env_free(b); return 1; }
    *out = b;
    return 0;
}
