// hl_astar_common.cuh -- shared pieces of the K4 search kernels (parameters, per-scenario workspace,
// grid keys, hash set, heapdict replay, pose filter, heuristic, rollout).
//
// Replaces HybridAStarSearch.hybrid_a_star_search (path_planner/hybrid_a_star_search.py:497-607,
// King mode): every popped node first gets a Reeds-Shepp analytic shot over all words
// (:232-287), then the 14 motion primitives are rolled out (:357-410), collision-checked
// against obstacles + field polygon + guide lane (:412-427), costed (:306-329) and merged
// into the open list (:580-596).  The pop->expand->push chain of ONE scenario is sequential;
// the parallelism is (a) scenarios across CTAs (persistent CTAs pull scenario ids from an
// atomic counter), (b) inside an expansion: 46 word solvers, the poses of a word, the
// 14 x (n+1) primitive poses, the guide-point argmin.
//
// Exactness: everything that feeds a discrete decision (rollout, grid keys, g-cost,
// heuristic, priorities, word validity/dedup/cost, heap order, sample counts) is float64 in
// the reference's operation order; the open list replays heapdict's tie behaviour
// (oracle/heapdict_port.py).  Footprint tests go through the float32 filter first and
// escalate to the float64 predicates only inside the error band, and only when no other
// pose of the same path already decided it.
// The search kernel runs many different phases on different warps/CTAs at once; with every helper
// inlined its SASS was 460 KB and 73 % of the non-barrier stall samples were instruction-fetch misses
// (profiles/r1b).  HL_SHARED_CODE makes the heavy helpers out-of-line so the kernel keeps one copy.
#define HL_SHARED_CODE 1
#pragma once
#include <cstring>
#include "hl_geom.cuh"
#include "hl_rs.cuh"
#include "hl_dubins.cuh"

#ifndef AS_THREADS
#define AS_THREADS 128
#endif
#ifndef AS_MIN_CTAS
#define AS_MIN_CTAS 4
#endif
#define AS_WARPS (AS_THREADS / 32)
#define AS_MAX_PLANS 8
#define AS_ROLL (HL_MAX_ROLLOUT + 1)
#define AS_ENV_FLOATS 768       // staged float32 environment (canonical: 8x28 + 16x12 + 4x4 = 432 floats)
#define KEY_EMPTY (-1LL)
// phase timers (cycles of thread 0 between barriers), the device analogue of the reference's three
// accumulating timers (hybrid_a_star_search.py:91-94): summed over scenarios into ctx->d_counters
enum { PH_POP = 0, PH_RS_CAND, PH_RS_SELECT, PH_RS_PLAN, PH_RS_SAMPLE, PH_ARRIVE, PH_ROLLOUT, PH_FILTER, PH_EXACT,
       PH_COST_HEUR, PH_MERGE, PH_SETUP, PH_OUTPUT, AS_N_PHASES };
#define TICK(ph) do { if (tid == 0) { long long _n = clock64(); S.t_phase[ph] += _n - S.t_last; S.t_last = _n; } } while (0)

struct AsParams {
    double res, yaw_res, maxc, max_steer, wheel_base;
    int n_prims;
    double steer[HL_MAX_PRIMS], dir[HL_MAX_PRIMS], yaw_step[HL_MAX_PRIMS], curv[HL_MAX_PRIMS], steer_eff[HL_MAX_PRIMS];
    double steer_cost, delta_steer_cost, dir_change_cost, reverse_cost, hybrid_cost, min_len_goal;
    int max_nodes, max_path_poses;
    int cap_nodes, hash_size;
    int pawn;                 // motion_type "Pawn": forward-only primitives, Dubins goal extension (:184-230, :289-304)
    int dub_cap;              // capacity (samples / course rows) of the Dubins scratch of one scenario slot
};

// per-CTA workspace in global memory (L2 resident while the scenario runs)
struct AsWs {
    double* nx; double* ny; double* nyaw; double* ng;
    long long* nkey;
    int* nparent; int* nheap; int* nhpos;
    signed char* nprim; signed char* nsteps; signed char* nstate;
    long long* hkey; int* hval;
    double* hprio; int* hslot;
    int* corder;
    long long* cref;          // algorithmic primitive-pose checks done BEFORE pop i was expanded
    double* dub;              // Pawn mode: 13 x dub_cap doubles (ts kx ky ks dx dy cp bx by | rx ry ryaw rk), else unused
};

__host__ __device__ inline size_t as_align(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ __device__ inline size_t as_ws_bytes(int cap, int hsize, int max_nodes, int dub_cap = 0) {
    size_t b = 0;
    b += 4 * as_align(sizeof(double) * cap);           // nx ny nyaw ng
    b += as_align(sizeof(long long) * cap);            // nkey
    b += 3 * as_align(sizeof(int) * cap);              // nparent nheap nhpos
    b += 3 * as_align(cap);                            // nprim nsteps nstate
    b += as_align(sizeof(long long) * hsize) + as_align(sizeof(int) * hsize);
    b += as_align(sizeof(double) * cap) + as_align(sizeof(int) * cap);
    b += as_align(sizeof(int) * (max_nodes + 4));
    b += as_align(sizeof(long long) * (max_nodes + 4));
    b += as_align(sizeof(double) * 13 * (size_t)dub_cap);
    return b;
}

__device__ inline AsWs as_carve(char* base, int cap, int hsize, int max_nodes, int dub_cap = 0) {
    AsWs w;
    char* p = base;
    auto take = [&](size_t bytes) { char* r = p; p += as_align(bytes); return r; };
    w.nx = (double*)take(sizeof(double) * cap); w.ny = (double*)take(sizeof(double) * cap);
    w.nyaw = (double*)take(sizeof(double) * cap); w.ng = (double*)take(sizeof(double) * cap);
    w.nkey = (long long*)take(sizeof(long long) * cap);
    w.nparent = (int*)take(sizeof(int) * cap); w.nheap = (int*)take(sizeof(int) * cap);
    w.nhpos = (int*)take(sizeof(int) * cap);
    w.nprim = (signed char*)take(cap); w.nsteps = (signed char*)take(cap); w.nstate = (signed char*)take(cap);
    w.hkey = (long long*)take(sizeof(long long) * hsize); w.hval = (int*)take(sizeof(int) * hsize);
    w.hprio = (double*)take(sizeof(double) * cap); w.hslot = (int*)take(sizeof(int) * cap);
    w.corder = (int*)take(sizeof(int) * (max_nodes + 4));
    w.cref = (long long*)take(sizeof(long long) * (max_nodes + 4));
    w.dub = (double*)take(sizeof(double) * 13 * (size_t)dub_cap);
    return w;
}

// ---- grid key (calculate_node_index, :82-89): Python round() = half-to-even = rint
__device__ __forceinline__ bool make_key(double x, double y, double yaw, double res, double yaw_res,
                                         int& ix, int& iy, int& iyaw, long long& key) {
    double fx = rint(xdiv(x, res)), fy = rint(xdiv(y, res)), fw = rint(xdiv(yaw, yaw_res));
    if (!(fabs(fx) < 8388607.0) || !(fabs(fy) < 8388607.0) || !(fabs(fw) < 127.0)) return false;
    ix = (int)fx; iy = (int)fy; iyaw = (int)fw;
    key = ((long long)(ix + 8388608) << 32) | ((long long)(iy + 8388608) << 8) | (long long)(iyaw + 128);
    return true;
}

__device__ __forceinline__ void unpack_key(long long key, int& ix, int& iy, int& iyaw) {
    ix = (int)(key >> 32) - 8388608;
    iy = (int)((key >> 8) & 0xFFFFFF) - 8388608;
    iyaw = (int)(key & 0xFF) - 128;
}

__device__ __forceinline__ unsigned hash_key(long long k) {
    unsigned long long z = (unsigned long long)k * 0x9E3779B97F4A7C15ULL;
    return (unsigned)(z >> 40);
}

// returns slot or -1; *pos = table position where the key is / would be inserted
__device__ __noinline__ int hash_find(const AsWs& w, int hmask, long long key, int* pos) {
    unsigned h = hash_key(key) & hmask;
    while (true) {
        long long k = w.hkey[h];
        if (k == key) { *pos = (int)h; return w.hval[h]; }
        if (k == KEY_EMPTY) { *pos = (int)h; return -1; }
        h = (h + 1) & hmask;
    }
}

// ---- heapdict replay (oracle/heapdict_port.py) on (hprio, hslot) with nheap[] positions
// heapdict moves entries by swaps; here the moving entry is carried in registers and the entries it passes are
// shifted into the hole -- the arrangement after every operation is the one the swaps produce (comparisons look at the
// priority only, strict `<`, left child first), with half the memory operations on the dependent chain.
// The first HS heap positions (the top log2(HS+1) levels, where every sift spends most of its steps) may live in
// SHARED memory (sp / ss): an LDS is ~30 cycles, a workspace load that misses L1 several hundred, and the sifts are a
// serial chain on one lane (K4 spec variant: pop + merge were 25 % of the expander's time on global memory).
template <int HS> struct HeapRef {
    const AsWs& w; double* sp; int* ss;
    __device__ __forceinline__ double prio(int i) const { return (HS > 0 && i < HS) ? sp[i] : w.hprio[i]; }
    __device__ __forceinline__ int slot(int i) const { return (HS > 0 && i < HS) ? ss[i] : w.hslot[i]; }
    __device__ __forceinline__ void put(int i, double p, int s) const {
        if (HS > 0 && i < HS) { sp[i] = p; ss[i] = s; } else { w.hprio[i] = p; w.hslot[i] = s; }
        w.nheap[s] = i;
    }
};

// _decrease_key (or, with `always`, the unconditional bubbling of __delitem__) of the entry (prio, slot) at position i
template <int HS> __device__ __forceinline__ void heap_up_t(const HeapRef<HS>& H, int i, double prio, int slot, bool always) {
    while (i) {
        const int parent = (i - 1) >> 1;
        const double pp = H.prio(parent);
        if (!always && pp < prio) break;
        H.put(i, pp, H.slot(parent));
        i = parent;
    }
    H.put(i, prio, slot);
}

template <int HS> __device__ __forceinline__ int heap_popitem_t(const HeapRef<HS>& H, int& n) {
    const int top = H.slot(0);
    --n;
    if (n > 0) {
        const double prio = H.prio(n);                 // heap[0] = heap.pop(); _min_heapify(0)
        const int slot = H.slot(n);
        int i = 0;
        while (true) {
            const int l = (i << 1) + 1, r = l + 1;
            int low = i;
            double lowp = prio;
            if (l < n) {
                const double pl = H.prio(l);
                const double pr = (r < n) ? H.prio(r) : 0.0;
                if (pl < lowp) { low = l; lowp = pl; }
                if (r < n && pr < lowp) { low = r; lowp = pr; }
            }
            if (low == i) break;
            H.put(i, lowp, H.slot(low));
            i = low;
        }
        H.put(i, prio, slot);
    }
    H.w.nheap[top] = -1;
    return top;
}

template <int HS> __device__ __forceinline__ void heap_set_t(const HeapRef<HS>& H, int& n, int slot, double prio) {
    const int at = H.w.nheap[slot];
    if (at >= 0) {                                     // __setitem__ on an existing key: pop(key) first
        heap_up_t<HS>(H, at, 0.0, slot, true);         // __delitem__: bubble to the root unconditionally
        heap_popitem_t<HS>(H, n);
    }
    heap_up_t<HS>(H, n++, prio, slot, false);
}

// workspace-only heap (one-warp-per-scenario and level-synchronous variants)
__device__ __noinline__ int heap_popitem(const AsWs& w, int& n) {
    const HeapRef<0> H{w, nullptr, nullptr};
    return heap_popitem_t<0>(H, n);
}
__device__ __noinline__ void heap_set(const AsWs& w, int& n, int slot, double prio) {
    const HeapRef<0> H{w, nullptr, nullptr};
    heap_set_t<0>(H, n, slot, prio);
}

// Ternary footprint status of one pose (body only): HL_FREE / HL_HIT / HL_AMBIG(+mask)
__device__ HL_CODE int pose_filter(const EnvDesc& D, const EnvSmem& E, double x, double y, double yaw,
                                           unsigned flags, unsigned* amb) {
    float px = (float)(x - D.origin[0]), py = (float)(y - D.origin[1]);
    if (fabsf(px) > E.reach || fabsf(py) > E.reach) return far_status(flags, E.n_seg);
    if (!(fabs(yaw) < 1e6) || !(px == px) || !(py == py)) { *amb = flags; return HL_AMBIG; }
    float sf, cf;
    sincosf((float)yaw, &sf, &cf);
    return filter_part(E, px, py, cf, sf, E.ext, flags, amb);
}

__device__ __noinline__ bool pose_exact(const EnvBatchDev& eb, const EnvDesc& D, double x, double y, double yaw,
                                           unsigned amb) {
    Pose64 p;
    p.x = x; p.y = y; p.c = m_cos(yaw); p.s = m_sin(yaw);
    return exact_part_check(p, D.body_ext, eb, D, amb);
}

// calculate_state_cost (reference_line_heuristic.py:131-158) for one pose, one warp.
__device__ __noinline__ double warp_state_cost(const EnvBatchDev& eb, const EnvDesc& D, double x, double y, double yaw, int lane) {
    const double* gx = eb.guide_x + D.guide_off;
    const double* gy = eb.guide_y + D.guide_off;
    const int n = D.n_guide;
    if (n <= 0) return 0.0;
    // pass 1: minimum squared distance (ordering filter only)
    double best = INFINITY;
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
        double dx = gx[i] - x, dy = gy[i] - y;
        best = fmin(best, dx * dx + dy * dy);
    }
    for (int o = 16; o; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
    // pass 2: exact hypot on the near-minimal candidates, first minimum wins (np.argmin)
    const double thr = best * (1.0 + 1e-9) + 1e-300;
    double bh = INFINITY;
    int bi = 0x7fffffff;
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
        double dx = xsub(gx[i], x), dy = xsub(gy[i], y);
        if (dx * dx + dy * dy <= thr) {
            double h = hypot_cr(dx, dy);
            if (h < bh || (h == bh && i < bi)) { bh = h; bi = i; }
        }
    }
    for (int o = 16; o; o >>= 1) {
        double oh = __shfl_xor_sync(0xffffffffu, bh, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oh < bh || (oh == bh && oi < bi)) { bh = oh; bi = oi; }
    }
    double dist = xmul(bh, 100.0);
    double yaw_diff = fabs(angle_wrap(xsub(eb.guide_yaw[D.guide_off + bi], yaw)));
    if (dist > 2.0) dist = 100.0;
    double to_goal = xsub(eb.guide_s[D.guide_off + n - 1], eb.guide_s[D.guide_off + bi]);
    return xadd(xadd(dist, xmul(yaw_diff, 0.2)), xmul(to_goal, 5.0));
}

// One step of the kinematic rollout (kinematic_simulation_node, :366-390): yaws[i] of
// np.linspace(init_yaw, init_yaw + yaw_step*(n+1), n+2) after angle_wrap.
__device__ HL_TINY double rollout_yaw(double init_yaw, double stop, double step, double delta, int div, int i) {
    double v;
    if (i == div) v = stop;                                     // y[-1] = stop
    else if (step != 0.0) v = xadd(xmul((double)i, step), init_yaw);
    else v = xadd(xmul(xdiv((double)i, (double)div), delta), init_yaw);
    return angle_wrap(v);
}

