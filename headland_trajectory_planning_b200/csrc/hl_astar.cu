// hl_astar.cu -- K4: batched Hybrid A* warm-start search, one CTA per scenario.
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
#include <cstring>
#include "hl_geom.cuh"
#include "hl_rs.cuh"

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
};

__host__ __device__ inline size_t as_align(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ __device__ inline size_t as_ws_bytes(int cap, int hsize, int max_nodes) {
    size_t b = 0;
    b += 4 * as_align(sizeof(double) * cap);           // nx ny nyaw ng
    b += as_align(sizeof(long long) * cap);            // nkey
    b += 3 * as_align(sizeof(int) * cap);              // nparent nheap nhpos
    b += 3 * as_align(cap);                            // nprim nsteps nstate
    b += as_align(sizeof(long long) * hsize) + as_align(sizeof(int) * hsize);
    b += as_align(sizeof(double) * cap) + as_align(sizeof(int) * cap);
    b += as_align(sizeof(int) * (max_nodes + 4));
    return b;
}

__device__ inline AsWs as_carve(char* base, int cap, int hsize, int max_nodes) {
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
__device__ __forceinline__ void heap_swap(const AsWs& w, int i, int j) {
    double pi = w.hprio[i], pj = w.hprio[j];
    int si = w.hslot[i], sj = w.hslot[j];
    w.hprio[i] = pj; w.hslot[i] = sj; w.nheap[sj] = i;
    w.hprio[j] = pi; w.hslot[j] = si; w.nheap[si] = j;
}

__device__ __noinline__ void heap_decrease_key(const AsWs& w, int i) {
    while (i) {
        int parent = (i - 1) >> 1;
        if (w.hprio[parent] < w.hprio[i]) break;
        heap_swap(w, i, parent);
        i = parent;
    }
}

__device__ __noinline__ int heap_popitem(const AsWs& w, int& n) {
    int top = w.hslot[0];
    --n;
    if (n > 0) {
        w.hprio[0] = w.hprio[n]; w.hslot[0] = w.hslot[n]; w.nheap[w.hslot[0]] = 0;
        int i = 0;
        while (true) {
            int l = (i << 1) + 1, r = (i + 1) << 1, low = i;
            if (l < n && w.hprio[l] < w.hprio[i]) low = l;
            if (r < n && w.hprio[r] < w.hprio[low]) low = r;
            if (low == i) break;
            heap_swap(w, i, low);
            i = low;
        }
    }
    w.nheap[top] = -1;
    return top;
}

__device__ __noinline__ void heap_set(const AsWs& w, int& n, int slot, double prio) {
    if (w.nheap[slot] >= 0) {                      // __setitem__ on an existing key: pop(key) first
        int i = w.nheap[slot];
        while (i) {                                // __delitem__: bubble to the root unconditionally
            int parent = (i - 1) >> 1;
            heap_swap(w, i, parent);
            i = parent;
        }
        heap_popitem(w, n);
    }
    int i = n++;
    w.hprio[i] = prio; w.hslot[i] = slot; w.nheap[slot] = i;
    heap_decrease_key(w, i);
}

struct AsSmem {
    // scenario
    double start[3], goal[3];
    int env, scen;
    long long start_key, goal_key;
    // search state (owned by thread 0)
    int n_nodes, heap_n, counter, n_closed;
    int status, arrival, rs_word;
    double goal_cost;
    // current node
    int cur; double cx, cy, cyaw, cg; int cprim; int nsteps;
    int stop_flag;
    // Reeds-Shepp shot
    double rs_lens[HL_RS_CANDIDATES][HL_RS_MAX_SEGS];
    double rs_L[HL_RS_CANDIDATES], rs_prio[HL_RS_CANDIDATES];
    int rs_acc[HL_RS_CANDIDATES], rs_order[HL_RS_CANDIDATES];
    unsigned char rs_valid[HL_RS_CANDIDATES + 2];
    unsigned char rs_accept[HL_RS_CANDIDATES + 2];
    double rs_Lc[HL_RS_CANDIDATES];
    RsProblem rs_prob;
    int rs_n, rs_pick;
    int vote[2][AS_WARPS];        // per-warp (hit | ambiguous << 1) bits of the current sample chunk
    RsPlan plans[AS_MAX_PLANS];
    RsPlan plan_tmp;
    // primitives
    double tx[HL_MAX_PRIMS][AS_ROLL], ty[HL_MAX_PRIMS][AS_ROLL];    // terms, then positions
    double pyaw[HL_MAX_PRIMS][AS_ROLL];                             // pose yaw (yaws[j+1])
    unsigned char pamb[HL_MAX_PRIMS][AS_ROLL];
    int phit[HL_MAX_PRIMS], pany_amb[HL_MAX_PRIMS];
    double pg[HL_MAX_PRIMS], pprio[HL_MAX_PRIMS];
    long long pkey[HL_MAX_PRIMS];
    int pkey_ok[HL_MAX_PRIMS];
    // stats
    unsigned long long n_checks, n_exact;
    long long t_last, t_phase[AS_N_PHASES];
    // backtrack
    int chain_len;
    long long path_off;
    int path_len;
};

// Ternary footprint status of one pose (body only): HL_FREE / HL_HIT / HL_AMBIG(+mask)
__device__ HL_CODE int pose_filter(const EnvDesc& D, const EnvSmem& E, double x, double y, double yaw,
                                           unsigned flags, unsigned* amb) {
    float px = (float)(x - D.origin[0]), py = (float)(y - D.origin[1]);
    if (fabsf(px) > E.reach || fabsf(py) > E.reach || !(fabs(yaw) < 1e6)) { *amb = flags; return HL_AMBIG; }
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
    for (int i = lane; i < n; i += 32) {
        double dx = gx[i] - x, dy = gy[i] - y;
        best = fmin(best, dx * dx + dy * dy);
    }
    for (int o = 16; o; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
    // pass 2: exact hypot on the near-minimal candidates, first minimum wins (np.argmin)
    const double thr = best * (1.0 + 1e-9) + 1e-300;
    double bh = INFINITY;
    int bi = 0x7fffffff;
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
__device__ HL_CODE double rollout_yaw(double init_yaw, double stop, double step, double delta, int div, int i) {
    double v;
    if (i == div) v = stop;                                     // y[-1] = stop
    else if (step != 0.0) v = xadd(xmul((double)i, step), init_yaw);
    else v = xadd(xmul(xdiv((double)i, (double)div), delta), init_yaw);
    return angle_wrap(v);
}

__global__ void __launch_bounds__(AS_THREADS, AS_MIN_CTAS)
k_hybrid_astar(EnvBatchDev eb, const HlScenario* __restrict__ scen, int n_scen, AsParams P, char* ws_base,
               size_t ws_stride, unsigned int* work_counter, HlPlanResult* __restrict__ results,
               int32_t* __restrict__ expanded_keys, double* __restrict__ path_x, double* __restrict__ path_y,
               double* __restrict__ path_yaw, double* __restrict__ path_k, int8_t* __restrict__ path_dir,
               long long path_capacity, unsigned long long* path_cursor, unsigned long long* phase_cycles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AsSmem& S = *reinterpret_cast<AsSmem*>(smem_raw);
    float* env_sm = reinterpret_cast<float*>(smem_raw + ((sizeof(AsSmem) + 15) & ~(size_t)15));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const AsWs W = as_carve(ws_base + (size_t)blockIdx.x * ws_stride, P.cap_nodes, P.hash_size, P.max_nodes);
    const int hmask = P.hash_size - 1;
    const unsigned FLAGS = HL_CHECK_OBSTACLES | HL_CHECK_BOUNDARY | HL_CHECK_LANE;
    const int env_sm_floats = AS_ENV_FLOATS;

    // hash table starts empty; afterwards only the used positions are reset
    for (int i = tid; i < P.hash_size; i += AS_THREADS) W.hkey[i] = KEY_EMPTY;
    __syncthreads();

    while (true) {
        if (tid == 0) S.scen = (int)atomicAdd(work_counter, 1u);
        __syncthreads();
        const int sc = S.scen;
        if (sc >= n_scen) break;
        if (tid == 0) {
            const HlScenario s = scen[sc];
            S.env = s.env_id;
            for (int k = 0; k < 3; ++k) { S.start[k] = s.start[k]; S.goal[k] = s.goal[k]; }
            S.n_nodes = 0; S.heap_n = 0; S.counter = 0; S.n_closed = 0;
            S.status = -1; S.arrival = 0; S.rs_word = -1; S.goal_cost = 0.0;
            S.n_checks = 0; S.n_exact = 0; S.path_len = 0; S.path_off = 0; S.stop_flag = 0;
            for (int k = 0; k < AS_N_PHASES; ++k) S.t_phase[k] = 0;
            S.t_last = clock64();
        }
        __syncthreads();
        const EnvDesc& D = eb.desc[S.env];
        EnvSmem E;
        bool staged;
        stage_env(eb, D, env_sm, env_sm_floats, E, staged);
        __syncthreads();

        // ---- start / goal feasibility (:76-80, :516-519) and start node (:500-510)
        if (wid == 0) {
            int bad = 0;
            if (lane < 2) {
                const double* q = lane == 0 ? S.start : S.goal;
                unsigned amb = FLAGS;
                int r = pose_filter(D, E, q[0], q[1], q[2], FLAGS, &amb);
                bad = (r == HL_HIT) || (r == HL_AMBIG && pose_exact(eb, D, q[0], q[1], q[2], amb));
            }
            bad = __any_sync(0xffffffffu, bad);
            double h = warp_state_cost(eb, D, S.start[0], S.start[1], S.start[2], lane);
            if (lane == 0) {
                int ix, iy, iw;
                long long sk = 0, gk = 0;
                bool ok = make_key(S.start[0], S.start[1], S.start[2], P.res, P.yaw_res, ix, iy, iw, sk);
                ok = make_key(S.goal[0], S.goal[1], S.goal[2], P.res, P.yaw_res, ix, iy, iw, gk) && ok;
                S.start_key = sk; S.goal_key = gk;
                if (!ok) S.status = HL_STATUS_CAPACITY;
                else if (bad) S.status = HL_STATUS_START_GOAL_BLOCKED;
                else {
                    W.nx[0] = S.start[0]; W.ny[0] = S.start[1]; W.nyaw[0] = S.start[2]; W.ng[0] = 0.0;
                    W.nkey[0] = sk; W.nparent[0] = 0; W.nprim[0] = -1; W.nsteps[0] = 0; W.nstate[0] = 0;
                    W.nheap[0] = -1;
                    int pos;
                    hash_find(W, hmask, sk, &pos);
                    W.hkey[pos] = sk; W.hval[pos] = 0; W.nhpos[0] = pos;
                    S.n_nodes = 1;
                    double prio = xmul(P.hybrid_cost, h);
                    prio = (prio > 0.0) ? prio : 0.0;           // max(start.cost = 0, 50*h)
                    heap_set(W, S.heap_n, 0, prio);
                }
            }
        }
        __syncthreads();

        TICK(PH_SETUP);
        // =============================== main loop (:525-596) ===============================
        // Control flow is decided ONLY by reads of S.status that directly follow a barrier, and thread 0
        // never rewrites S.status between such a read and the next barrier -- otherwise a late warp could
        // see the new value, leave the loop alone and desynchronise the CTA's barriers.
        while (true) {
            if (tid == 0 && S.status < 0) {
                if (S.counter > P.max_nodes) S.status = HL_STATUS_MAX_NODES;
                else {
                    S.counter += 1;
                    if (S.heap_n == 0) S.status = HL_STATUS_OPEN_EMPTY;
                    else {
                        int cur = heap_popitem(W, S.heap_n);
                        W.nstate[cur] = 1;
                        W.corder[S.n_closed++] = cur;
                        S.cur = cur; S.cx = W.nx[cur]; S.cy = W.ny[cur]; S.cyaw = W.nyaw[cur]; S.cg = W.ng[cur];
                        S.cprim = W.nprim[cur];
                        S.rs_pick = -1;
                        const double q0n[3] = {S.cx, S.cy, S.cyaw};
                        S.rs_prob = rs_normalise(q0n, S.goal, P.maxc);      // generate_path (:565-572), once per pop
                    }
                }
            }
            __syncthreads();
            if (S.status >= 0) break;
            TICK(PH_POP);

            // ---- analytic shot: 46 candidate words (:249-258)
            {
                const double q0[3] = {S.cx, S.cy, S.cyaw};
                if (tid < HL_RS_CANDIDATES) {
                    double l[HL_RS_MAX_SEGS] = {0, 0, 0, 0, 0};
                    bool ok = rs_candidate(tid, S.rs_prob, l);
                    S.rs_valid[tid] = ok ? 1 : 0;
                    for (int k = 0; k < HL_RS_MAX_SEGS; ++k) S.rs_lens[tid][k] = l[k];
                }
                __syncthreads();
                TICK(PH_RS_CAND);
                if (tid < RS_N_GROUPS) rs_select_group(tid, S.rs_valid, S.rs_lens, S.rs_accept, S.rs_Lc);
                __syncthreads();
                if (tid == 0) {
                    int m = rs_select_compact(S.rs_accept, S.rs_Lc, S.rs_acc, S.rs_L);
                    if (m < 0) { S.status = HL_STATUS_RS_ASSERT; m = 0; }
                    S.rs_n = m;
                    for (int k = 0; k < m; ++k)
                        S.rs_prio[k] = rs_path_cost(S.cg, S.rs_acc[k], S.rs_lens[S.rs_acc[k]], P.max_steer,
                                                    P.reverse_cost, P.dir_change_cost, P.steer_cost);
                    if (m > 0) heapdict_order(S.rs_prio, m, S.rs_order);
                }
                __syncthreads();
                if (S.status >= 0) break;
                TICK(PH_RS_SELECT);
                const int m = S.rs_n;
                const double stepn = xmul(P.res, P.maxc);
                const double cq = m_cos(-q0[2]), sq = m_sin(-q0[2]);
                // sampling plans of the first AS_MAX_PLANS words in pop order, one thread each
                if (tid < m && tid < AS_MAX_PLANS) {
                    int c = S.rs_acc[S.rs_order[tid]];
                    rs_make_plan(c, S.rs_lens[c], P.maxc, stepn, S.plans[tid]);
                    rs_plan_world32(S.plans[tid], q0, cq, sq, D.origin);
                }
                __syncthreads();
                TICK(PH_RS_PLAN);
                // sampled poses are float32 for the filter (1 sincosf per pose); their error (~1e-5 m) widens the band
                EnvSmem Ers = E;
                Ers.eps = E.eps + 6e-5f;
                const float inv_maxc = (float)(1.0 / P.maxc);
                int vb = 0;
                for (int r = 0; r < m; ++r) {
                    const int k = S.rs_order[r];
                    const int c = S.rs_acc[k];
                    if (r >= AS_MAX_PLANS) {
                        if (tid == 0) {
                            rs_make_plan(c, S.rs_lens[c], P.maxc, stepn, S.plan_tmp);
                            rs_plan_world32(S.plan_tmp, q0, cq, sq, D.origin);
                        }
                        __syncthreads();
                    }
                    const RsPlan& plan = (r < AS_MAX_PLANS) ? S.plans[r] : S.plan_tmp;
                    const int npts = plan.npts;
                    int infeasible = 0;
                    for (int base = 0; base < npts && !infeasible; base += AS_THREADS) {
                        const int j = base + tid;
                        int st = HL_FREE;
                        unsigned amb = 0;
                        if (j < npts) {
                            float fx, fy, fc, fs;
                            rs_sample_world32(plan, j, inv_maxc, fx, fy, fc, fs);
                            if (fabsf(fx) > Ers.reach || fabsf(fy) > Ers.reach) { st = HL_AMBIG; amb = FLAGS; }
                            else st = filter_part(Ers, fx, fy, fc, fs, Ers.ext, FLAGS, &amb);
                        }
                        // one barrier per chunk: warps publish (any hit | any ambiguous << 1)
                        const unsigned hitm = __ballot_sync(0xffffffffu, st == HL_HIT);
                        const unsigned ambm = __ballot_sync(0xffffffffu, st == HL_AMBIG);
                        if (lane == 0) S.vote[vb][wid] = (hitm ? 1 : 0) | (ambm ? 2 : 0);
                        __syncthreads();
                        int bits = 0;
#pragma unroll
                        for (int w = 0; w < AS_WARPS; ++w) bits |= S.vote[vb][w];
                        vb ^= 1;
                        infeasible = bits & 1;
                        if (!infeasible && (bits & 2)) {          // float64 sample + exact predicate, ambiguous poses only
                            int bad = 0;
                            if (st == HL_AMBIG) {
                                double lx, ly, lyaw, wx, wy, wyaw;
                                int cs, dir;
                                rs_sample_local(plan, j, P.maxc, lx, ly, lyaw, cs, dir);
                                rs_to_world(q0, cq, sq, lx, ly, lyaw, wx, wy, wyaw);
                                bad = pose_exact(eb, D, wx, wy, wyaw, amb) ? 1 : 0;
                                atomicAdd(&S.n_exact, 1ULL);
                            }
                            infeasible = __syncthreads_or(bad);
                        }
                        if (tid == 0) S.n_checks += (unsigned long long)min(AS_THREADS, npts - base);
                    }
                    const bool short_enough = xdiv(S.rs_L[k], P.maxc) < P.min_len_goal;     // path.L < MIN_LENGTH_TO_GOAL
                    if (!infeasible && short_enough) {
                        if (tid == 0) { S.rs_pick = r; S.arrival = 1; S.rs_word = c; S.goal_cost = S.rs_prio[k]; }
                        break;
                    }
                    if (r + 1 >= AS_MAX_PLANS) __syncthreads();      // plan_tmp is rewritten next round
                }
                __syncthreads();
                TICK(PH_RS_SAMPLE);
            }
            // ---- tolerance arrival (:464-495) overrides the shot
            if (tid == 0) {
                double xd = fabs(xsub(S.cx, S.goal[0])), yd = fabs(xsub(S.cy, S.goal[1]));
                double wd = fabs(angle_wrap(xsub(S.cyaw, S.goal[2])));
                if (xd < P.res && yd < P.res && wd < P.yaw_res) { S.arrival = 2; S.goal_cost = S.cg; S.rs_word = -1; }
                if (S.arrival) S.status = HL_STATUS_OK;
                else {
                    // ---- primitive expansion (:558-596): search length of this node (get_search_length, :368)
                    int seg = exact_search_segment(eb, D, S.cx, S.cy);
                    double len = seg < 0 ? D.default_len : eb.seg_len[D.seg_off + seg];
                    S.nsteps = (int)rint(xdiv(len, P.res));                    // Python round()
                    if (S.nsteps + 1 > HL_MAX_ROLLOUT || S.nsteps < 1) S.status = HL_STATUS_CAPACITY;
                }
            }
            if (tid < HL_MAX_PRIMS) { S.phit[tid] = 0; S.pany_amb[tid] = 0; }
            __syncthreads();
            if (S.status >= 0) break;
            TICK(PH_ARRIVE);
            const int n = S.nsteps, np1 = n + 1;
            const int total = P.n_prims * np1;
            // phase A: per (p, i) displacement terms  (res*m_cos(yaws[i]))*dir, i = 0..n
            for (int idx = tid; idx < total; idx += AS_THREADS) {
                const int p = idx / np1, i = idx - p * np1;
                const double ys = P.yaw_step[p];
                const double init_yaw = angle_wrap(xadd(S.cyaw, ys));
                const double stop = xadd(init_yaw, xmul(ys, (double)(n + 1)));
                const double delta = xsub(stop, init_yaw);
                const double step = xdiv(delta, (double)(n + 1));
                const double yw = rollout_yaw(init_yaw, stop, step, delta, n + 1, i);
                S.tx[p][i] = xmul(xmul(P.res, m_cos(yw)), P.dir[p]);
                S.ty[p][i] = xmul(xmul(P.res, m_sin(yw)), P.dir[p]);
                S.pyaw[p][i] = rollout_yaw(init_yaw, stop, step, delta, n + 1, i + 1);
            }
            __syncthreads();
            // phase B: sequential cumsum per primitive (np.cumsum), then + init
            if (tid < P.n_prims) {
                double ax = 0.0, ay = 0.0;
                for (int i = 0; i < np1; ++i) {
                    ax = (i == 0) ? S.tx[tid][0] : xadd(ax, S.tx[tid][i]);
                    ay = (i == 0) ? S.ty[tid][0] : xadd(ay, S.ty[tid][i]);
                    S.tx[tid][i] = xadd(S.cx, ax);
                    S.ty[tid][i] = xadd(S.cy, ay);
                }
            }
            __syncthreads();
            TICK(PH_ROLLOUT);
            // phase C: float32 filter of every pose
            for (int idx = tid; idx < total; idx += AS_THREADS) {
                const int p = idx / np1, j = idx - p * np1;
                unsigned amb = 0;
                int st = pose_filter(D, E, S.tx[p][j], S.ty[p][j], S.pyaw[p][j], FLAGS, &amb);
                S.pamb[p][j] = (st == HL_AMBIG) ? (unsigned char)amb : 0;
                if (st == HL_HIT) atomicOr(&S.phit[p], 1);
                else if (st == HL_AMBIG) atomicOr(&S.pany_amb[p], 1);
            }
            if (tid == 0) S.n_checks += (unsigned long long)total;
            __syncthreads();
            TICK(PH_FILTER);
            // phase D: float64 escalation only where it can still change the answer
            for (int idx = tid; idx < total; idx += AS_THREADS) {
                const int p = idx / np1, j = idx - p * np1;
                if (S.pamb[p][j] && !S.phit[p]) {
                    atomicAdd(&S.n_exact, 1ULL);
                    if (pose_exact(eb, D, S.tx[p][j], S.ty[p][j], S.pyaw[p][j], S.pamb[p][j])) atomicOr(&S.phit[p], 2);
                }
            }
            __syncthreads();
            TICK(PH_EXACT);
            // phase E: cost, key (thread per primitive) and heuristic (warp per primitive)
            if (tid < P.n_prims && !S.phit[tid]) {
                const int p = tid;
                double len = 0.0;                                   // calculate_path_length (path_utils.py:5-12)
                for (int i = 0; i + 1 < np1; ++i) {
                    double ds = hypot_cr(xsub(S.tx[p][i + 1], S.tx[p][i]), xsub(S.ty[p][i + 1], S.ty[p][i]));
                    len = (i == 0) ? ds : xadd(len, ds);
                }
                double cost = xadd(S.cg, len);                       // simulated_path_cost (:306-329)
                if (P.dir[p] == -1.0) cost = xadd(cost, P.reverse_cost);
                cost = xadd(cost, xmul(P.steer[p], P.steer_cost));
                const double parent_steer = S.cprim < 0 ? 0.0 : P.steer_eff[S.cprim];
                cost = xadd(cost, xmul(fabs(xsub(P.steer[p], parent_steer)), P.delta_steer_cost));
                const double parent_dir = S.cprim < 0 ? 1.0 : P.dir[S.cprim];
                if (parent_dir != P.dir[p]) cost = xadd(cost, P.dir_change_cost);
                S.pg[p] = cost;
                int ix, iy, iw;
                long long key = 0;
                S.pkey_ok[p] = make_key(S.tx[p][n], S.ty[p][n], S.pyaw[p][n], P.res, P.yaw_res, ix, iy, iw, key) ? 1 : 0;
                S.pkey[p] = key;
            }
            for (int p = wid; p < P.n_prims; p += AS_WARPS) {
                if (!S.phit[p]) {
                    double h = warp_state_cost(eb, D, S.tx[p][n], S.ty[p][n], S.pyaw[p][n], lane);
                    if (lane == 0) S.pprio[p] = xmul(P.hybrid_cost, h);
                }
            }
            __syncthreads();
            TICK(PH_COST_HEUR);
            // phase F: merge into the open list in primitive order (:580-596)
            if (tid == 0) {
                for (int p = 0; p < P.n_prims; ++p) {
                    if (S.phit[p]) continue;
                    if (!S.pkey_ok[p]) { S.status = HL_STATUS_CAPACITY; break; }
                    int pos;
                    int slot = hash_find(W, hmask, S.pkey[p], &pos);
                    const double g = S.pg[p];
                    const double prio = (S.pprio[p] > g) ? S.pprio[p] : g;      // max(sim.cost, 50*h)
                    if (slot >= 0) {
                        if (W.nstate[slot] == 1) continue;                         // in closed_set
                        if (!(g < W.ng[slot])) continue;                           // not strictly better
                    } else {
                        if (S.n_nodes >= P.cap_nodes) { S.status = HL_STATUS_CAPACITY; break; }
                        slot = S.n_nodes++;
                        W.hkey[pos] = S.pkey[p]; W.hval[pos] = slot; W.nhpos[slot] = pos;
                        W.nkey[slot] = S.pkey[p]; W.nstate[slot] = 0; W.nheap[slot] = -1;
                    }
                    W.nx[slot] = S.tx[p][n]; W.ny[slot] = S.ty[p][n]; W.nyaw[slot] = S.pyaw[p][n];
                    W.ng[slot] = g; W.nparent[slot] = S.cur; W.nprim[slot] = (signed char)p;
                    W.nsteps[slot] = (signed char)n;
                    heap_set(W, S.heap_n, slot, prio);
                }
            }
            __syncthreads();
            TICK(PH_MERGE);
        }

        // =============================== results ===============================
        // expanded keys in pop order
        {
            int32_t* ek = expanded_keys + (size_t)sc * (P.max_nodes + 2) * 3;
            for (int i = tid; i < S.n_closed; i += AS_THREADS) {
                int ix, iy, iw;
                unpack_key(W.nkey[W.corder[i]], ix, iy, iw);
                ek[3 * i] = ix; ek[3 * i + 1] = iy; ek[3 * i + 2] = iw;
            }
        }
        // path (get_path_from_expanded_nodes, :429-454)
        if (tid == 0 && S.status == HL_STATUS_OK) {
            // Walk cur -> parent -> ... -> start (slot 0 is the only node with the start key: the
            // start cell is closed at the first pop and never re-inserted).  The chain is kept in
            // hslot[] (the heap is dead once the search is over), goal side first.
            // closed_set[goal_key] is the goal node, so a goal in the start's own cell makes the
            // reference's `while current_node_index != start_node_index` loop a no-op: empty path.
            int len = 0, poses = 0, rs_pts = 0;
            bool ok = true;
            if (S.goal_key != S.start_key) {
                for (int node = S.cur; node != 0; node = W.nparent[node]) {
                    if (len >= P.cap_nodes) { ok = false; break; }
                    W.hslot[len++] = node;
                    poses += W.nsteps[node] + 1;
                }
                if (S.arrival == 1) rs_pts = (S.rs_pick < AS_MAX_PLANS ? S.plans[S.rs_pick] : S.plan_tmp).npts;
            }
            S.chain_len = len;
            S.path_len = poses + rs_pts;
            if (!ok || S.path_len > P.max_path_poses) { S.status = HL_STATUS_CAPACITY; S.path_len = 0; }
            else if (S.path_len > 0) {
                unsigned long long off = atomicAdd(path_cursor, (unsigned long long)S.path_len);
                if ((long long)(off + S.path_len) > path_capacity) { S.status = HL_STATUS_CAPACITY; S.path_len = 0; }
                S.path_off = (long long)off;
            }
        }
        __syncthreads();
        if (S.status == HL_STATUS_OK && S.path_len > 0) {
            // node trajectories, start side first (the chain in hslot[] is goal side first); each
            // thread re-derives its write offset from the step counts of the nodes before it.
            const int len = S.chain_len;
            for (int c = tid; c < len; c += AS_THREADS) {
                const int node = W.hslot[len - 1 - c];          // c-th node from the start side
                long long off = S.path_off;
                for (int q = 0; q < c; ++q) off += W.nsteps[W.hslot[len - 1 - q]] + 1;
                const int par = W.nparent[node];
                const int p = W.nprim[node], n = W.nsteps[node];
                const double ys = P.yaw_step[p];
                const double init_yaw = angle_wrap(xadd(W.nyaw[par], ys));
                const double stop = xadd(init_yaw, xmul(ys, (double)(n + 1)));
                const double delta = xsub(stop, init_yaw);
                const double step = xdiv(delta, (double)(n + 1));
                double ax = 0.0, ay = 0.0;
                for (int i = 0; i <= n; ++i) {
                    const double yw = rollout_yaw(init_yaw, stop, step, delta, n + 1, i);
                    const double txv = xmul(xmul(P.res, m_cos(yw)), P.dir[p]);
                    const double tyv = xmul(xmul(P.res, m_sin(yw)), P.dir[p]);
                    ax = (i == 0) ? txv : xadd(ax, txv);
                    ay = (i == 0) ? tyv : xadd(ay, tyv);
                    path_x[off + i] = xadd(W.nx[par], ax);
                    path_y[off + i] = xadd(W.ny[par], ay);
                    path_yaw[off + i] = rollout_yaw(init_yaw, stop, step, delta, n + 1, i + 1);
                    path_k[off + i] = P.curv[p];
                    path_dir[off + i] = (int8_t)P.dir[p];
                }
            }
            if (S.arrival == 1) {
                const RsPlan& plan = (S.rs_pick < AS_MAX_PLANS) ? S.plans[S.rs_pick] : S.plan_tmp;
                const double q0[3] = {S.cx, S.cy, S.cyaw};
                const double cq = m_cos(-q0[2]), sq = m_sin(-q0[2]);
                const long long off = S.path_off + (S.path_len - plan.npts);
                for (int j = tid; j < plan.npts; j += AS_THREADS) {
                    double lx, ly, lyaw, wx, wy, wyaw;
                    int cs, dir;
                    rs_sample_local(plan, j, P.maxc, lx, ly, lyaw, cs, dir);
                    rs_to_world(q0, cq, sq, lx, ly, lyaw, wx, wy, wyaw);
                    path_x[off + j] = wx; path_y[off + j] = wy; path_yaw[off + j] = wyaw;
                    path_k[off + j] = cs == 0 ? 0.0 : (cs > 0 ? P.maxc : -P.maxc);
                    path_dir[off + j] = (int8_t)dir;
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            HlPlanResult r;
            r.status = S.status;
            r.counter = (S.status == HL_STATUS_START_GOAL_BLOCKED) ? 0 : S.counter;
            r.n_expanded = S.n_closed;
            r.arrival = S.arrival;
            r.path_len = S.path_len;
            r.rs_word = S.rs_word;
            r.path_offset = S.path_off;
            r.goal_cost = S.goal_cost;
            r.n_pose_checks = (long long)S.n_checks;
            r.n_exact = (long long)S.n_exact;
            results[sc] = r;
            long long _n = clock64(); S.t_phase[PH_OUTPUT] += _n - S.t_last;
            { long long tot = 0; for (int k = 0; k < AS_N_PHASES; ++k) tot += S.t_phase[k]; results[sc].cycles = tot; }
            for (int k = 0; k < AS_N_PHASES; ++k) atomicAdd(phase_cycles + k, (unsigned long long)S.t_phase[k]);
        }
        // reset the used hash positions for the next scenario of this CTA
        for (int i = tid; i < S.n_nodes; i += AS_THREADS) W.hkey[W.nhpos[i]] = KEY_EMPTY;
        __syncthreads();
    }
}

static void fill_params(const hl_ctx* ctx, const HlSearchParams* h, AsParams& P) {
    memset(&P, 0, sizeof(P));
    P.res = h->plan_resolution; P.yaw_res = h->yaw_resolution; P.maxc = h->maxc;
    P.max_steer = h->max_steer; P.wheel_base = h->wheel_base; P.n_prims = h->n_prims;
    for (int i = 0; i < HL_MAX_PRIMS; ++i) {
        P.steer[i] = h->prim_steer[i]; P.dir[i] = h->prim_dir[i]; P.yaw_step[i] = h->prim_yaw_step[i];
        P.curv[i] = h->prim_curv[i]; P.steer_eff[i] = h->prim_steer_eff[i];
    }
    P.steer_cost = h->steer_cost; P.delta_steer_cost = h->delta_steer_cost;
    P.dir_change_cost = h->direction_change_cost; P.reverse_cost = h->reverse_cost;
    P.hybrid_cost = h->hybrid_cost; P.min_len_goal = h->min_length_to_goal;
    P.max_nodes = h->max_nodes; P.max_path_poses = h->max_path_poses;
    P.cap_nodes = h->n_prims * (h->max_nodes + 1) + 2;
    int hs = 1024;
    while (hs < 2 * P.cap_nodes) hs <<= 1;
    P.hash_size = hs;
}

static int astar_grid(const hl_ctx* ctx, int n_scen, size_t smem) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hybrid_astar, AS_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)ctx->sm_count * per_sm;
    return (int)(g < n_scen ? g : n_scen);
}

static size_t astar_smem() { return ((sizeof(AsSmem) + 15) & ~(size_t)15) + AS_ENV_FLOATS * sizeof(float); }

extern "C" int hl_astar_phase_cycles(hl_ctx* ctx, uint64_t* h_out, int32_t n, int32_t reset) {
    if (!ctx || !h_out || n < 1 || n > AS_N_PHASES) { hl_set_error("hl_astar_phase_cycles: bad arguments"); return 1; }
    HL_CUDA_OK(cudaSetDevice(ctx->device));
    HL_CUDA_OK(cudaMemcpy(h_out, ctx->d_counters + 16, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
    if (reset) HL_CUDA_OK(cudaMemset(ctx->d_counters + 16, 0, sizeof(uint64_t) * AS_N_PHASES));
    return 0;
}

extern "C" int64_t hl_hybrid_astar_workspace_bytes(const hl_ctx* ctx, const HlSearchParams* h) {
    if (!ctx || !h) return -1;
    AsParams P;
    fill_params(ctx, h, P);
    return (int64_t)as_ws_bytes(P.cap_nodes, P.hash_size, P.max_nodes);
}

extern "C" int hl_hybrid_astar_batch(hl_ctx* ctx, const hl_env_batch* envs, const HlScenario* d_scen,
                                     int32_t n_scen, const HlSearchParams* h_params,
                                     HlPlanResult* d_results, int32_t* d_expanded_keys,
                                     double* d_path_x, double* d_path_y, double* d_path_yaw,
                                     double* d_path_k, int8_t* d_path_dir, int64_t path_capacity,
                                     unsigned long long* d_path_cursor, void* stream) {
    if (!ctx || !envs || !d_scen || !h_params || !d_results || !d_expanded_keys || !d_path_x || !d_path_y ||
        !d_path_yaw || !d_path_k || !d_path_dir || !d_path_cursor || n_scen < 0) {
        hl_set_error("hl_hybrid_astar_batch: bad arguments"); return 1;
    }
    if (n_scen == 0) return 0;
    if (h_params->n_prims < 1 || h_params->n_prims > HL_MAX_PRIMS || h_params->max_nodes < 0) {
        hl_set_error("hl_hybrid_astar_batch: n_prims/max_nodes out of range"); return 1;
    }
    HL_CUDA_OK(cudaSetDevice(ctx->device));
    AsParams P;
    fill_params(ctx, h_params, P);
    const size_t smem = astar_smem();
    HL_CUDA_OK(cudaFuncSetAttribute(k_hybrid_astar, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = astar_grid(ctx, n_scen, smem);
    const size_t stride = as_ws_bytes(P.cap_nodes, P.hash_size, P.max_nodes);
    const size_t need = stride * (size_t)grid;
    if (need > ctx->astar_ws_bytes) {
        if (ctx->astar_ws) cudaFree(ctx->astar_ws);
        ctx->astar_ws = nullptr; ctx->astar_ws_bytes = 0;
        HL_CUDA_OK(cudaMalloc(&ctx->astar_ws, need));
        ctx->astar_ws_bytes = need;
    }
    cudaStream_t st = (cudaStream_t)stream;
    HL_CUDA_OK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned int), st));
    HL_CUDA_OK(cudaMemsetAsync(d_path_cursor, 0, sizeof(unsigned long long), st));
    k_hybrid_astar<<<grid, AS_THREADS, smem, st>>>(envs->dev, d_scen, n_scen, P, (char*)ctx->astar_ws, stride,
                                                   ctx->d_counters, d_results, d_expanded_keys, d_path_x, d_path_y,
                                                   d_path_yaw, d_path_k, d_path_dir, (long long)path_capacity,
                                                   d_path_cursor, (unsigned long long*)(ctx->d_counters + 16));
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}
