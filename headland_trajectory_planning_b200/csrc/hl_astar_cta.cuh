// hl_astar_cta.cu -- K4 variant A: one 128-thread CTA per scenario (kept for A/B measurements;
// the shipped variant is hl_astar.cu, one warp per scenario).  See DESIGN.md section 5.
#define HL_SHARED_CODE 1
#include "hl_astar_common.cuh"

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

