// hl_grid.cu -- K6: holonomic grid distance field; K7: occupancy-grid footprint check.
//
// K6 replaces holonomic_costs_with_obstacles (path_planner/utils/a_star_utils.py:75-142).
// The reference runs a binary-heap Dijkstra from the goal cell with edge costs 1 and
// hypot(1,1) accumulated as sequential float64 sums.  Because fl(d + w) is monotone in d,
// its result is the least fixed point of D(v) = min_m fl(D(v - m) + w(m)), which ANY
// relaxation order reaches bit-exactly in float64 -- so the GPU runs a tiled wavefront:
// a CTA pulls a DF_TILE x DF_TILE tile (16 x 16, +1 halo) into shared memory, relaxes it to its local fixed
// point there, writes it back and flags the neighbouring tiles whose halo changed.  Only
// flagged tiles do work in the next launch (the frontier), so the traffic per launch is
// the frontier's, not the map's.  Roofline: nominally HBM, really launch/dependency
// depth.
//
// K7 is the grid-based footprint check the reference left commented out
// (orchard_geometry_environment.py:8,35-43,460-461) on the grid layout of
// occupancy_grid_utils.py:70-101; parity unpinned (no reference implementation).
#include <cstring>
#include "hl_common.cuh"

#ifndef DF_TILE
#define DF_TILE 16                      // tile edge: 16 -> 310 launches of 13 us for the 4096 x 4096 field (4.03 ms), 32 -> 157 of 27 us
#endif                                   // (4.26 ms), 64 -> 79 of 80 us (6.37 ms): short visits and a long frontier list win
#define DF_THREADS 256

struct DfMoves { int n; int di[8]; int dj[8]; double w[8]; };

__global__ void k_df_init(const uint8_t* __restrict__ occ, int W, int H, int gi, int gj, double* __restrict__ out,
                          int* __restrict__ list, int* __restrict__ count, int tiles_i, int tiles_j, int* __restrict__ bad) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)W * H) return;
    int i = (int)(idx / H), j = (int)(idx % H);
    out[idx] = (i == gi && j == gj) ? 0.0 : INFINITY;
    if ((i == 0 || j == 0 || i == W - 1 || j == H - 1) && !occ[idx]) atomicOr(bad, 1);   // open border: aliases reachable
    if (i == gi && j == gj) {
        int ti = i / DF_TILE, tj = j / DF_TILE;
        for (int a = -1; a <= 1; ++a)
            for (int b = -1; b <= 1; ++b) {
                int x = ti + a, y = tj + b;
                if (x >= 0 && y >= 0 && x < tiles_i && y < tiles_j) list[atomicAdd(count, 1)] = x * tiles_j + y;
            }
    }
}

// DF_CPT cells per thread of a DF_TILE x DF_TILE tile (DF_TILE^2 / DF_CPT threads), the move set a template parameter: the 8 (King) or 5
// (Pawn) predecessors are compile-time offsets into the shared tile.  Only the neighbours that read an improved border
// cell in their halo are flagged for the next launch.  History on the 4096 x 4096 King field:
//   16.1 ms  one CTA per map tile, move table in a rolled loop, a host look per launch
//   14.4     frontier list, launches queued 16 at a time          13.3  constant offsets
//    9.2     directional flags                                      7.8  two cells per thread, no memset between launches
//   (tools/variants_df.sh: 1 / 2 / 4 cells per thread 9.2 / 7.8 / 10.0 ms; tried and slower: one WARP per tile with
//    Gauss-Seidel sweeps, 12.1 ms -- what a launch costs is the latency of the slowest tile visit)
// ncu on that version (profiles/r2v_k6_*): a launch is ISSUE-bound, 74 % issue-active, FP64 pipe 26 %, and two thirds of
// the instructions were fmin's NaN handling (DSETP.MIN + fix-up + moves, 7 per minimum).  Costs are never NaN, so:
//    6.1 ms  minimum as compare + two selects                       5.7  a thread's two cells vertically ADJACENT (10 loads
//                                                                        of the shared 3-column neighbourhood instead of 16)
//    5.0     the compare on the FP64 pipe (DSETP.LT, 3 instructions) instead of a 64-bit integer compare (4)
//    4.6     a warp skips an iteration when no row it can see changed (the wave crosses a tile as a band)
//    4.26    minimum of the predecessors FIRST, one addition per weight (exact: fl(x + w) is monotone in x)
// (3 CTAs per SM at 35 registers: 4.79; 4 cells per thread: 4.71; first host look after (tiles_i + tiles_j) / 2 launches.)
//    4.03    16 x 16 tiles of 128 threads (310 launches of 13 us, four times the frontier list)
//    3.47    launches chained by programmatic dependent launch (DF_PDL)
#ifndef DF_CPT
#define DF_CPT 2
#endif
#ifndef DF_ADJ
#define DF_ADJ 1                       // 1: a thread's cells are vertically ADJACENT (rows a0 .. a0 + DF_CPT - 1): the three-column
#endif                                 // neighbourhood of the cells overlaps, 5 + 5 (King) loads for 2 cells instead of 16;
                                       // 0: rows DF_TILE / DF_CPT apart (round-2 first version)
#define DF_RELAX_THREADS (DF_TILE * DF_TILE / DF_CPT)
#if DF_ADJ
#define DF_ROW0(tid) (1 + DF_CPT * ((tid) / DF_TILE))
#define DF_ROW(a0, q) ((a0) + (q))
#else
#define DF_ROW0(tid) (1 + ((tid) / DF_TILE))
#define DF_ROW(a0, q) ((a0) + (q) * (DF_TILE / DF_CPT))
#endif
// min of two costs.  Costs are +0, positive or +inf, never NaN: their bit patterns order like integers, so the minimum is
// an integer compare + two selects (4 ALU instructions) instead of fmin's DSETP.MIN + NaN fix-up + moves (7, one of
// them on the FP64 pipe).  The relaxation kernel is issue-bound (ncu: issue-active 74 %), instructions are time.
#ifndef DF_MIN_MODE
#define DF_MIN_MODE 1                  // 0: integer compare (4 ALU instructions); 1: DSETP.LT + two selects (3, one on the FP64 pipe);
                                       // 2: half each.  Measured (4096 x 4096 King, row skip on): 4.93 / 4.57 / 4.81 ms
#endif
__device__ __forceinline__ double df_min_i(double a, double b) {
    const long long x = __double_as_longlong(a), y = __double_as_longlong(b);
    return __longlong_as_double(y < x ? y : x);
}
__device__ __forceinline__ double df_min_d(double a, double b) { return (b < a) ? b : a; }
// df_min serves the straight moves, df_min2 the diagonal ones (mode 2: integer / FP64 compare, half each)
__device__ __forceinline__ double df_min(double a, double b) { return DF_MIN_MODE == 1 ? df_min_d(a, b) : df_min_i(a, b); }
__device__ __forceinline__ double df_min2(double a, double b) { return DF_MIN_MODE == 0 ? df_min_i(a, b) : df_min_d(a, b); }
__device__ __forceinline__ bool df_less(double a, double b) { return __double_as_longlong(a) < __double_as_longlong(b); }
#ifndef DF_MIN_CTAS
#define DF_MIN_CTAS 2                    // launch bound (with 32 x 32 tiles: 64 registers = two 512-thread CTAs per SM, 66 would leave one)
#endif
#ifndef DF_PDL
#define DF_PDL 1                        // relaxation launches with programmatic stream serialization (griddepcontrol)
#endif
#ifndef DF_MIN_TREE
#define DF_MIN_TREE 1                   // one addition per weight: min of the predecessors first (exact, see the loop)
#endif
#ifndef DF_ROWSKIP
#define DF_ROWSKIP 1                    // a warp sits an iteration out when no row its cells can see changed in the previous one
#endif
template <bool KING>
__global__ void __launch_bounds__(DF_RELAX_THREADS, DF_MIN_CTAS)
k_df_relax(const uint8_t* __restrict__ occ, int W, int H, int gi, int gj, double* __restrict__ D,
           const int* __restrict__ list_cur, const int* __restrict__ count_cur, int* __restrict__ list_next,
           int* __restrict__ count_next, int* __restrict__ mark_next, int* __restrict__ mark_cur, int* __restrict__ count_clear,
           int tiles_i, int tiles_j, int* __restrict__ any_next) {
    // the frontier is a LIST of tiles walked by a small grid with a stride (instead of one CTA per tile of the map, 16 k
    // CTAs of which a few hundred have work).  No memset between launches: a tile clears its own mark when it is
    // processed, and the counter the launch after next appends to is zeroed here (three counters rotate).
#if DF_PDL
    // programmatic dependent launch: this grid was scheduled while the previous relaxation launch was still running;
    // nothing that launch wrote is read before it has completed, and the NEXT launch may be scheduled from here on
    // (its CTAs wait at the same point), so the ~2 us between two dependent launches overlap with the work
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    if (blockIdx.x == 0 && threadIdx.x == 0) *count_clear = 0;
    __shared__ double sd[DF_TILE + 2][DF_TILE + 3];
    __shared__ unsigned char so[DF_TILE + 2][DF_TILE + 2];
    __shared__ int s_border_changed;
#if DF_ROWSKIP
    __shared__ int s_rowchg[2][DF_TILE + 2];            // rows that changed in the previous / this iteration
#endif
    const double W1 = 1.0, W2 = 1.4142135623730951;       // hypot(1, 0), hypot(1, 1) as the host's libm gives them
    const int n_cur = *count_cur;
    const int tid = threadIdx.x;
    const int a0 = DF_ROW0(tid), b = 1 + (tid % DF_TILE);      // cell q of this thread: row DF_ROW(a0, q), column b
    for (int li = blockIdx.x; li < n_cur; li += gridDim.x) {
        const int tile = list_cur[li];
        const int ti = tile / tiles_j, tj = tile % tiles_j;
        const int i0 = ti * DF_TILE - 1, j0 = tj * DF_TILE - 1;
        if (tid == 0) { s_border_changed = 0; mark_cur[tile] = 0; }
#if DF_ROWSKIP
        if (tid < DF_TILE + 2) { s_rowchg[0][tid] = 1; s_rowchg[1][tid] = 0; }      // first iteration: every row counts as changed
        int it = 0;
#endif
        for (int k = tid; k < (DF_TILE + 2) * (DF_TILE + 2); k += DF_RELAX_THREADS) {
            const int x = k / (DF_TILE + 2), y = k % (DF_TILE + 2);
            const int i = i0 + x, j = j0 + y;
            const bool in = i >= 0 && j >= 0 && i < W && j < H;
            sd[x][y] = in ? D[(long long)i * H + j] : INFINITY;
            so[x][y] = in ? occ[(long long)i * H + j] : 1;
        }
        __syncthreads();
        double init[DF_CPT], cur[DF_CPT];
        bool enter[DF_CPT];
#pragma unroll
        for (int q = 0; q < DF_CPT; ++q) {
            const int a = DF_ROW(a0, q);
            init[q] = cur[q] = sd[a][b];
            // occupied cells are never entered (the goal cell itself is exempt: it starts closed at 0)
            enter[q] = !so[a][b] && i0 + a < W && j0 + b < H;
        }
        while (true) {
            double best[DF_CPT];
#if DF_ROWSKIP && DF_ADJ
            // a cell can only change if a cell of rows a0 - 1 .. a0 + DF_CPT changed in the previous iteration: the warp
            // (one group of DF_CPT rows) sits the iteration out otherwise (the wave crosses a tile as a band)
            int seen = 0;
#pragma unroll
            for (int r = 0; r < DF_CPT + 2; ++r) seen |= s_rowchg[it & 1][a0 - 1 + r];
            if (!seen) {
#pragma unroll
                for (int q = 0; q < DF_CPT; ++q) best[q] = cur[q];
            } else {
#endif
#if DF_ADJ
            // columns b - 1, b, b + 1 of rows a0 - 1 .. a0 + DF_CPT; the thread's own cells (column b) are in registers,
            // and a cell sees the value its upper neighbour in the same thread was just given (any relaxation order
            // reaches the same fixed point)
            double L[DF_CPT + 2], M[DF_CPT + 2], R[DF_CPT + 2];
#pragma unroll
            for (int r = 0; r < DF_CPT + 2; ++r) {
                L[r] = sd[a0 - 1 + r][b - 1];
                R[r] = KING ? sd[a0 - 1 + r][b + 1] : 0.0;
            }
            M[0] = sd[a0 - 1][b]; M[DF_CPT + 1] = sd[a0 + DF_CPT][b];
#pragma unroll
            for (int q = 0; q < DF_CPT; ++q) M[q + 1] = cur[q];
#if DF_MIN_TREE
            // fl(x + w) is monotone in x, so min(fl(x + w), fl(y + w)) == fl(min(x, y) + w) bit for bit: ONE addition per
            // weight and cell (the minimum of the straight predecessors + 1, of the diagonal ones + sqrt 2) instead of
            // eight, and the minimum of a row's left and right cell serves the cell between them (straight) and the two
            // cells above / below it (diagonal).  Pawn moves only come from column b - 1.
            double hr[DF_CPT + 2];
#pragma unroll
            for (int r = 0; r < DF_CPT + 2; ++r) hr[r] = KING ? df_min(L[r], R[r]) : L[r];
#pragma unroll
            for (int q = 0; q < DF_CPT; ++q) {
                const int r = q + 1;
                double v = cur[q];
                if (enter[q]) {
                    const double m1 = df_min(df_min(M[r - 1], M[r + 1]), hr[r]);
                    const double m2 = df_min(hr[r - 1], hr[r + 1]);
                    v = df_min(v, df_min(__dadd_rn(m1, W1), __dadd_rn(m2, W2)));
                }
                best[q] = v;
                M[r] = v;
            }
#else
#pragma unroll
            for (int q = 0; q < DF_CPT; ++q) {
                const int r = q + 1;
                double v = cur[q];
                if (enter[q]) {                                     // predecessor of move (di, dj) is (a - di, b - dj)
                    v = df_min(v, __dadd_rn(M[r + 1], W1));                             // (-1,  0)
                    v = df_min(v, __dadd_rn(L[r], W1));                                 // ( 0,  1)
                    v = df_min2(v, __dadd_rn(L[r + 1], W2));                             // (-1,  1)
                    v = df_min2(v, __dadd_rn(L[r - 1], W2));                             // ( 1,  1)
                    v = df_min(v, __dadd_rn(M[r - 1], W1));                             // ( 1,  0)
                    if (KING) {
                        v = df_min2(v, __dadd_rn(R[r - 1], W2));                         // ( 1, -1)
                        v = df_min(v, __dadd_rn(R[r], W1));                             // ( 0, -1)
                        v = df_min2(v, __dadd_rn(R[r + 1], W2));                         // (-1, -1)
                    }
                }
                best[q] = v;
                M[r] = v;
            }
#endif
#if DF_ROWSKIP
            }
#endif
#else
#pragma unroll
            for (int q = 0; q < DF_CPT; ++q) {
                const int a = DF_ROW(a0, q);
                double v = cur[q];
                if (enter[q]) {                                     // predecessor of move (di, dj) is (a - di, b - dj)
                    v = df_min(v, __dadd_rn(sd[a + 1][b], W1));                           // (-1,  0)
                    v = df_min(v, __dadd_rn(sd[a][b - 1], W1));                           // ( 0,  1)
                    v = df_min(v, __dadd_rn(sd[a + 1][b - 1], W2));                       // (-1,  1)
                    v = df_min(v, __dadd_rn(sd[a - 1][b - 1], W2));                       // ( 1,  1)
                    v = df_min(v, __dadd_rn(sd[a - 1][b], W1));                           // ( 1,  0)
                    if (KING) {
                        v = df_min(v, __dadd_rn(sd[a - 1][b + 1], W2));                   // ( 1, -1)
                        v = df_min(v, __dadd_rn(sd[a][b + 1], W1));                       // ( 0, -1)
                        v = df_min(v, __dadd_rn(sd[a + 1][b + 1], W2));                   // (-1, -1)
                    }
                }
                best[q] = v;
            }
#endif
            __syncthreads();
            bool ch = false;
#pragma unroll
            for (int q = 0; q < DF_CPT; ++q)
                if (df_less(best[q], cur[q])) {
                    sd[DF_ROW(a0, q)][b] = best[q]; cur[q] = best[q]; ch = true;
#if DF_ROWSKIP && DF_ADJ
                    s_rowchg[(it + 1) & 1][DF_ROW(a0, q)] = 1;
#endif
                }
#if DF_ROWSKIP && DF_ADJ
            if (tid < DF_TILE + 2) s_rowchg[it & 1][tid] = 0;      // read before the barrier above, written again in iteration it + 2
            ++it;
#endif
            if (!__syncthreads_or(ch)) break;
        }
#pragma unroll
        for (int q = 0; q < DF_CPT; ++q) {
            const int a = DF_ROW(a0, q);
            const int i = i0 + a, j = j0 + b;
            if (cur[q] < init[q] && i < W && j < H) {
                D[(long long)i * H + j] = cur[q];
                // which neighbouring tiles read this cell in their halo: bit 3 * (x + 1) + (y + 1) for neighbour (x, y)
                // (a side cell is read by that side's neighbour only, a corner cell also by the diagonal neighbour)
                const int x = (a == 1) ? -1 : (a == DF_TILE ? 1 : 0), y = (b == 1) ? -1 : (b == DF_TILE ? 1 : 0);
                if (x || y) {
                    unsigned m = 0;
                    if (x) m |= 1u << (3 * (x + 1) + 1);
                    if (y) m |= 1u << (3 + (y + 1));
                    if (x && y) m |= 1u << (3 * (x + 1) + (y + 1));
                    atomicOr(&s_border_changed, (int)m);
                }
            }
        }
        __syncthreads();
        if (tid < 9 && tid != 4 && (s_border_changed >> tid) & 1) {
            const int tx = ti + tid / 3 - 1, ty = tj + tid % 3 - 1;
            if (tx >= 0 && ty >= 0 && tx < tiles_i && ty < tiles_j && atomicExch(&mark_next[tx * tiles_j + ty], 1) == 0)
                list_next[atomicAdd(count_next, 1)] = tx * tiles_j + ty;
            *any_next = 1;
        }
        __syncthreads();                                   // the shared tile is reused by the next list entry
    }
}

// ---- the reference's index wrap-around (a_star_utils.py:49-64) --------------------------------------------
// Validity is `abs(index) < dim` and obstacles[i][j] is read with Python's negative indexing, so the reference
// really searches the EXTENDED grid of nodes (i, j), -W < i < W, -H < j < H, whose occupancy is the map's at
// (i mod W, j mod H); every closed node writes holonomicCost[i][j] (again with negative indexing) in closing order,
// so a map cell ends up with the cost of its LAST-closed alias.  Nodes close in the order of their heap entries
// (priority at push time, index tuple) -- there is no decrease-key (:131), so the priority of a node is the cost
// offered by its FIRST-closed neighbour, not its final cost.  All of it has a relaxation-order-free form:
//   C = least fixed point of the cost relaxation on the extended grid            (k_df_relax, as for closed maps)
//   P(v) = C(u*) + w(u* -> v),  u* = the neighbour of v with the smallest (P(u), i_u, j_u)      (fixed point, k_df_prio)
//   out[cell] = C(alias of the cell with the largest (P, i, j))                                   (k_df_fold)
// oracle/distance_field.py (pinned bit-for-bit on the reference) agrees on every tested grid
// (tests/test_grid_gpu.py::test_wrap_around_quirk_matches_reference).
__global__ void k_df_ext_occ(const uint8_t* __restrict__ occ, int W, int H, uint8_t* __restrict__ eocc) {
    const int EW = 2 * W - 1, EH = 2 * H - 1;
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)EW * EH) return;
    const int a = (int)(idx / EH), b = (int)(idx % EH);
    const int i = a - (W - 1), j = b - (H - 1);
    eocc[idx] = occ[(long long)(i < 0 ? i + W : i) * H + (j < 0 ? j + H : j)];
}

__global__ void k_df_prio(const double* __restrict__ C, const double* __restrict__ Pin, double* __restrict__ Pout,
                          int EW, int EH, DfMoves mv, int ga, int gb, int* __restrict__ changed) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)EW * EH) return;
    const int a = (int)(idx / EH), b = (int)(idx % EH);
    double np_ = INFINITY;
    if (a == ga && b == gb) np_ = 0.0;
    else if (C[idx] < INFINITY) {
        double bp = INFINITY; int ba = 0x7fffffff, bb = 0x7fffffff;
        for (int m = 0; m < mv.n; ++m) {
            const int ua = a - mv.di[m], ub = b - mv.dj[m];
            if (ua < 0 || ub < 0 || ua >= EW || ub >= EH) continue;
            const long long u = (long long)ua * EH + ub;
            const double cu = C[u];
            if (!(cu < INFINITY)) continue;
            const double pu = Pin[u];
            if (pu < bp || (pu == bp && (ua < ba || (ua == ba && ub < bb)))) { bp = pu; ba = ua; bb = ub; np_ = __dadd_rn(cu, mv.w[m]); }
        }
    }
    if (np_ != Pin[idx]) *changed = 1;
    Pout[idx] = np_;
}

__global__ void k_df_fold(const double* __restrict__ C, const double* __restrict__ P, int W, int H, double* __restrict__ out) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)W * H) return;
    const int ci = (int)(idx / H), cj = (int)(idx % H);
    const int EH = 2 * H - 1;
    double best_c = INFINITY, best_p = -INFINITY;
    int best_i = -0x7fffffff, best_j = -0x7fffffff;
    for (int ai = 0; ai < 2; ++ai) {
        if (ai == 1 && ci < 1) continue;
        const int i = ai == 0 ? ci : ci - W;
        for (int aj = 0; aj < 2; ++aj) {
            if (aj == 1 && cj < 1) continue;
            const int j = aj == 0 ? cj : cj - H;
            const long long v = (long long)(i + W - 1) * EH + (j + H - 1);
            const double c = C[v];
            if (!(c < INFINITY)) continue;
            const double p = P[v];
            if (p > best_p || (p == best_p && (i > best_i || (i == best_i && j > best_j)))) { best_p = p; best_i = i; best_j = j; best_c = c; }
        }
    }
    out[idx] = best_c;
}

// Tiled wavefront to the fixed point on a w x h grid (cells outside count as occupied).  *open_border is set when
// a border cell is free (only when `open_border` is given; nothing is relaxed then).
static int df_solve(const uint8_t* d_occ, int w, int h, int gi, int gj, const DfMoves& mv, double* d_out,
                    cudaStream_t st, int* open_border, int* sweeps_out) {
    const int tiles_i = (w + DF_TILE - 1) / DF_TILE, tiles_j = (h + DF_TILE - 1) / DF_TILE;
    const int n_tiles = tiles_i * tiles_j;
    int* ws = nullptr;                       // list[2][n_tiles] | mark[2][n_tiles] | count[2] | flags[1 + DF_BATCH]
    // The wavefront needs one launch per tile it crosses (~150 for a 4096 x 4096 map), so the launches are queued
    // DF_BATCH at a time and the host looks at the "frontier not empty" flags once per batch (one stream
    // synchronisation per launch before); a launch with an empty frontier list costs a few microseconds.
    // Launches are queued DF_BATCH at a time with one host look per batch (launches behind the first empty frontier do
    // nothing and cost ~2 us each).  The wavefront needs about one launch per tile it crosses, so the first look comes
    // after (tiles_i + tiles_j) / 2 launches, the later ones every 32: 2 host looks for the 4096 x 4096 field (157
    // launches) instead of 10 -- every look is a stream synchronisation, i.e. an idle GPU for as long as the host
    // thread takes to come back.
    enum { DF_BATCH_MAX = 128, DF_BATCH_NEXT = 32 };
    int DF_BATCH = (tiles_i + tiles_j) / 2 < 16 ? 16 : ((tiles_i + tiles_j) / 2 > DF_BATCH_MAX ? DF_BATCH_MAX : (tiles_i + tiles_j) / 2);
    const size_t ws_ints = 4 * (size_t)n_tiles + 4 + 1 + DF_BATCH_MAX;
    HL_CUDA_OK(cudaMalloc(&ws, ws_ints * sizeof(int)));
    int* list[2] = {ws, ws + n_tiles};
    int* mark[2] = {ws + 2 * (size_t)n_tiles, ws + 3 * (size_t)n_tiles};
    int* count = ws + 4 * (size_t)n_tiles;
    int* flags = count + 4;                  // count[0..2]: three rotating frontier counters
    int rc = 0, sweeps = 0;
    int sm = 148;
    { int dev_id = 0; cudaGetDevice(&dev_id); cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev_id); }
    // CTAs per SM in the grid: MORE than are resident.  The frontier list is walked with a grid
    // stride, so with a grid of exactly the resident CTAs a CTA that drew a slow tile keeps its second tile waiting; with
    // twice as many the block scheduler hands the next list entry to whichever SM frees a slot first
    // (4096 x 4096 King, 32 x 32 tiles of 512 threads: 2 per SM 4.87 ms, 4 per SM 4.26 ms; 16 x 16 tiles of 128 threads:
    // 16 or 32 per SM 4.03 ms).
#ifndef DF_GRID_PER_SM
#define DF_GRID_PER_SM 16
#endif
    const int per_sm = DF_GRID_PER_SM;
    const int grid = n_tiles < sm * per_sm ? n_tiles : sm * per_sm;
    do {
        if (cudaMemsetAsync(ws, 0, ws_ints * sizeof(int), st) != cudaSuccess) { rc = 1; break; }
        long long cells = (long long)w * h;
        k_df_init<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(d_occ, w, h, gi, gj, d_out, list[0], count, tiles_i, tiles_j, flags);
        int host_flags[1 + DF_BATCH_MAX] = {0};
        if (cudaMemcpyAsync(host_flags, flags, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = 1; break; }
        if (open_border) { *open_border = host_flags[0]; if (host_flags[0]) break; }
        int cur = 0;                            // launch number: lists / marks alternate (cur & 1), counters rotate (cur % 3)
        const int max_sweeps = 8 * (tiles_i + tiles_j) * DF_TILE + 64;
        bool converged = false;
        while (!converged && sweeps < max_sweeps) {
            cudaMemsetAsync(flags + 1, 0, DF_BATCH * sizeof(int), st);
            for (int k = 0; k < DF_BATCH; ++k) {
                const int lc = cur & 1, ln = lc ^ 1;
#ifdef DF_MEMSET
                cudaMemsetAsync(mark[ln], 0, sizeof(int) * (size_t)n_tiles, st);
#endif
                int* c_cur = count + cur % 3; int* c_nxt = count + (cur + 1) % 3; int* c_clr = count + (cur + 2) % 3;
#if DF_PDL
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(DF_RELAX_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                const uint8_t* a_occ = d_occ; double* a_out = d_out;
                const int* a_lc = list[lc]; const int* a_cc = c_cur;
                int* a_fl = flags + 1 + k;
                cudaError_t le;
                if (mv.n == 8)
                    le = cudaLaunchKernelEx(&cfg, k_df_relax<true>, a_occ, w, h, gi, gj, a_out, a_lc, a_cc, list[ln], c_nxt, mark[ln],
                                            mark[lc], c_clr, tiles_i, tiles_j, a_fl);
                else
                    le = cudaLaunchKernelEx(&cfg, k_df_relax<false>, a_occ, w, h, gi, gj, a_out, a_lc, a_cc, list[ln], c_nxt, mark[ln],
                                            mark[lc], c_clr, tiles_i, tiles_j, a_fl);
                if (le != cudaSuccess) { rc = 1; break; }       // a launch that did not happen would read as "frontier empty"
#else
                if (mv.n == 8)
                    k_df_relax<true><<<grid, DF_RELAX_THREADS, 0, st>>>(d_occ, w, h, gi, gj, d_out, list[lc], c_cur, list[ln], c_nxt,
                                                                        mark[ln], mark[lc], c_clr, tiles_i, tiles_j, flags + 1 + k);
                else
                    k_df_relax<false><<<grid, DF_RELAX_THREADS, 0, st>>>(d_occ, w, h, gi, gj, d_out, list[lc], c_cur, list[ln], c_nxt,
                                                                         mark[ln], mark[lc], c_clr, tiles_i, tiles_j, flags + 1 + k);
                if (cudaGetLastError() != cudaSuccess) { rc = 1; break; }
#endif
                ++cur;
            }
            if (rc) { cudaStreamSynchronize(st); break; }
            if (cudaMemcpyAsync(host_flags + 1, flags + 1, DF_BATCH * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = 1; break; }
            for (int k = 0; k < DF_BATCH && !converged; ++k) {   // launches after the first empty frontier did nothing
                ++sweeps;
                if (!host_flags[1 + k]) converged = true;
            }
            DF_BATCH = DF_BATCH_NEXT;
        }
        if (rc == 0 && !converged) { hl_set_error("hl_distance_field: no convergence after %d launches", sweeps); rc = 2; }
    } while (0);
    if (rc == 1) hl_set_error("hl_distance_field: CUDA error: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(ws);
    if (sweeps_out) *sweeps_out += sweeps;
    return rc;
}

extern "C" int hl_distance_field(hl_ctx* ctx, const uint8_t* d_occ, int32_t w, int32_t h, int32_t gi,
                                 int32_t gj, int32_t motion_type, double* d_out, int32_t* h_sweeps,
                                 void* stream) {
    if (!ctx || !d_occ || !d_out || w < 1 || h < 1) { hl_set_error("hl_distance_field: bad arguments"); return 1; }
    if (gi < 0 || gj < 0 || gi >= w || gj >= h) { hl_set_error("hl_distance_field: goal (%d,%d) outside the %dx%d grid", gi, gj, w, h); return 1; }
    if (hl_enter(ctx, nullptr, d_out, "hl_distance_field")) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    DfMoves mv;
    memset(&mv, 0, sizeof(mv));
    if (motion_type == 0) {                 // a_star_utils.py:8-21
        const int k[8][2] = {{-1, 0}, {-1, 1}, {0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}};
        mv.n = 8;
        for (int m = 0; m < 8; ++m) { mv.di[m] = k[m][0]; mv.dj[m] = k[m][1]; mv.w[m] = hypot((double)k[m][0], (double)k[m][1]); }
    } else if (motion_type == 1) {          // :24-34
        const int k[5][2] = {{-1, 0}, {0, 1}, {-1, 1}, {1, 1}, {1, 0}};
        mv.n = 5;
        for (int m = 0; m < 5; ++m) { mv.di[m] = k[m][0]; mv.dj[m] = k[m][1]; mv.w[m] = hypot((double)k[m][0], (double)k[m][1]); }
    } else { hl_set_error("hl_distance_field: motion_type must be 0 (King) or 1 (Pawn)"); return 1; }
    int sweeps = 0;
    // fast path: every border cell occupied and the goal strictly inside -> no alias is reachable, the map itself is
    // the whole search space
    if (w >= 3 && h >= 3 && gi > 0 && gj > 0 && gi < w - 1 && gj < h - 1) {
        int open_border = 0;
        const int rc = df_solve(d_occ, w, h, gi, gj, mv, d_out, st, &open_border, &sweeps);
        if (rc) return 1;
        if (!open_border) { if (h_sweeps) *h_sweeps = sweeps; return 0; }
    }
    // general path: the extended grid of the reference's index wrap-around
    const int EW = 2 * w - 1, EH = 2 * h - 1;
    const long long ecells = (long long)EW * EH;
    uint8_t* eocc = nullptr;
    double* buf = nullptr;                   // C | P0 | P1
    int* changed = nullptr;
    if (cudaMalloc(&eocc, (size_t)ecells) != cudaSuccess || cudaMalloc(&buf, 3 * sizeof(double) * (size_t)ecells) != cudaSuccess ||
        cudaMalloc(&changed, sizeof(int)) != cudaSuccess) {
        if (eocc) cudaFree(eocc);
        if (buf) cudaFree(buf);
        hl_set_error("hl_distance_field: cudaMalloc failed for the %dx%d extended grid", EW, EH);
        return 1;
    }
    double* C = buf; double* P0 = buf + ecells; double* P1 = buf + 2 * ecells;
    const unsigned eblocks = (unsigned)((ecells + 255) / 256);
    int rc = 0;
    do {
        k_df_ext_occ<<<eblocks, 256, 0, st>>>(d_occ, w, h, eocc);
        if (df_solve(eocc, EW, EH, gi + w - 1, gj + h - 1, mv, C, st, nullptr, &sweeps)) { rc = 1; break; }
        if (cudaMemcpyAsync(P0, C, sizeof(double) * (size_t)ecells, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { rc = 2; break; }
        int it = 0, h_changed = 1;
        while (h_changed && it < 4 * (EW + EH)) {
            cudaMemsetAsync(changed, 0, sizeof(int), st);
            k_df_prio<<<eblocks, 256, 0, st>>>(C, P0, P1, EW, EH, mv, gi + w - 1, gj + h - 1, changed);
            if (cudaMemcpyAsync(&h_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = 2; break; }
            double* t = P0; P0 = P1; P1 = t;
            ++it; ++sweeps;
        }
        if (rc) break;
        if (h_changed) { hl_set_error("hl_distance_field: closing-order priorities did not converge"); rc = 1; break; }
        k_df_fold<<<(unsigned)(((long long)w * h + 255) / 256), 256, 0, st>>>(C, P0, w, h, d_out);
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = 2; break; }
    } while (0);
    if (rc == 2) hl_set_error("hl_distance_field: CUDA error: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(eocc); cudaFree(buf); cudaFree(changed);
    if (h_sweeps) *h_sweeps = sweeps;
    return rc ? 1 : 0;
}

// ------------------------------------------------------------------------- K7
__global__ void k_grid_pack(const uint8_t* __restrict__ occ, long long cells, uint32_t* __restrict__ bits) {
    long long wd = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long n_words = (cells + 31) / 32;
    if (wd >= n_words) return;
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
        long long c = wd * 32 + b;
        if (c < cells && occ[c]) v |= 1u << b;
    }
    bits[wd] = v;
}

extern "C" int hl_grid_pack(hl_ctx* ctx, const uint8_t* d_occ, int32_t w, int32_t h, uint32_t* d_bits, void* stream) {
    if (!ctx || !d_occ || !d_bits || w <= 0 || h <= 0) { hl_set_error("hl_grid_pack: bad arguments"); return 1; }
    if (hl_enter(ctx, nullptr, d_bits, "hl_grid_pack")) return 1;
    long long cells = (long long)w * h, n_words = (cells + 31) / 32;
    k_grid_pack<<<(unsigned)((n_words + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_occ, cells, d_bits);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}

// any occupied bit among cells (i, j0..j1) of the bit-packed grid
__device__ __forceinline__ bool row_any(const uint32_t* __restrict__ bits, long long base, int j0, int j1) {
    long long a = base + j0, b = base + j1;
    long long wa = a >> 5, wb = b >> 5;
    for (long long wd = wa; wd <= wb; ++wd) {
        uint32_t m = 0xFFFFFFFFu;
        if (wd == wa) m &= 0xFFFFFFFFu << (a & 31);
        if (wd == wb) m &= 0xFFFFFFFFu >> (31 - (b & 31));
        if (__ldg(bits + wd) & m) return true;
    }
    return false;
}

__global__ void __launch_bounds__(256)
k_grid_footprint(const uint32_t* __restrict__ bits, int W, int H, double res, const double* __restrict__ poses,
                 long long n, double x0, double x1, double y0, double y1, uint8_t* __restrict__ out) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const double px = poses[3 * idx], py = poses[3 * idx + 1], yaw = poses[3 * idx + 2];
    const double c = cos(yaw), s = sin(yaw);
    const double lx[4] = {x0, x0, x1, x1}, ly[4] = {y1, y0, y0, y1};
    double cx[4], cy[4];
    double xmin = INFINITY, xmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cx[k] = xadd(xsub(xmul(c, lx[k]), xmul(s, ly[k])), px);
        cy[k] = xadd(xadd(xmul(s, lx[k]), xmul(c, ly[k])), py);
        xmin = fmin(xmin, cx[k]); xmax = fmax(xmax, cx[k]);
    }
    int ia = (int)floor(xmin / res), ib = (int)floor(xmax / res);
    // a rectangle that touches x = i*res from above also meets cell i-1 (closed sets)
    if ((double)ia * res == xmin) ia -= 1;
    ia = max(ia, 0); ib = min(ib, W - 1);
    bool hit = false;
    for (int i = ia; i <= ib && !hit; ++i) {
        const double a = (double)i * res, b = (double)(i + 1) * res;
        double ymin = INFINITY, ymax = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (cx[k] >= a && cx[k] <= b) { ymin = fmin(ymin, cy[k]); ymax = fmax(ymax, cy[k]); }
            const int k2 = (k + 1) & 3;
            const double ex = cx[k2] - cx[k];
            if (ex != 0.0) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double X = e ? b : a;
                    if ((cx[k] - X) * (cx[k2] - X) < 0.0) {
                        double y = cy[k] + (X - cx[k]) * (cy[k2] - cy[k]) / ex;
                        ymin = fmin(ymin, y); ymax = fmax(ymax, y);
                    }
                }
            }
        }
        if (ymin > ymax) continue;
        int ja = (int)floor(ymin / res), jb = (int)floor(ymax / res);
        if ((double)ja * res == ymin) ja -= 1;
        ja = max(ja, 0); jb = min(jb, H - 1);
        if (ja <= jb) hit = row_any(bits, (long long)i * H, ja, jb);
    }
    out[idx] = hit ? 1 : 0;
}

extern "C" int hl_grid_footprint_check(hl_ctx* ctx, const uint32_t* d_occ_bits, int32_t w, int32_t h,
                                       double res, const double* d_poses, int64_t n,
                                       const double body_ext[4], uint8_t* d_out, void* stream) {
    if (!ctx || !d_occ_bits || !d_poses || !d_out || !body_ext || n < 0 || w <= 0 || h <= 0 || !(res > 0)) {
        hl_set_error("hl_grid_footprint_check: bad arguments"); return 1;
    }
    if (n == 0) return 0;
    if (hl_enter(ctx, nullptr, d_out, "hl_grid_footprint_check")) return 1;
    k_grid_footprint<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_occ_bits, w, h, res, d_poses, (long long)n, body_ext[0], body_ext[1], body_ext[2], body_ext[3], d_out);
    HL_CUDA_OK(cudaGetLastError());
    return 0;
}
