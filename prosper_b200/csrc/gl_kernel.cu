// Posterior kernel of the Gaussian-linear ET models (BSC, TSC, DSC): one warp per datapoint.
//
// Input per datapoint n: the score row YW[n,:] = y_n . W (from the score GEMM), ||y_n||^2,
// and per iteration G = W^T W.  With these, for a latent state s supported on the H'
// candidates c_1..c_H' (values v_j),
//     ||y - sum_j v_j W_cj||^2 = ||y||^2 - 2 sum_j v_j YW[c_j] + sum_jk v_j v_k G[c_j,c_k]
// so the whole truncated state space (<= gamma non-zeros of H') is evaluated from H'
// gathered scalars and an H'xH' gathered block held in shared memory -- no D-dimensional
// work per state (the reference does an S x H' x D product per datapoint, bsc_et.py:180-184).
//
// Phases (each cites what it replaces):
//   1 top-H' preselection, scores in registers   bsc_et.py:110-112 / tsc_et.py:198-210 / dsc_et.py:398-408
//   2 gather YW[cand], G[cand,cand]
//   3 log-joint of every column, running max     bsc_et.py:168-190 / tsc_et.py:337-355 / dsc_et.py:558-584
//   4 exp / sum (log-sum-exp), sigma statistic   bsc_et.py:271-272,362,395-415
//   5 pair sums over the state space from a shared-memory gather table (16-bit state ids)
//   6 first/second posterior moments, <s> row for the statistics GEMM, Wq scatter   bsc_et.py:349-366
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "gl_kernel.cuh"

namespace pet {

constexpr int GL_MAX_GROUPS = 8;           // datapoints in flight per CTA of the state kernel (2 warps each)
constexpr int ROW_WARPS = 8;               // warps per CTA of the row kernel
constexpr double GL_EXP_CUTOFF = -100.0;   // exp(x) for x below this contributes < 4e-44 relative

static __host__ __device__ inline int r2(int x) { return (x + 1) & ~1; }

constexpr int GS = PET_MAXHP + 1;          // stride of the gathered Gram block; row/col Hp is all zero

constexpr int GRP_LANES = PET_GRP_LANES;   // four warps work on one datapoint in the state kernel
constexpr int GRP_WARPS = GRP_LANES / 32;
constexpr int GL_CHUNK = 8;                // entries per gather chunk (fixed: the inner loop is fully unrolled)

struct SmemLayout {
    int shared_doubles;   // state records + gather table + chunk table
    int off_ids, off_chunk;
    int per_dp;           // doubles per datapoint group
    int off_G, off_lin, off_mom, off_P, off_red, off_cand;
};

static __host__ __device__ inline SmemLayout smem_layout(const GLStatic &s) {
    SmemLayout L;
    const int nch = s.n_chunks > 0 ? s.n_chunks : 1;
    L.off_ids = r2(s.S);
    L.off_chunk = L.off_ids + nch * s.chunk_len * (GRP_LANES / 4);   // nch*CH*64 u16
    L.shared_doubles = L.off_chunk + nch * (GRP_LANES / 4);          // nch*64 u16
    L.off_G = r2(s.S + 1 + PET_MAXHP);                               // qbuf first: states, zero slot, singletons
    L.off_lin = L.off_G + r2(GS * GS);
    L.off_mom = L.off_lin + r2(GS);
    L.off_P = L.off_mom + r2(s.n_out + 1);
    L.off_red = L.off_P + r2(s.Hp * (s.n_cnt > 0 ? s.n_cnt : 1));
    L.off_cand = L.off_red + 16;
    L.per_dp = L.off_cand + PET_MAXHP;     // cand[16] + live[16] ints
    return L;
}

// bytes of the state kernel with `groups` datapoints in flight per CTA
size_t gl_smem_bytes(const GLStatic &s, int groups) {
    SmemLayout L = smem_layout(s);
    return (size_t(L.shared_doubles) + size_t(L.per_dp) * groups) * sizeof(double);
}

__constant__ double c_winv[8] = {1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8};

// log-prior of multi-state s.  Binary spaces are enumerated by size (camodels/__init__.py:30-33), so
// |s| follows from the index (size boundaries held in registers) and nothing is loaded; valued spaces
// read the per-iteration table.
template <int GMAX>
struct SizeBounds {
    int b[GMAX + 1];
    __device__ __forceinline__ void load(const GLStatic &st) {
#pragma unroll
        for (int g = 3; g <= GMAX; ++g) b[g] = st.size_start[g];
    }
    __device__ __forceinline__ int members(int s) const {
        int n = 2;
#pragma unroll
        for (int g = 3; g <= GMAX; ++g) n += (s >= b[g]) ? 1 : 0;
        return n;
    }
};

template <int GMAX, bool BINARY>
__device__ __forceinline__ double prior_of(const GLArgs &a, const SizeBounds<GMAX> &sb, int s) {
    if (BINARY) return a.it.lp[0] * double(sb.members(s));
    return a.state_prior[s];
}

__device__ __forceinline__ double combine(const GLIter &it, double prior, double q) {
    return it.anneal_prior ? it.beta * (prior + it.pre1 * q) : prior + it.beta * (it.pre1 * q);
}

// squared error of one multi-state record.
// BINARY: member bytes are positions; unused slots hold the dummy position Hp whose lin[] entry and
// Gram row/column are zero, so there are no multiplies, selects or branches:
//     q = yy + sum_m lin[p_m] + 2 sum_{m2<m} G[p_m][p_m2],   lin[p] = G[p][p] - 2 YW[c_p]
// valued states (TSC/DSC): byte = pos | vidx<<4, 0xFF unused; lin[] holds YW[c_p].
template <int GMAX, bool BINARY>
__device__ __forceinline__ double eval_state(unsigned long long rec, const GLStatic &st, const double *lin,
                                             const double *Gc, double yy) {
    if (BINARY) {
        int p[GMAX];
#pragma unroll
        for (int m = 0; m < GMAX; ++m) p[m] = int(unsigned(rec >> (8 * m)) & 0xFFu);
        double a1 = yy, a2 = 0.0;
#pragma unroll
        for (int m = 0; m < GMAX; ++m) {
            a1 += lin[p[m]];
            const double *g = Gc + p[m] * GS;
#pragma unroll
            for (int m2 = 0; m2 < m; ++m2) a2 += g[p[m2]];
        }
        return fma(2.0, a2, a1);
    } else {
        int pos[GMAX];
        double val[GMAX];
#pragma unroll
        for (int m = 0; m < GMAX; ++m) {
            unsigned b = unsigned(rec >> (8 * m)) & 0xFFu;
            bool ok = (b != 0xFFu);
            pos[m] = ok ? int(b & 15u) : 0;
            val[m] = ok ? st.vals[b >> 4] : 0.0;
        }
        double acc = yy;
#pragma unroll
        for (int m = 0; m < GMAX; ++m) {
            double l = fma(val[m], Gc[pos[m] * GS + pos[m]], -2.0 * lin[pos[m]]);
            double cross = 0.0;
#pragma unroll
            for (int m2 = 0; m2 < m; ++m2) cross = fma(val[m2], Gc[pos[m] * GS + pos[m2]], cross);
            acc = fma(val[m], fma(2.0, cross, l), acc);
        }
        return acc;
    }
}

// log-prior of every multi-state, once per iteration (it depends on pi only):
//   BSC pil_bar*|s| (bsc_et.py:164); TSC sum_j log p(s_j) over H' entries (tsc_et.py:316-326);
//   DSC state_abs . log pi with H - nnz zeros (dsc_et.py:525-527)
__global__ void state_prior_kernel(GLStatic st, GLIter it, double *out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.S) return;
    unsigned long long rec = st.states[s];
    const unsigned unused = st.binary ? unsigned(st.Hp) : 0xFFu;
    double pr = double(st.zbase) * it.lp0;
    for (int m = 0; m < PET_MAXG; ++m) {
        unsigned b = unsigned(rec >> (8 * m)) & 0xFFu;
        if (b != unused && b != 0xFFu) pr += it.lp[(b >> 4) & 15u] - it.lp0;
    }
    out[s] = pr;
}

int launch_state_prior(const GLStatic &st, const GLIter &it, double *out, cudaStream_t stream) {
    if (st.S <= 0) return PET_OK;
    state_prior_kernel<<<(st.S + 127) / 128, 128, 0, stream>>>(st, it, out);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// selection score of item i (a cause h, or for TSC a signed singleton (sign block, h))
__device__ __forceinline__ double sel_score(const GLArgs &a, const double *row, int i, double yy) {
    const GLStatic &st = a.st;
    switch (st.select_mode) {
        case SEL_BSC:
            return (a.wmu ? row[i] + a.wmu[i] : row[i]) * a.invn[i];
        case SEL_NEGDIST:
            return 2.0 * row[i] - a.wn2[i];
        case SEL_GIVEN:
            return -row[i];
        case SEL_TSC: {   // -1 block first, then +1 (tsc_et.py:54-65); the log-prior is the same for all
            int h = i % st.H;
            double sg = (i < st.H) ? -1.0 : 1.0;
            return a.it.pre1 * (yy + (a.wn2[h] - 2.0 * sg * row[h]));
        }
        default: {        // SEL_DSC: best valued singleton of this h
            double best = -INFINITY;
            for (int b = 0; b < st.n_blocks; ++b) {
                double v = st.block_val[b];
                double q = yy + v * (v * a.wn2[i] - 2.0 * row[i]);
                best = fmax(best, a.it.sel_prior[b] + a.it.pre1 * q);
            }
            return best;
        }
    }
}

__device__ __forceinline__ void warp_argmax(double &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if ((ov > v) || (ov == v && oi > i)) { v = ov; i = oi; }
    }
}

__device__ __forceinline__ void store_cand(const GLStatic &st, int *cand, int rnd, int item) {
    const bool tsc = (st.select_mode == SEL_TSC);
    const bool ascending = (st.select_mode == SEL_BSC) || tsc;
    cand[ascending ? (st.Hp - 1 - rnd) : rnd] = tsc ? (item % st.H) : item;
}

// Top-H' with the scores of this lane's items held in registers (items <= 32*HC).
// Order: value descending, ties -> larger item index first (same rule as the generic path).
template <int HC>
__device__ __noinline__ void select_regs(const GLArgs &a, const double *row, double yy, int items, int *cand) {
    const int lane = threadIdx.x & 31;
    double sc[HC];
#pragma unroll
    for (int k = 0; k < HC; ++k) {
        int i = k * 32 + lane;
        sc[k] = (i < items) ? sel_score(a, row, i, yy) : -INFINITY;
        if (sc[k] != sc[k]) sc[k] = -INFINITY;       // NaN scores never win
    }
    for (int rnd = 0; rnd < a.st.Hp; ++rnd) {
        double lm = -INFINITY;
        int lk = 0;
#pragma unroll
        for (int k = 0; k < HC; ++k)
            if (sc[k] >= lm) { lm = sc[k]; lk = k; }     // '>=': the larger item index wins ties
        double bv = lm;
        int bi = lk * 32 + lane;
        warp_argmax(bv, bi);
        if (bi < 0 || bi >= items) bi = 0;
        if (lane == 0) store_cand(a.st, cand, rnd, bi);
        if ((bi & 31) == lane) {
            int wk = bi >> 5;
#pragma unroll
            for (int k = 0; k < HC; ++k)
                if (k == wk) sc[k] = -INFINITY;
        }
    }
}

// generic path: any number of items, scores recomputed every round
__device__ __noinline__ void select_generic(const GLArgs &a, const double *row, double yy, int items, int *cand) {
    const int lane = threadIdx.x & 31;
    double prev_v = INFINITY;
    int prev_i = 0x7fffffff;
    for (int rnd = 0; rnd < a.st.Hp; ++rnd) {
        double best_v = -INFINITY;
        int best_i = -1;
        for (int i = lane; i < items; i += 32) {
            double v = sel_score(a, row, i, yy);
            bool below = (v < prev_v) || (v == prev_v && i < prev_i);
            bool better = (v > best_v) || (v == best_v && i > best_i);
            if (below && better) { best_v = v; best_i = i; }
        }
        warp_argmax(best_v, best_i);
        if (best_i < 0) best_i = 0;
        prev_v = best_v;
        prev_i = best_i;
        if (lane == 0) store_cand(a.st, cand, rnd, best_i);
    }
}

// =================================================================================================
// Kernel A -- row kernel: everything that is a function of the H-long score row.  One warp per
// datapoint, the row in shared memory, selection scores in registers, no state-space tables, so it
// runs at high occupancy.  Produces: candidates, their scores, the un-normalised singleton posteriors
// (written to the <s> row, relative to the singleton max m1) and the per-datapoint partial sums.
//   rs[n] = { m1, Z1, sig1, -, cnt1[0..5] }
// =================================================================================================
constexpr int RS = 4 + PET_MAXV;

__global__ void __launch_bounds__(ROW_WARPS * 32, 2) gl_row_kernel(const __grid_constant__ GLArgs a) {
    extern __shared__ __align__(16) double smem[];
    const GLStatic &st = a.st;
    const GLIter &it = a.it;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = st.H, Hp = st.Hp;
    double *row = smem + size_t(r2(H) + PET_MAXHP) * warp;
    int *cand_s = reinterpret_cast<int *>(row + r2(H));
    const bool rd = (a.flags & GLF_READ_LOGPJ) != 0, wr = (a.flags & GLF_WRITE_LOGPJ) != 0;
    const bool do_stats = !(a.flags & (GLF_LSE_ONLY | GLF_SELECT_ONLY));
    const int items = (st.select_mode == SEL_TSC) ? 2 * H : H;

    const int64_t wstride = int64_t(gridDim.x) * ROW_WARPS;
    for (int64_t r = int64_t(blockIdx.x) * ROW_WARPS + warp; r < a.n_rows; r += wstride) {
        const int64_t n = a.row0 + r;
        const double *yw = a.YW + r * st.ldH;
        const double yy = a.yy[n];
        for (int h = lane; h < H; h += 32) row[h] = yw[h];
        __syncwarp();
        // ---- top-H' preselection ---------------------------------------------------------------
        if (a.flags & GLF_SELECT) {
            if (items <= 32) select_regs<1>(a, row, yy, items, cand_s);
            else if (items <= 128) select_regs<4>(a, row, yy, items, cand_s);
            else if (items <= 1024) select_regs<32>(a, row, yy, items, cand_s);
            else select_generic(a, row, yy, items, cand_s);
            __syncwarp();
            if (lane < Hp) a.cand[n * Hp + lane] = cand_s[lane];
        } else {
            if (lane < Hp) cand_s[lane] = a.cand[n * Hp + lane];
            __syncwarp();
        }
        if (a.flags & GLF_SELECT_ONLY) { __syncwarp(); continue; }
        if (lane < Hp) a.ywc[n * Hp + lane] = row[cand_s[lane]];

        // ---- null state and all-H singleton blocks: log-joints and their max ------------------------
        double *logpj_row = a.logpj ? a.logpj + n * a.ld_logpj : nullptr;
        double m1 = -INFINITY, F0 = 0.0;
        if (st.has_null) {
            F0 = rd ? logpj_row[0] : combine(it, it.prior_null, yy);
            if (wr && lane == 0) logpj_row[0] = F0;
            m1 = F0;
        }
        for (int b = 0; b < st.n_blocks; ++b) {
            const double v = st.block_val[b];
            for (int h = lane; h < H; h += 32) {
                double F;
                if (rd) F = logpj_row[st.has_null + b * H + h];
                else {
                    double q = yy + v * (v * a.wn2[h] - 2.0 * row[h]);
                    F = combine(it, it.prior_block[b], q);
                    if (wr) logpj_row[st.has_null + b * H + h] = F;
                }
                m1 = fmax(m1, F);
            }
        }
        m1 = warp_max(m1);
        if (wr && (a.flags & GLF_LSE_ONLY)) { __syncwarp(); continue; }      // compat E_step: logpj only

        // ---- exp relative to m1, partial sums, un-normalised <s> row --------------------------------
        double Z1 = 0.0, sig1 = 0.0;
        double cntb[PET_MAXV];
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v) cntb[v] = 0.0;
        if (st.has_null && lane == 0) {
            double x = F0 - m1;
            double p = (x > GL_EXP_CUTOFF) ? exp_nonpos(x) : 0.0;
            Z1 += p;
            sig1 += p * yy;
        }
        for (int h = lane; h < st.ldH; h += 32) {
            double snew = 0.0, s2new = 0.0;
            if (h < H) {
                const double ywh = row[h], wn2h = (st.n_blocks > 0) ? a.wn2[h] : 0.0;
#pragma unroll
                for (int b = 0; b < PET_MAXV; ++b) {
                    if (b < st.n_blocks) {
                        const double v = st.block_val[b];
                        double q = yy + v * (v * wn2h - 2.0 * ywh);
                        double F = rd ? logpj_row[st.has_null + b * H + h] : combine(it, it.prior_block[b], q);
                        double x = F - m1;
                        double p = (x > GL_EXP_CUTOFF) ? exp_nonpos(x) : 0.0;
                        Z1 += p;
                        sig1 += p * q;
                        cntb[b] += p;
                        snew = fma(p, v, snew);
                        s2new = fma(p, v * v, s2new);
                    }
                }
            }
            if (do_stats) {
                a.S[r * st.ldH + h] = snew;              // scaled by exp(m1 - m)/Z in the scale kernel
                if (a.S2) a.S2[r * st.ldH + h] = s2new;
            }
        }
        Z1 = warp_sum(Z1);
        sig1 = warp_sum(sig1);
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v)
            if (v < st.n_blocks) cntb[v] = warp_sum(cntb[v]);
        if (lane == 0) {
            double *rs = a.rs + n * RS;
            rs[0] = m1; rs[1] = Z1; rs[2] = sig1;
#pragma unroll
            for (int v = 0; v < PET_MAXV; ++v) rs[4 + v] = cntb[v];
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------------
// Row kernel, fast path for the binary single-block layout [null | h = 0..H-1 | states] with H <= 1024,
// no logpj I/O and no second-moment row (BSC): same results as gl_row_kernel, ~1/3 of the instructions.
//   * selection: every lane keeps the best two of its 32 scores; a round is three warp REDUX.MAX over the
//     order-preserving integer image of the score (value, then item index: larger index wins ties, the
//     rule of the generic path) and a pop from the winning lane; a lane that has given both of its
//     entries rebuilds them from its registers (taken items masked)
//   * one latent value (1.0): no per-block loops
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long order_key(double v) {
    const long long b = __double_as_longlong(v);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}

template <int HC>
__device__ __forceinline__ void top2_of(const double (&sc)[HC], unsigned taken, double &v1, int &k1, double &v2, int &k2) {
    v1 = v2 = -INFINITY;
    k1 = k2 = 0;
#pragma unroll
    for (int k = 0; k < HC; ++k) {
        const double s = ((taken >> k) & 1u) ? -INFINITY : sc[k];
        const bool b1 = s >= v1, b2 = s >= v2;            // '>=': the later (larger) item index wins ties
        v2 = b1 ? v1 : (b2 ? s : v2);
        k2 = b1 ? k1 : (b2 ? k : k2);
        v1 = b1 ? s : v1;
        k1 = b1 ? k : k1;
    }
}

__global__ void __launch_bounds__(ROW_WARPS * 32, 3) gl_row_fast_kernel(const __grid_constant__ GLArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr int HC = 32;
    const GLStatic &st = a.st;
    const GLIter &it = a.it;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = st.H, Hp = st.Hp;
    double *row = smem + size_t(r2(H) + PET_MAXHP) * warp;
    int *cand_s = reinterpret_cast<int *>(row + r2(H));
    const bool do_stats = !(a.flags & (GLF_LSE_ONLY | GLF_SELECT_ONLY | GLF_NO_SROW));
    const double pb = it.prior_block[0];

    const int64_t wstride = int64_t(gridDim.x) * ROW_WARPS;
    for (int64_t r = int64_t(blockIdx.x) * ROW_WARPS + warp; r < a.n_rows; r += wstride) {
        const int64_t n = a.row0 + r;
        const double *yw = a.YW + r * st.ldH;
        const double yy = a.yy[n];
        double sc[HC];
        double m1 = -INFINITY;
        // ---- one pass over the score row: shared copy, selection scores, singleton log-joints' maximum ----
#pragma unroll
        for (int k = 0; k < HC; ++k) {
            const int h = k * 32 + lane;
            if (h < H) {
                const double v = yw[h];
                row[h] = v;
                double s = (st.select_mode == SEL_BSC) ? (a.wmu ? v + a.wmu[h] : v) * a.invn[h]
                         : (st.select_mode == SEL_NEGDIST) ? 2.0 * v - a.wn2[h] : -v;
                s += 0.0;                                  // -0.0 ties with +0.0, as in a floating-point compare
                sc[k] = (s != s) ? -INFINITY : s;          // NaN scores never win
                m1 = fmax(m1, combine(it, pb, yy + (a.wn2[h] - 2.0 * v)));
            } else {
                sc[k] = -INFINITY;
            }
        }
        __syncwarp();
        if (a.flags & GLF_SELECT) {
            unsigned taken = 0;
            double v1, v2;
            int k1, k2, left = 2;
            top2_of<HC>(sc, taken, v1, k1, v2, k2);
            for (int rnd = 0; rnd < Hp; ++rnd) {
                const unsigned long long key = order_key(v1);
                const unsigned hi = unsigned(key >> 32), lo = unsigned(key);
                const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
                const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
                const unsigned mine = (hi == mhi && lo == mlo) ? unsigned(k1 * 32 + lane) + 1u : 0u;
                int bi = int(__reduce_max_sync(0xffffffffu, mine)) - 1;
                if (bi < 0 || bi >= H) bi = 0;
                if (lane == 0) store_cand(st, cand_s, rnd, bi);          // BSC: ascending by score (bsc_et.py:113)
                if ((bi & 31) == lane && (bi >> 5) == k1 && mine != 0u) {
                    taken |= 1u << k1;
                    v1 = v2; k1 = k2; v2 = -INFINITY;
                    --left;
                }
                if (__any_sync(0xffffffffu, left == 0)) {                 // rare: a lane ran out of prepared entries
                    if (left == 0) { top2_of<HC>(sc, taken, v1, k1, v2, k2); left = 2; }
                }
            }
            __syncwarp();
            if (lane < Hp) a.cand[n * Hp + lane] = cand_s[lane];
        } else {
            if (lane < Hp) cand_s[lane] = a.cand[n * Hp + lane];
            __syncwarp();
        }
        if (a.flags & GLF_SELECT_ONLY) { __syncwarp(); continue; }
        if (lane < Hp) a.ywc[n * Hp + lane] = row[cand_s[lane]];

        const double F0 = combine(it, it.prior_null, yy);
        m1 = warp_max(fmax(m1, F0));
        double Z1 = 0.0, sig1 = 0.0, cnt = 0.0;
        if (lane == 0) {
            const double x = F0 - m1;
            const double p = (x > GL_EXP_CUTOFF) ? exp_nonpos(x) : 0.0;
            Z1 += p;
            sig1 += p * yy;
        }
        double *Srow = a.S + r * st.ldH;
#pragma unroll 4
        for (int h = lane; h < st.ldH; h += 32) {
            double p = 0.0;
            if (h < H) {
                const double q = yy + (a.wn2[h] - 2.0 * row[h]);
                const double x = combine(it, pb, q) - m1;
                p = (x > GL_EXP_CUTOFF) ? exp_nonpos(x) : 0.0;
                Z1 += p;
                sig1 = fma(p, q, sig1);
                cnt += p;
            }
            if (do_stats) Srow[h] = p;                       // scaled by exp(m1 - m)/Z in the scale kernel
        }
        Z1 = warp_sum(Z1);
        sig1 = warp_sum(sig1);
        cnt = warp_sum(cnt);
        if (lane == 0) {
            double *rs = a.rs + n * RS;
            rs[0] = m1; rs[1] = Z1; rs[2] = sig1; rs[4] = cnt;
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------------
// The fast path as TWO kernels (the default): the selection needs the 32 scores of a lane in registers and runs at
// 24 warps/SM; the singleton posterior is a streaming pass that needs few registers and runs at full occupancy,
// which is what hides the FP64 dependency chains of exp.  Results are those of gl_row_fast_kernel.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_WARPS * 32, 3) gl_row_select_kernel(const __grid_constant__ GLArgs a) {
    __shared__ int cand_all[ROW_WARPS][PET_MAXHP];
    constexpr int HC = 32;
    const GLStatic &st = a.st;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = st.H, Hp = st.Hp;
    int *cand_s = cand_all[warp];
    const int64_t wstride = int64_t(gridDim.x) * ROW_WARPS;
    for (int64_t r = int64_t(blockIdx.x) * ROW_WARPS + warp; r < a.n_rows; r += wstride) {
        const int64_t n = a.row0 + r;
        const double *yw = a.YW + r * st.ldH;
        if (a.flags & GLF_SELECT) {
            double sc[HC];
#pragma unroll
            for (int k = 0; k < HC; ++k) {
                const int h = k * 32 + lane;
                if (h < H) {
                    const double v = yw[h];
                    double s = (st.select_mode == SEL_BSC) ? (a.wmu ? v + a.wmu[h] : v) * a.invn[h]
                             : (st.select_mode == SEL_NEGDIST) ? 2.0 * v - a.wn2[h] : -v;
                    s += 0.0;                                  // -0.0 ties with +0.0, as in a floating-point compare
                    sc[k] = (s != s) ? -INFINITY : s;          // NaN scores never win
                } else {
                    sc[k] = -INFINITY;
                }
            }
            unsigned taken = 0;
            double v1, v2;
            int k1, k2, left = 2;
            top2_of<HC>(sc, taken, v1, k1, v2, k2);
            for (int rnd = 0; rnd < Hp; ++rnd) {
                const unsigned long long key = order_key(v1);
                const unsigned hi = unsigned(key >> 32), lo = unsigned(key);
                const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
                const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
                const unsigned mine = (hi == mhi && lo == mlo) ? unsigned(k1 * 32 + lane) + 1u : 0u;
                int bi = int(__reduce_max_sync(0xffffffffu, mine)) - 1;
                if (bi < 0 || bi >= H) bi = 0;
                if (lane == 0) store_cand(st, cand_s, rnd, bi);
                if ((bi & 31) == lane && (bi >> 5) == k1 && mine != 0u) {
                    taken |= 1u << k1;
                    v1 = v2; k1 = k2; v2 = -INFINITY;
                    --left;
                }
                if (__any_sync(0xffffffffu, left == 0)) {                 // rare: a lane ran out of prepared entries
                    if (left == 0) { top2_of<HC>(sc, taken, v1, k1, v2, k2); left = 2; }
                }
            }
            __syncwarp();
            if (lane < Hp) a.cand[n * Hp + lane] = cand_s[lane];
        } else {
            if (lane < Hp) cand_s[lane] = a.cand[n * Hp + lane];
            __syncwarp();
        }
        if (!(a.flags & GLF_SELECT_ONLY) && lane < Hp) a.ywc[n * Hp + lane] = yw[cand_s[lane]];
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------------
// Selection and singleton reductions in ONE pass over the score row (used whenever the <s> row is not written: the
// int8 statistics path, the log-denominator sweep, select_Hprimes).  The row lives in registers (32 scores per lane).
//   * selection by threshold on FLOAT32 keys, exact ranking in float64: rounding to float32 is monotone, so with tau =
//     the H'-th best of the 32 lane maxima (H' rounds of two 32-bit warp REDUX.MAX; each lane contributes one item, no
//     refills) every item whose key is below tau is below at least H' others in float64 as well.  The items with
//     key >= tau (H' + a few) are compacted into a shared-memory list with their float64 keys and ranked exactly there
//     (score descending, ties -> the larger item index first, as everywhere).  A list overflow (massive float32 ties)
//     falls back to the generic selection.
//   * null + all-H singleton log-joints from the same registers, F_h = (B + A yy) + A wn2_h - 2 A v_h with
//     A = beta pre1: max, partition sum, sigma / prior partial sums (table-based exp); nothing of length H is written.
// -------------------------------------------------------------------------------------------------
constexpr int ROWF_WARPS = 8, ROWF_LIST = 32 * PET_MAXHP;

__device__ __forceinline__ bool item_ge(unsigned long long ka, int ia, unsigned long long kb, int ib) {
    return ka > kb || (ka == kb && ia >= ib);
}
// selection score of item h from its score-row entry; sc1 = 1/||W_h|| (BSC) or ||W_h||^2 (NEGDIST), sc0 = W_h . mu (BSC)
template <int MODE, bool WMU>
__device__ __forceinline__ double row_score(double vk, double sc0, double sc1) {
    double s = (MODE == SEL_BSC) ? (WMU ? vk + sc0 : vk) * sc1 : (MODE == SEL_NEGDIST) ? 2.0 * vk - sc1 : -vk;
    s += 0.0;                                          // -0.0 ties with +0.0, as in a floating-point compare
    return (s != s) ? -INFINITY : s;                   // NaN scores never win
}
__device__ __forceinline__ unsigned key32(double s) {  // order-preserving image of the score rounded to float32; > 0
    const int b = __float_as_int(__double2float_rn(s));
    return unsigned(b ^ ((b >> 31) | int(0x80000000)));
}

template <int MODE, bool WMU>
__global__ void __launch_bounds__(ROWF_WARPS * 32, 2) gl_row_threshold_kernel(const __grid_constant__ GLArgs a) {
    __shared__ double exptab[32];
    __shared__ unsigned long long list_key[ROWF_WARPS][ROWF_LIST];
    __shared__ short list_idx[ROWF_WARPS][ROWF_LIST];
    __shared__ int cand_all[ROWF_WARPS][PET_MAXHP];
    constexpr int HC = 32;
    const GLStatic &st = a.st;
    const GLIter &it = a.it;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = st.H, Hp = st.Hp;
    int *cand_s = cand_all[warp];
    unsigned long long *lk = list_key[warp];
    short *li = list_idx[warp];
    // combine(prior, q) = B + A q
    const double A = it.beta * it.pre1, invA = 1.0 / A;
    const double Bp = it.anneal_prior ? it.beta * it.prior_block[0] : it.prior_block[0];
    const double B0 = it.anneal_prior ? it.beta * it.prior_null : it.prior_null;
    const double m2A = -2.0 * A;
    const double *sc1p = ((MODE == SEL_BSC) ? a.invn : a.wn2) + lane, *sc0p = (WMU ? a.wmu : a.wn2) + lane, *wn2p = a.wn2 + lane;
    const int nk = (H - lane + 31) >> 5;               // items of this lane: h = k * 32 + lane < H  <=>  k < nk
    const bool do_select = (a.flags & GLF_SELECT) != 0, select_only = (a.flags & GLF_SELECT_ONLY) != 0;
    exp_tab32_init(exptab);
    __syncthreads();
    const int64_t wstride = int64_t(gridDim.x) * ROWF_WARPS;
    for (int64_t r = int64_t(blockIdx.x) * ROWF_WARPS + warp; r < a.n_rows; r += wstride) {
        const int64_t n = a.row0 + r;
        const double *yw = a.YW + r * st.ldH;
        const double *ywl = yw + lane;
        // straight-line code over the 32 items of the lane: out-of-range items (h >= H) are masked by selects, never by
        // branches, so that the 32 dependency chains interleave
        double v[HC];
#pragma unroll
        for (int k = 0; k < HC; ++k) v[k] = (k < nk) ? ywl[k * 32] : 0.0;
        if (do_select) {
            // ---- float32 keys, best item of every lane ----
            unsigned key[HC];
            unsigned bk = 0u;
            int bidx = lane;
#pragma unroll
            for (int k = 0; k < HC; ++k) {
                const double s0 = (WMU && k < nk) ? sc0p[k * 32] : 0.0;
                const double s1 = (MODE != SEL_GIVEN && k < nk) ? sc1p[k * 32] : 0.0;
                unsigned kk = key32(row_score<MODE, WMU>(v[k], s0, s1));
                kk = (k < nk) ? kk : 0u;
                key[k] = kk;
                const bool better = kk >= bk;
                bk = better ? kk : bk;
                bidx = better ? k * 32 + lane : bidx;
            }
            // ---- tau = the H'-th best lane maximum ----
            unsigned tk = 0u;
            {
                unsigned ck = bk;
                for (int rnd = 0; rnd < Hp; ++rnd) {
                    tk = __reduce_max_sync(0xffffffffu, ck);
                    const unsigned mi = __reduce_max_sync(0xffffffffu, ck == tk ? unsigned(bidx) + 1u : 0u);
                    if (ck == tk && unsigned(bidx) + 1u == mi) ck = 0u;       // this lane's maximum is out
                }
            }
            // ---- compact the items with key >= tau ----
            unsigned mask = 0u;
#pragma unroll
            for (int k = 0; k < HC; ++k) mask |= (key[k] != 0u && key[k] >= tk) ? (1u << k) : 0u;
            const int cnt = __popc(mask);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t2;
            }
            const int M = __shfl_sync(0xffffffffu, incl, 31);
            if (M <= ROWF_LIST) {
                int off = incl - cnt;
                while (mask) {
                    const int k = __ffs(mask) - 1;
                    mask &= mask - 1;
                    li[off++] = short(k * 32 + lane);
                }
                __syncwarp();
                // ---- float64 keys of the listed items, exact rank inside the list (M >= H') ----
                for (int e = lane; e < M; e += 32) {
                    const int h = li[e];
                    lk[e] = order_key(row_score<MODE, WMU>(yw[h], WMU ? a.wmu[h] : 0.0,
                                                           (MODE == SEL_BSC) ? a.invn[h] : (MODE == SEL_NEGDIST) ? a.wn2[h] : 0.0));
                }
                __syncwarp();
                for (int e = lane; e < M; e += 32) {
                    const unsigned long long mk = lk[e];
                    const int mi = li[e];
                    int rank = 0;
                    for (int j = 0; j < M; ++j) rank += (j != e && item_ge(lk[j], li[j], mk, mi)) ? 1 : 0;
                    if (rank < Hp) store_cand(st, cand_s, rank, mi);
                }
            } else {
                select_generic(a, yw, a.yy[n], H, cand_s);
            }
            __syncwarp();
            if (lane < Hp) a.cand[n * Hp + lane] = cand_s[lane];
        } else {
            if (lane < Hp) cand_s[lane] = a.cand[n * Hp + lane];
            __syncwarp();
        }
        if (select_only) { __syncwarp(); continue; }
        if (lane < Hp) a.ywc[n * Hp + lane] = yw[cand_s[lane]];

        // ---- null state + all-H singletons: max, partition sum, sigma / prior partial sums ----
        const double yy = a.yy[n];
        double tmax = -INFINITY;
#pragma unroll
        for (int k = 0; k < HC; ++k) {
            const double w2 = (k < nk) ? wn2p[k * 32] : 0.0;
            const double th = fma(m2A, v[k], A * w2);                          // t_h = F_h - (B + A yy)
            v[k] = (k < nk) ? th : -1.0e300;
            tmax = fmax(tmax, v[k]);
        }
        tmax = warp_max(tmax);
        const double Ryy = fma(A, yy, Bp), F0 = fma(A, yy, B0);
        const double m1 = fmax(F0, Ryy + tmax);
        const double d = Ryy - m1;
        double Zs = 0.0, S1 = 0.0;
#pragma unroll
        for (int g = 0; g < HC / 8; ++g) {                                      // eight interleaved chains at a time
#pragma unroll
            for (int k = 8 * g; k < 8 * g + 8; ++k) {
                double p = exp_tab32(fmax(v[k] + d, GL_EXP_CUTOFF), exptab);
                p = (k < nk) ? p : 0.0;                                         // (masked items: t = -1e300 -> cutoff -> 0)
                Zs += p;
                S1 = fma(p, (k < nk) ? v[k] : 0.0, S1);
            }
            asm volatile("" ::: "memory");                                      // keep the groups apart: register pressure
        }
        Zs = warp_sum(Zs);
        S1 = warp_sum(S1);
        if (lane == 0) {
            const double x0 = F0 - m1;
            const double p0 = (x0 > GL_EXP_CUTOFF) ? exp_tab32(x0, exptab) : 0.0;
            double *rs = a.rs + n * RS;
            rs[0] = m1;
            rs[1] = Zs + p0;
            rs[2] = fma(p0 + Zs, yy, S1 * invA);             // sum_c p_c q_c,  q_h = yy + t_h / A
            rs[4] = Zs;
        }
        __syncwarp();
    }
}

// null state + all-H singletons of one datapoint per warp: max, partition sum, sigma / prior partial sums and the
// un-normalised posterior row.  F_h decreases with q_h = yy + wn2_h - 2 yw_h (beta pre1 < 0), so the maximum is the
// log-joint of the smallest q_h.
__global__ void __launch_bounds__(256, 4) gl_row_post_kernel(const __grid_constant__ GLArgs a) {
    __shared__ double exptab[32];
    exp_tab32_init(exptab);
    __syncthreads();
    const GLStatic &st = a.st;
    const GLIter &it = a.it;
    const int lane = threadIdx.x & 31;
    const int H = st.H;
    const bool do_stats = !(a.flags & (GLF_LSE_ONLY | GLF_NO_SROW));
    const double pb = it.prior_block[0];
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.n_rows) return;
    const int64_t n = a.row0 + r;
    const double *yw = a.YW + r * st.ldH;
    const double yy = a.yy[n];
    double tmin = INFINITY;
    for (int h = lane; h < H; h += 32) tmin = fmin(tmin, a.wn2[h] - 2.0 * yw[h]);
    tmin = -warp_max(-tmin);
    const double F0 = combine(it, it.prior_null, yy);
    const double m1 = fmax(F0, combine(it, pb, yy + tmin));
    double Z1 = 0.0, sig1 = 0.0, cnt = 0.0;
    if (lane == 0) {
        const double x = F0 - m1;
        const double p = (x > GL_EXP_CUTOFF) ? exp_tab32(x, exptab) : 0.0;
        Z1 += p;
        sig1 += p * yy;
    }
    double *Srow = a.S + r * st.ldH;
#pragma unroll 2
    for (int h = lane; h < st.ldH; h += 32) {
        double p = 0.0;
        if (h < H) {
            const double q = yy + (a.wn2[h] - 2.0 * yw[h]);
            const double x = combine(it, pb, q) - m1;
            p = (x > GL_EXP_CUTOFF) ? exp_tab32(x, exptab) : 0.0;
            Z1 += p;
            sig1 = fma(p, q, sig1);
            cnt += p;
        }
        if (do_stats) Srow[h] = p;                           // scaled by exp(m1 - m)/Z downstream
    }
    Z1 = warp_sum(Z1);
    sig1 = warp_sum(sig1);
    cnt = warp_sum(cnt);
    if (lane == 0) {
        double *rs = a.rs + n * RS;
        rs[0] = m1; rs[1] = Z1; rs[2] = sig1; rs[4] = cnt;
    }
}

// =================================================================================================
// Kernel B -- state kernel: the truncated multi-cause state space of one datapoint, TWO warps per
// datapoint (the shared-memory footprint is per datapoint, so more lanes per datapoint is what hides
// latency).  Combines with the row kernel's partial sums by log-sum-exp merging:
//   m = max(m1, m2),  Z = Z1 e^(m1-m) + Z2,  lse = m + log Z.
// =================================================================================================
__device__ __forceinline__ void grp_sync(int gid) { asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "n"(GRP_LANES) : "memory"); }

__device__ __forceinline__ double grp_max(double v, double *red, int gid, int wig) {
    v = warp_max(v);
    grp_sync(gid);
    if ((threadIdx.x & 31) == 0) red[wig] = v;
    grp_sync(gid);
    double m = red[0];
#pragma unroll
    for (int w = 1; w < GRP_WARPS; ++w) m = fmax(m, red[w]);
    return m;
}
__device__ __forceinline__ double grp_sum(double v, double *red, int gid, int wig) {
    v = warp_sum(v);
    grp_sync(gid);
    if ((threadIdx.x & 31) == 0) red[wig] = v;
    grp_sync(gid);
    double t = red[0];
#pragma unroll
    for (int w = 1; w < GRP_WARPS; ++w) t += red[w];
    return t;
}

template <int GMAX, bool BINARY>
__global__ void __launch_bounds__(GL_MAX_GROUPS * GRP_LANES) gl_state_kernel(const __grid_constant__ GLArgs a) {
    extern __shared__ __align__(16) double smem[];
    const GLStatic &st = a.st;
    const GLIter &it = a.it;
    const int ngroups = blockDim.x / GRP_LANES;
    const int gid = threadIdx.x / GRP_LANES, l64 = threadIdx.x % GRP_LANES, wig = l64 >> 5;
    const int Hp = st.Hp, S = st.S, H = st.H;
    const SmemLayout L = smem_layout(st);
    const int CH = GL_CHUNK, NCH = st.n_chunks;

    unsigned long long *states_s = reinterpret_cast<unsigned long long *>(smem);
    unsigned short *ids_s = reinterpret_cast<unsigned short *>(smem + L.off_ids);
    unsigned short *chunk_s = reinterpret_cast<unsigned short *>(smem + L.off_chunk);
    const bool inc = BINARY && GMAX <= 5 && st.inc_states != nullptr;
    for (int s = threadIdx.x; s < S; s += blockDim.x) states_s[s] = inc ? st.inc_states[s] : st.states[s];
    for (int i = threadIdx.x; i < NCH * CH * GRP_LANES; i += blockDim.x) ids_s[i] = st.entries[i];
    for (int i = threadIdx.x; i < NCH * GRP_LANES; i += blockDim.x) chunk_s[i] = st.chunk_tab[i];
    double *gb = smem + L.shared_doubles + size_t(L.per_dp) * gid;
    double *qbuf = gb;
    double *Gc = gb + L.off_G;
    double *lin = gb + L.off_lin;
    double *mom = gb + L.off_mom;
    double *Pj = gb + L.off_P;
    double *red = gb + L.off_red;
    int *cand_s = reinterpret_cast<int *>(gb + L.off_cand);
    int *live_s = cand_s + PET_MAXHP;
    for (int i = l64; i < GS * GS; i += GRP_LANES) Gc[i] = 0.0;       // row/column Hp stay zero: dummy position
    for (int i = l64; i < GS; i += GRP_LANES) lin[i] = 0.0;
    __syncthreads();

    const bool do_stats = !(a.flags & (GLF_LSE_ONLY | GLF_SELECT_ONLY));
    const int n_cnt = st.n_cnt, n_g = st.n_g;
    const double cut = (a.flags & GLF_USE_CUT) ? *a.cut : 0.0;
    const bool rd = (a.flags & GLF_READ_LOGPJ) != 0, wr = (a.flags & GLF_WRITE_LOGPJ) != 0;
    const int col_states = st.has_null + st.n_blocks * H;
    SizeBounds<GMAX> sb;
    sb.load(st);

    double acc_n = 0.0, acc_lse = 0.0, acc_sig = 0.0;      // lane 0 of each group accumulates
    double acc_cnt[PET_MAXV];
#pragma unroll
    for (int v = 0; v < PET_MAXV; ++v) acc_cnt[v] = 0.0;

    const int64_t gstride = int64_t(gridDim.x) * ngroups;
    for (int64_t r = int64_t(blockIdx.x) * ngroups + gid; r < a.n_rows; r += gstride) {
        const int64_t n = a.row0 + r;
        const double yy = a.yy[n];
        double *scl = a.scl + n * (1 + PET_MAXHP);
        if ((a.flags & GLF_USE_CUT) && do_stats) {
            double l = a.lse[n];
            bool keep = (a.flags & GLF_CUT_STRICT) ? (l > cut) : (l >= cut);
            if (!keep) {   // truncated away: contributes nothing (bsc_et.py:254-257)
                if (l64 <= Hp) scl[l64] = 0.0;
                continue;
            }
        }
        // ---- gather ------------------------------------------------------------------------------
        grp_sync(gid);
        if (l64 < Hp) cand_s[l64] = a.cand[n * Hp + l64];
        grp_sync(gid);
        for (int idx = l64; idx < Hp * Hp; idx += GRP_LANES) {
            int j = idx / Hp, k = idx % Hp;
            Gc[j * GS + k] = a.G[int64_t(cand_s[j]) * st.ldH + cand_s[k]];
        }
        if (l64 < Hp) {
            int c = cand_s[l64];
            int lv = 1;
            for (int j = l64 + 1; j < Hp; ++j) lv &= (cand_s[j] != c);   // numpy "last write wins"
            live_s[l64] = lv;
        }
        grp_sync(gid);
        if (l64 < Hp) {
            double ywc = a.ywc[n * Hp + l64];
            lin[l64] = BINARY ? fma(-2.0, ywc, Gc[l64 * GS + l64]) : ywc;
        }
        grp_sync(gid);

        double *logpj_row = a.logpj ? a.logpj + n * a.ld_logpj : nullptr;
        // ---- log-joints of the multi-cause states, max ------------------------------------------------
        double m2 = -INFINITY;
        if (inc) {
            // size by size: q(s) = q(s without its largest member e) + lin[e] + 2 sum_i G[i][e]; the shorter state
            // was evaluated one level earlier (singletons: yy + lin[j], kept behind the zero slot)
            if (l64 < Hp) qbuf[S + 1 + l64] = yy + lin[l64];
            grp_sync(gid);
#pragma unroll
            for (int g = 2; g <= GMAX; ++g) {
                const int s_end = (g < GMAX) ? st.size_start[g + 1] : S;
                const double pr = it.lp[0] * double(g);
                for (int s = st.size_start[g] + l64; s < s_end; s += GRP_LANES) {
                    const unsigned long long rec = states_s[s];
                    const int e = int(unsigned(rec) & 0xFFu);
                    const double *ge = Gc + e;
                    double cross = 0.0;
#pragma unroll
                    for (int m = 1; m < GMAX; ++m) cross += ge[int(unsigned(rec >> (8 * m)) & 0xFFu) * GS];
                    const double q = fma(2.0, cross, qbuf[int(rec >> 48)] + lin[e]);
                    qbuf[s] = q;
                    double F;
                    if (rd) F = logpj_row[col_states + s];
                    else {
                        F = combine(it, pr, q);
                        if (wr) logpj_row[col_states + s] = F;
                    }
                    m2 = fmax(m2, F);
                }
                if (g < GMAX) grp_sync(gid);
            }
        } else {
#pragma unroll 2
            for (int s = l64; s < S; s += GRP_LANES) {
                double q = eval_state<GMAX, BINARY>(states_s[s], st, lin, Gc, yy);
                qbuf[s] = q;
                double F;
                if (rd) F = logpj_row[col_states + s];
                else {
                    F = combine(it, prior_of<GMAX, BINARY>(a, sb, s), q);
                    if (wr) logpj_row[col_states + s] = F;
                }
                m2 = fmax(m2, F);
            }
        }
        if (wr && (a.flags & GLF_LSE_ONLY)) continue;   // compat E_step: logpj only
        m2 = grp_max(m2, red, gid, wig);
        const double *rs = a.rs + n * RS;
        const double m1 = rs[0];
        const double mx = fmax(m1, m2);
        // ---- exp, merge with the singleton partial sums -------------------------------------------------
        double Z2 = 0.0, sig2 = 0.0;
        if (inc) {
#pragma unroll
            for (int g = 2; g <= GMAX; ++g) {                 // by size: the log-prior is a constant of the level
                const int s_end = (g < GMAX) ? st.size_start[g + 1] : S;
                const double pr = it.lp[0] * double(g);
#pragma unroll 2
                for (int s = st.size_start[g] + l64; s < s_end; s += GRP_LANES) {
                    double q = qbuf[s];
                    double F = rd ? logpj_row[col_states + s] : combine(it, pr, q);
                    double x = F - mx;
                    double p = (x > GL_EXP_CUTOFF) ? exp_nonpos(x) : 0.0;
                    Z2 += p;
                    sig2 += p * q;
                    qbuf[s] = p;
                }
            }
        } else {
#pragma unroll 2
            for (int s = l64; s < S; s += GRP_LANES) {
                double q = qbuf[s];
                double F = rd ? logpj_row[col_states + s] : combine(it, prior_of<GMAX, BINARY>(a, sb, s), q);
                double x = F - mx;
                double p = (x > GL_EXP_CUTOFF) ? exp_nonpos(x) : 0.0;
                Z2 += p;
                sig2 += p * q;
                qbuf[s] = p;
            }
        }
        if (l64 == 0) qbuf[S] = 0.0;   // zero slot read by padding entries of the gather table
        Z2 = grp_sum(Z2, red, gid, wig);
        const double e1 = (m1 == -INFINITY) ? 0.0 : exp(m1 - mx);
        const double Z = fma(rs[1], e1, Z2);
        const double lse = mx + log(Z);
        if (l64 == 0) a.lse[n] = lse;
        if (!do_stats) continue;
        sig2 = grp_sum(sig2, red, gid, wig);
        const double inv = 1.0 / Z;

        // ---- pair sums from the shared-memory gather table (64 lanes, exclusive owners) ---------------
        for (int o = l64; o <= st.n_out; o += GRP_LANES) mom[o] = 0.0;
        grp_sync(gid);
        for (int i = l64; i < st.n_direct; i += GRP_LANES) {
            unsigned d = st.direct[i];
            mom[d >> 16] = qbuf[d & 0xFFFFu];
        }
        {
            double acc = 0.0;
            const unsigned short *e = ids_s + l64;            // [(chunk*8 + i)*64 + lane]
            const unsigned short *ct = chunk_s + l64;
            for (int c = 0; c < NCH; ++c, e += GL_CHUNK * GRP_LANES, ct += GRP_LANES) {
                double s0 = qbuf[e[0]] + qbuf[e[GRP_LANES]];
                double s1 = qbuf[e[2 * GRP_LANES]] + qbuf[e[3 * GRP_LANES]];
                double s2 = qbuf[e[4 * GRP_LANES]] + qbuf[e[5 * GRP_LANES]];
                double s3 = qbuf[e[6 * GRP_LANES]] + qbuf[e[7 * GRP_LANES]];
                acc += (s0 + s1) + (s2 + s3);
                const unsigned co = *ct;
                if (co & 0x8000u) { mom[co & 0x7FFFu] = acc; acc = 0.0; }
            }
        }
        grp_sync(gid);

        // ---- moments -------------------------------------------------------------------------------
        // P[j][a] = posterior mass of (s_j = v_a): singleton state + size-weighted pair sums
        for (int idx = l64; idx < Hp * n_cnt; idx += GRP_LANES) {
            const int j = idx / n_cnt, av = idx % n_cnt;
            const int sid = st.single_idx[idx];
            double P = (sid >= 0) ? qbuf[sid] : 0.0;
            for (int k = 0; k < Hp; ++k) {
                if (k == j) continue;
                const int lo = min(j, k), hi = max(j, k);
                const int pair = lo * Hp - lo * (lo + 1) / 2 + (hi - lo - 1);
                for (int bv = 0; bv < n_cnt; ++bv) {
                    const int base = ((pair * n_cnt + (j < k ? av : bv)) * n_cnt + (j < k ? bv : av)) * n_g;
                    for (int g = 0; g < n_g; ++g) P = fma(mom[base + g], c_winv[g], P);
                }
            }
            Pj[idx] = P;
        }
        grp_sync(gid);
        // scale of the singleton row and the candidate marginals for the scale kernel
        double cnt_states[PET_MAXV];
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v) cnt_states[v] = 0.0;
        const bool fold = (a.flags & GLF_FOLD_SCALE) != 0;
        double sce = e1 * inv;                                // <s>[n,:] = singles[n,:] * sce (+ candidate marginals)
        double *Srow = a.S + r * st.ldH;
        if (fold) {
            if (sce == 0.0) {       // the singletons vanish next to the multi-cause states: zero row, unit scale
                for (int h = l64; h < st.ldH; h += GRP_LANES) Srow[h] = 0.0;
                sce = 1.0;
            }
            grp_sync(gid);
        }
        if (l64 == 0) scl[0] = sce;
        if (l64 < Hp) {
            double m1j = 0.0;
#pragma unroll
            for (int v = 0; v < PET_MAXV; ++v)
                if (v < n_cnt) {
                    double P = Pj[l64 * n_cnt + v];
                    m1j = fma(BINARY ? 1.0 : st.vals[v], P, m1j);
                    cnt_states[v] = P;
                }
            const double mj = live_s[l64] ? m1j * inv : 0.0;  // non-live duplicates carry 0 and do not write
            if (fold) {
                if (mj != 0.0) Srow[cand_s[l64]] += mj / sce;
                scl[1 + l64] = 0.0;
            } else {
                scl[1 + l64] = mj;
            }
        }
        // second moments scattered into Wq (numpy fancy-index semantics for duplicates)
        for (int idx = l64; idx < Hp * Hp; idx += GRP_LANES) {
            const int j = idx / Hp, k = idx % Hp;
            if (!(live_s[j] && live_s[k])) continue;
            double m2v = 0.0;
            if (j == k) {
                if (st.diag_from_colsum) continue;
                for (int v = 0; v < n_cnt; ++v) {
                    double vv = BINARY ? 1.0 : st.vals[v];
                    m2v = fma(vv * vv, Pj[j * n_cnt + v], m2v);
                }
            } else {
                const int lo = min(j, k), hi = max(j, k);
                const int pair = lo * Hp - lo * (lo + 1) / 2 + (hi - lo - 1);
                for (int av = 0; av < n_cnt; ++av)
                    for (int bv = 0; bv < n_cnt; ++bv) {
                        double w = BINARY ? 1.0 : st.vals[av] * st.vals[bv];
                        const int base = ((pair * n_cnt + av) * n_cnt + bv) * n_g;
                        double sacc = 0.0;
                        for (int g = 0; g < n_g; ++g) sacc += mom[base + g];
                        m2v = fma(w, sacc, m2v);
                    }
            }
            if (m2v != 0.0) atomicAdd(&a.Wq[int64_t(cand_s[j]) * st.ldH + cand_s[k]], m2v * inv);
        }
        // scalar statistics (only warp 0 of the group holds candidate lanes)
        if (wig == 0) {
#pragma unroll
            for (int v = 0; v < PET_MAXV; ++v)
                if (v < n_cnt) cnt_states[v] = warp_sum(cnt_states[v]);
            if (l64 == 0) {
                acc_n += 1.0;
                acc_lse += lse;
                acc_sig += fma(rs[2], e1, sig2) * inv;
#pragma unroll
                for (int v = 0; v < PET_MAXV; ++v) {
                    double c = cnt_states[v];
#pragma unroll
                    for (int b = 0; b < PET_MAXV; ++b)
                        if (b < st.n_blocks && st.block_vidx[b] == v) c = fma(rs[4 + b], e1, c);
                    acc_cnt[v] += c * inv;
                }
            }
        }
    }
    if (do_stats && l64 == 0) {
        atomicAdd(&a.scalars[0], acc_n);
        atomicAdd(&a.scalars[1], acc_lse);
        atomicAdd(&a.scalars[2], acc_sig);
        for (int v = 0; v < st.n_cnt; ++v) atomicAdd(&a.scalars[3 + v], acc_cnt[v]);
    }
}

// =================================================================================================
// Kernel C -- scale kernel: <s>[n,:] = singles[n,:] * e^(m1-m)/Z  (+ candidate marginals).
// =================================================================================================
__global__ void __launch_bounds__(256) gl_scale_kernel(const __grid_constant__ GLArgs a) {
    const GLStatic &st = a.st;
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.n_rows) return;
    const int64_t n = a.row0 + r;
    const double *scl = a.scl + n * (1 + PET_MAXHP);
    const double sc = scl[0];
    double *Srow = a.S + r * st.ldH;
    for (int h = lane; h < st.ldH; h += 32) {
        Srow[h] *= sc;
        if (a.S2) a.S2[r * st.ldH + h] *= sc;
    }
    __syncwarp();
    if (lane < st.Hp) {
        double mj = scl[1 + lane];
        if (mj != 0.0) Srow[a.cand[n * st.Hp + lane]] += mj;      // non-live duplicates carry 0 and do not write
    }
}


// =================================================================================================
// Kernel D -- posterior rows straight into the int8 slices of the statistics GEMM (binary single-block layout).
// Replaces, for the fused int8 path: the <s> write of the singleton kernel, the read-modify-write of the candidate
// marginals, the column-maximum pass and the slicing pass (bsc_et.py:349,355,362-366).  CTA tile: 128 datapoints x 32
// causes; thread (tx, ty) owns cause c0 + tx and the 16 consecutive datapoints r0 + 16 ty ..; the candidate marginals of
// the tile are scattered into a shared-memory tile first, the digits are staged in shared memory so that both the
// loads (score rows) and the stores (slice rows, datapoints contiguous) are coalesced.
// =================================================================================================
constexpr int PS_R = 128, PS_C = 32, PS_PITCH = PS_R + 4;
// FULL: the tile lies inside the chunk and the cause range (no bounds checks in the element loop)
template <bool FULL>
__global__ void __launch_bounds__(256) gl_post_slice_kernel(const __grid_constant__ GLArgs a, int ns, int Kp, int8_t *out,
                                                            int64_t row_stride, int64_t slice_stride, double *scale_out,
                                                            int col_tile0, int row_tile0) {
    extern __shared__ __align__(16) int8_t ps_smem[];
    double *add = reinterpret_cast<double *>(ps_smem);                     // [PS_R][PS_C + 1] candidate marginals
    double *rowc = add + PS_R * (PS_C + 1);                                // [PS_R][2]  (B + A yy - m1), scale
    double *exptab = rowc + PS_R * 2;                                      // [32]
    int32_t *tile32 = reinterpret_cast<int32_t *>(exptab + 32);            // [ns][PS_C][PS_PITCH] bytes
    const GLStatic &st = a.st;
    const GLIter &it = a.it;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = (blockIdx.x + col_tile0) * PS_C, c = c0 + tx;
    const int64_t r0 = int64_t(blockIdx.y + row_tile0) * PS_R;
    const int Hp = st.Hp;
    // combine(prior, q) = B + A q,  q = yy + wn2_h - 2 v:  x = F_h - m1 = (B + A yy - m1) + A wn2_h - 2 A v
    const double A = it.beta * it.pre1, m2A = -2.0 * A;
    const double Bp = it.anneal_prior ? it.beta * it.prior_block[0] : it.prior_block[0];
    exp_tab32_init(exptab);
    for (int i = threadIdx.x; i < PS_R * (PS_C + 1); i += 256) add[i] = 0.0;
    if (threadIdx.x < PS_R) {
        const int64_t rr = r0 + threadIdx.x;
        double rt = -1.0e300, sc = 0.0;                                    // rows beyond the chunk: exp -> cutoff, scale 0
        if (rr < a.n_rows) {
            const int64_t n = a.row0 + rr;
            rt = fma(A, a.yy[n], Bp) - a.rs[n * RS];
            sc = a.scl[n * (1 + PET_MAXHP)];
        }
        rowc[threadIdx.x * 2 + 0] = rt; rowc[threadIdx.x * 2 + 1] = sc;
    }
    __syncthreads();
    {                                                                      // candidates are distinct: plain stores
        const int rl = threadIdx.x >> 1;                                   // two threads per datapoint, alternate candidates
        const int64_t rr = r0 + rl;
        if (rr < a.n_rows) {
            const int64_t n = a.row0 + rr;
            for (int j = threadIdx.x & 1; j < Hp; j += 2) {
                const int h = a.cand[n * Hp + j] - c0;
                if (h >= 0 && h < PS_C) add[rl * (PS_C + 1) + h] = a.scl[n * (1 + PET_MAXHP) + 1 + j];
            }
        }
    }
    __syncthreads();
    const bool col_ok = FULL || c < st.H;
    const double awn2 = col_ok ? A * a.wn2[c] : 0.0;
    // <s> in [0, 1] as the integer rint(<s> 2^48): slice 0 = bits 42.. (<= 64), slice t = the 7 bits below (0..127, valid
    // as signed bytes); value = sum_t slice_t 2^-(6+7t), i.e. the slicing convention of ozaki.cu with column scale 1
    const double *ywp = a.YW + (r0 + ty * 16) * st.ldH + c;
    const double *rc = rowc + ty * 32, *ad = add + ty * 16 * (PS_C + 1) + tx;
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        double vv = 0.0;
        if (FULL || (col_ok && r0 + ty * 16 + i < a.n_rows)) vv = ywp[int64_t(i) * st.ldH];
        const double x = fmax(fma(m2A, vv, awn2 + rc[2 * i]), GL_EXP_CUTOFF);
        const double p = exp_tab32(x, exptab);                             // e^-100 2^48 rounds to 0: no select needed
        double val = fma(p, rc[2 * i + 1], ad[i * (PS_C + 1)]);
        if (!FULL && !col_ok) val = 0.0;
        const long long xi = __double2ll_rn(val * 281474976710656.0);      // 2^48; val in [0, 1 + eps]
        lo[i] = uint32_t(xi);
        hi[i] = uint32_t(xi >> 32);
    }
    if (blockIdx.y + row_tile0 == 0 && ty == 0 && col_ok) scale_out[c] = 1.0;
#pragma unroll
    for (int t = 0; t < 7; ++t) {
        if (t < ns) {
            const int sh = 42 - 7 * t;                                      // slice t = bits [sh, sh + 7)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t w = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = 4 * g + i;
                    uint32_t d;
                    if (sh >= 32) d = hi[e] >> (sh - 32);
                    else if (sh + 7 <= 32) d = lo[e] >> sh;
                    else d = __funnelshift_r(lo[e], hi[e], sh);
                    w |= (d & 0x7Fu) << (8 * i);
                }
                tile32[((t * PS_C + tx) * PS_PITCH) / 4 + ty * 4 + g] = int32_t(w);
            }
        }
    }
    __syncthreads();
    const int n_rows_out = ns * PS_C;
    for (int ro = ty; ro < n_rows_out; ro += 8) {                          // one warp stores one (slice, cause) row of 128 bytes
        const int t = ro / PS_C, cc = ro % PS_C;
        const int col = c0 + cc;
        const int64_t r = r0 + tx * 4;
        if ((FULL || col < st.H) && r < Kp)
            *reinterpret_cast<int32_t *>(out + t * slice_stride + col * row_stride + r) = tile32[((t * PS_C + cc) * PS_PITCH) / 4 + tx];
    }
}

int launch_gl_post_slice(const GLArgs &a, int ns, int Kp, int8_t *out, int64_t row_stride, int64_t slice_stride, double *scale_out,
                         cudaStream_t st) {
    if (a.n_rows <= 0) return PET_OK;
    const size_t smem = size_t(PS_R) * (PS_C + 1) * 8 + size_t(PS_R) * 2 * 8 + 32 * 8 + size_t(ns) * PS_C * PS_PITCH;
    static size_t configured = 0;
    if (smem > configured) {
        PET_CUDA(cudaFuncSetAttribute(gl_post_slice_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        PET_CUDA(cudaFuncSetAttribute(gl_post_slice_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    // interior tiles (whole 128-row, 32-cause tiles) without bounds checks, then the ragged right / bottom borders
    const unsigned cols_full = unsigned(a.st.H / PS_C), rows_full = unsigned(a.n_rows / PS_R);
    const unsigned cols_all = (unsigned)ceil_div(a.st.H, PS_C), rows_all = (unsigned)ceil_div(Kp, PS_R);
    if (cols_full > 0 && rows_full > 0) {
        gl_post_slice_kernel<true><<<dim3(cols_full, rows_full), 256, smem, st>>>(a, ns, Kp, out, row_stride, slice_stride, scale_out, 0, 0);
        PET_LAUNCH_CHECK();
    }
    if (cols_all > cols_full && rows_full > 0) {                           // right border: the last, partial cause tile
        gl_post_slice_kernel<false><<<dim3(cols_all - cols_full, rows_full), 256, smem, st>>>(a, ns, Kp, out, row_stride, slice_stride,
                                                                                               scale_out, int(cols_full), 0);
        PET_LAUNCH_CHECK();
    }
    if (rows_all > rows_full) {                                            // bottom border: all cause tiles of the last rows
        gl_post_slice_kernel<false><<<dim3(cols_all, rows_all - rows_full), 256, smem, st>>>(a, ns, Kp, out, row_stride, slice_stride,
                                                                                              scale_out, 0, int(rows_full));
        PET_LAUNCH_CHECK();
    }
    return PET_OK;
}

// datapoint groups per CTA of the state kernel whose shared memory fits one SM (0 = does not fit at all)
int gl_pick_warps(const GLStatic &s) {
    static int cap = []() { const char *e = getenv("PET_GL_GROUPS"); int v = e ? atoi(e) : GL_MAX_GROUPS; return v < 1 ? 1 : (v > GL_MAX_GROUPS ? GL_MAX_GROUPS : v); }();
    for (int w = cap; w >= 1; --w)
        if (gl_smem_bytes(s, w) <= 227 * 1024) return w;
    return 0;
}

template <int GMAX, bool BINARY>
static int launch_state(const GLArgs &a, int sm_count, cudaStream_t stream) {
    auto kern = gl_state_kernel<GMAX, BINARY>;
    const int groups = gl_pick_warps(a.st);
    if (groups == 0) {
        set_error("state kernel needs %zu bytes of shared memory per datapoint (H'=%d, states=%d): unsupported size",
                  gl_smem_bytes(a.st, 1), a.st.Hp, a.st.S);
        return PET_EINVAL;
    }
    const size_t smem = gl_smem_bytes(a.st, groups);
    static size_t configured = 0;
    if (smem > configured) {
        PET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    int per_sm = int(std::max<size_t>(1, std::min<size_t>(4, (227 * 1024) / (smem + 1024))));
    per_sm = std::min(per_sm, std::max(1, 64 / (GRP_WARPS * groups)));
    int64_t want = ceil_div(a.n_rows, groups);
    int64_t grid = std::min<int64_t>(want, int64_t(sm_count) * per_sm);
    if (grid <= 0) return PET_OK;
    kern<<<(unsigned)grid, groups * GRP_LANES, smem, stream>>>(a);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// The posterior pipeline over a chunk is: row kernel -> state kernel -> scale kernel.
int launch_gl_row(const GLArgs &a, int sm_count, cudaStream_t stream) {
    if (a.n_rows <= 0) return PET_OK;
    const size_t smem = size_t(r2(a.st.H) + PET_MAXHP) * ROW_WARPS * sizeof(double);
    if (smem > 227 * 1024) {
        set_error("row kernel needs %zu bytes of shared memory (H=%d): unsupported size", smem, a.st.H);
        return PET_EINVAL;
    }
    static size_t configured = 0;
    if (smem > configured) {
        PET_CUDA(cudaFuncSetAttribute(gl_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    int per_sm = int(std::max<size_t>(1, std::min<size_t>(2, (227 * 1024) / (smem + 1024))));
    int64_t grid = std::min<int64_t>(ceil_div(a.n_rows, ROW_WARPS), int64_t(sm_count) * per_sm);
    static const bool no_fast = getenv("PET_GL_NO_FAST_ROW") != nullptr;
    const bool fast = !no_fast && a.st.n_blocks == 1 && a.st.block_val[0] == 1.0 && a.st.has_null && a.st.H <= 1024 &&
                      !a.S2 && !(a.flags & (GLF_READ_LOGPJ | GLF_WRITE_LOGPJ)) &&
                      (a.st.select_mode == SEL_BSC || a.st.select_mode == SEL_NEGDIST || a.st.select_mode == SEL_GIVEN);
    if (fast) {
        per_sm = int(std::max<size_t>(1, std::min<size_t>(3, (227 * 1024) / (smem + 1024))));
        grid = std::min<int64_t>(ceil_div(a.n_rows, ROW_WARPS), int64_t(sm_count) * per_sm);
        static size_t configured_fast = 0;
        if (smem > configured_fast) {
            PET_CUDA(cudaFuncSetAttribute(gl_row_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            configured_fast = smem;
        }
        static const bool one_kernel = getenv("PET_GL_ROW_FUSED") != nullptr;
        if (one_kernel) {
            gl_row_fast_kernel<<<(unsigned)grid, ROW_WARPS * 32, smem, stream>>>(a);
            PET_LAUNCH_CHECK();
            return PET_OK;
        }
        static const bool no_threshold = getenv("PET_GL_NO_THRESHOLD_ROW") != nullptr;
        if (!no_threshold && (a.flags & (GLF_NO_SROW | GLF_LSE_ONLY | GLF_SELECT_ONLY))) {      // nothing of length H to write
            grid = std::min<int64_t>(ceil_div(a.n_rows, ROWF_WARPS), int64_t(sm_count) * 2 * 4);
            if (a.st.select_mode == SEL_BSC && a.wmu) gl_row_threshold_kernel<SEL_BSC, true><<<(unsigned)grid, ROWF_WARPS * 32, 0, stream>>>(a);
            else if (a.st.select_mode == SEL_BSC) gl_row_threshold_kernel<SEL_BSC, false><<<(unsigned)grid, ROWF_WARPS * 32, 0, stream>>>(a);
            else if (a.st.select_mode == SEL_NEGDIST) gl_row_threshold_kernel<SEL_NEGDIST, false><<<(unsigned)grid, ROWF_WARPS * 32, 0, stream>>>(a);
            else gl_row_threshold_kernel<SEL_GIVEN, false><<<(unsigned)grid, ROWF_WARPS * 32, 0, stream>>>(a);
            PET_LAUNCH_CHECK();
            return PET_OK;
        }
        grid = std::min<int64_t>(ceil_div(a.n_rows, ROW_WARPS), int64_t(sm_count) * 3);
        gl_row_select_kernel<<<(unsigned)grid, ROW_WARPS * 32, 0, stream>>>(a);
        PET_LAUNCH_CHECK();
        if (!(a.flags & GLF_SELECT_ONLY)) {
            gl_row_post_kernel<<<(unsigned)ceil_div(a.n_rows * 32, 256), 256, 0, stream>>>(a);
            PET_LAUNCH_CHECK();
        }
        return PET_OK;
    }
    gl_row_kernel<<<(unsigned)grid, ROW_WARPS * 32, smem, stream>>>(a);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

int launch_gl_state(const GLArgs &a, int gamma, bool binary, int sm_count, cudaStream_t stream) {
    if (a.n_rows <= 0 || (a.flags & GLF_SELECT_ONLY)) return PET_OK;
    if (binary) {
        if (gamma <= 3) return launch_state<3, true>(a, sm_count, stream);
        if (gamma <= 5) return launch_state<5, true>(a, sm_count, stream);
        return launch_state<8, true>(a, sm_count, stream);
    }
    if (gamma <= 4) return launch_state<4, false>(a, sm_count, stream);
    return launch_state<8, false>(a, sm_count, stream);
}

// Second half of a GLF_DEFER_STATS evaluation: one thread per datapoint of the chunk.
__global__ void __launch_bounds__(256) gl_finalize_cut_kernel(const __grid_constant__ GLArgs a) {
    __shared__ double red[4][8];
    const GLStatic &st = a.st;
    const int Hp = st.Hp;
    const int64_t rr = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    double acc_n = 0.0, acc_lse = 0.0, acc_sig = 0.0, acc_cnt = 0.0;
    if (rr < a.n_rows) {
        const int64_t n = a.row0 + rr;
        const double l = a.lse[n];
        bool keep = true;
        if (a.flags & GLF_USE_CUT) {
            const double cut = *a.cut;
            keep = (a.flags & GLF_CUT_STRICT) ? (l > cut) : (l >= cut);
        }
        double *scl = a.scl + n * (1 + PET_MAXHP);
        if (!keep) {                                          // truncated away: contributes nothing (bsc_et.py:254-257)
            for (int j = 0; j <= Hp; ++j) scl[j] = 0.0;
        } else {
            const int *cand = a.cand + n * Hp;
            int f = 0;
            for (int j = 0; j < Hp; ++j) {
                const int cj = cand[j];
                for (int k = j + 1; k < Hp; ++k, ++f) {
                    const double w = a.pairs[int64_t(f) * a.pairs_ld + n];
                    if (w != 0.0) {
                        const int ck = cand[k];
                        atomicAdd(&a.Wq[int64_t(cj) * st.ldH + ck], w);
                        atomicAdd(&a.Wq[int64_t(ck) * st.ldH + cj], w);
                    }
                }
            }
            const double *rs = a.rs + n * (4 + PET_MAXV);
            acc_n = 1.0; acc_lse = l; acc_sig = rs[5]; acc_cnt = rs[6];
        }
    }
    acc_n = warp_sum(acc_n); acc_lse = warp_sum(acc_lse); acc_sig = warp_sum(acc_sig); acc_cnt = warp_sum(acc_cnt);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = acc_n; red[1][warp] = acc_lse; red[2][warp] = acc_sig; red[3][warp] = acc_cnt; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        if (s != 0.0) atomicAdd(&a.scalars[threadIdx.x], s);
    }
}

int launch_gl_finalize_cut(const GLArgs &a, cudaStream_t st) {
    if (a.n_rows <= 0) return PET_OK;
    if (!a.pairs || !a.Wq || !a.scalars) { set_error("gl_finalize_cut: missing buffers"); return PET_EINVAL; }
    gl_finalize_cut_kernel<<<(unsigned)ceil_div(a.n_rows, 256), 256, 0, st>>>(a);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

int launch_gl_scale(const GLArgs &a, cudaStream_t stream) {
    if (a.n_rows <= 0 || (a.flags & (GLF_LSE_ONLY | GLF_SELECT_ONLY))) return PET_OK;
    gl_scale_kernel<<<(unsigned)ceil_div(a.n_rows * 32, 256), 256, 0, stream>>>(a);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

int launch_gl_kernel(const GLArgs &a, int gamma, bool binary, int sm_count, cudaStream_t stream) {
    PET_CHECK(launch_gl_row(a, sm_count, stream));
    PET_CHECK(launch_gl_state(a, gamma, binary, sm_count, stream));
    return launch_gl_scale(a, stream);
}

}  // namespace pet
