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

constexpr int GL_MAX_WARPS = 8;
constexpr double GL_EXP_CUTOFF = -100.0;   // exp(x) for x below this contributes < 4e-44 relative

static __host__ __device__ inline int r2(int x) { return (x + 1) & ~1; }

constexpr int GS = PET_MAXHP + 1;          // stride of the gathered Gram block; row/col Hp is all zero

struct SmemLayout {
    int shared_doubles;   // state records + gather table + chunk table
    int off_ids, off_chunk;
    int per_warp;         // doubles
    int off_q, off_G, off_lin, off_mom, off_P, off_cand;
};

static __host__ __device__ inline SmemLayout smem_layout(const GLStatic &s) {
    SmemLayout L;
    const int nch = s.n_chunks > 0 ? s.n_chunks : 1;
    L.off_ids = r2(s.S);
    L.off_chunk = L.off_ids + nch * s.chunk_len * 8;      // nch*CH*32 u16
    L.shared_doubles = L.off_chunk + nch * 8;             // nch*32 u16
    L.off_q = r2(s.H);
    L.off_G = L.off_q + r2(s.S + 1);
    L.off_lin = L.off_G + r2(GS * GS);
    L.off_mom = L.off_lin + r2(GS);
    L.off_P = L.off_mom + r2(s.n_out + 1);
    L.off_cand = L.off_P + r2(s.Hp * (s.n_cnt > 0 ? s.n_cnt : 1));
    L.per_warp = L.off_cand + PET_MAXHP;     // cand[16] + live[16] ints
    return L;
}

size_t gl_smem_bytes(const GLStatic &s, int warps) {
    SmemLayout L = smem_layout(s);
    return (size_t(L.shared_doubles) + size_t(L.per_warp) * warps) * sizeof(double);
}

__constant__ double c_winv[8] = {1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8};

// log-prior of multi-state s.  Binary spaces are enumerated by size (camodels/__init__.py:30-33), so
// |s| follows from the index and nothing is loaded; valued spaces read the per-iteration table.
template <bool BINARY>
__device__ __forceinline__ double prior_of(const GLArgs &a, int s) {
    if (BINARY) {
        int n = 2;
#pragma unroll
        for (int g = 3; g <= PET_MAXG; ++g) n += (s >= a.st.size_start[g]) ? 1 : 0;
        return a.it.lp[0] * double(n);
    }
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
            return row[i] * a.invn[i];
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

template <int GMAX, bool BINARY>
__global__ void __launch_bounds__(GL_MAX_WARPS * 32) gl_kernel(const __grid_constant__ GLArgs a) {
    extern __shared__ __align__(16) double smem[];
    const GLStatic &st = a.st;
    const GLIter &it = a.it;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int H = st.H, Hp = st.Hp, S = st.S;
    const SmemLayout L = smem_layout(st);
    const int CH = st.chunk_len, NCH = st.n_chunks;

    unsigned long long *states_s = reinterpret_cast<unsigned long long *>(smem);
    unsigned short *ids_s = reinterpret_cast<unsigned short *>(smem + L.off_ids);
    unsigned short *chunk_s = reinterpret_cast<unsigned short *>(smem + L.off_chunk);
    for (int s = threadIdx.x; s < S; s += blockDim.x) states_s[s] = st.states[s];
    for (int i = threadIdx.x; i < NCH * CH * 32; i += blockDim.x) ids_s[i] = st.entries[i];
    for (int i = threadIdx.x; i < NCH * 32; i += blockDim.x) chunk_s[i] = st.chunk_tab[i];
    double *wbase = smem + L.shared_doubles + size_t(L.per_warp) * warp;
    double *row = wbase;
    double *qbuf = wbase + L.off_q;
    double *Gc = wbase + L.off_G;
    double *lin = wbase + L.off_lin;
    double *mom = wbase + L.off_mom;
    double *Pj = wbase + L.off_P;
    int *cand_s = reinterpret_cast<int *>(wbase + L.off_cand);
    int *live_s = cand_s + PET_MAXHP;
    // zero row/column Hp of the Gram block and lin[Hp] once: the dummy position of unused member slots
    for (int i = lane; i < GS * GS; i += 32) Gc[i] = 0.0;
    for (int i = lane; i < GS; i += 32) lin[i] = 0.0;
    __syncthreads();

    const bool do_stats = !(a.flags & (GLF_LSE_ONLY | GLF_SELECT_ONLY));
    const int n_cnt = st.n_cnt, n_g = st.n_g;
    const double cut = (a.flags & GLF_USE_CUT) ? *a.cut : 0.0;
    const bool rd = (a.flags & GLF_READ_LOGPJ) != 0, wr = (a.flags & GLF_WRITE_LOGPJ) != 0;
    const int col_states = st.has_null + st.n_blocks * H;
    const int items = (st.select_mode == SEL_TSC) ? 2 * H : H;

    // per-warp running sums over its datapoints (lane 0 holds them)
    double acc_n = 0.0, acc_lse = 0.0, acc_sig = 0.0;
    double acc_cnt[PET_MAXV];
#pragma unroll
    for (int v = 0; v < PET_MAXV; ++v) acc_cnt[v] = 0.0;

    const int64_t wstride = int64_t(gridDim.x) * nwarps;
    for (int64_t r = int64_t(blockIdx.x) * nwarps + warp; r < a.n_rows; r += wstride) {
        const int64_t n = a.row0 + r;
        const double *yw = a.YW + r * st.ldH;
        const double yy = a.yy[n];

        if ((a.flags & GLF_USE_CUT) && do_stats) {
            double l = a.lse[n];
            bool keep = (a.flags & GLF_CUT_STRICT) ? (l > cut) : (l >= cut);
            if (!keep) {   // truncated away: contributes nothing (bsc_et.py:254-257)
                for (int h = lane; h < st.ldH; h += 32) {
                    a.S[r * st.ldH + h] = 0.0;
                    if (a.S2) a.S2[r * st.ldH + h] = 0.0;
                }
                continue;
            }
        }

        // ---- phase 0: score row into shared memory ---------------------------------
        for (int h = lane; h < H; h += 32) row[h] = yw[h];
        __syncwarp();

        // ---- phase 1: top-H' preselection -------------------------------------------
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
        if (a.flags & GLF_SELECT_ONLY) continue;

        // ---- phase 2: gather scores and Gram block of the candidates ----------------
        for (int idx = lane; idx < Hp * Hp; idx += 32) {
            int j = idx / Hp, k = idx % Hp;
            Gc[j * GS + k] = a.G[int64_t(cand_s[j]) * st.ldH + cand_s[k]];
        }
        if (lane < Hp) {
            int c = cand_s[lane];
            int lv = 1;
            for (int j = lane + 1; j < Hp; ++j) lv &= (cand_s[j] != c);   // numpy "last write wins"
            live_s[lane] = lv;
        }
        __syncwarp();
        if (lane < Hp) {
            double ywc = row[cand_s[lane]];
            lin[lane] = BINARY ? fma(-2.0, ywc, Gc[lane * GS + lane]) : ywc;
        }
        __syncwarp();

        double *logpj_row = a.logpj ? a.logpj + n * a.ld_logpj : nullptr;

        // ---- phase 3: log-joints, running max ---------------------------------------
        double mx = -INFINITY;
        double F0 = 0.0;
        if (st.has_null) {
            F0 = rd ? logpj_row[0] : combine(it, it.prior_null, yy);
            if (wr && lane == 0) logpj_row[0] = F0;
            mx = F0;
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
                mx = fmax(mx, F);
            }
        }
#pragma unroll 2
        for (int s = lane; s < S; s += 32) {
            double q = eval_state<GMAX, BINARY>(states_s[s], st, lin, Gc, yy);
            qbuf[s] = q;
            double F;
            if (rd) F = logpj_row[col_states + s];
            else {
                F = combine(it, prior_of<BINARY>(a, s), q);
                if (wr) logpj_row[col_states + s] = F;
            }
            mx = fmax(mx, F);
        }
        mx = warp_max(mx);
        if (wr && (a.flags & GLF_LSE_ONLY)) continue;   // compat E_step: logpj only

        // ---- phase 4: exp, denominators, scalar statistics --------------------------
        double denom = 0.0, sig = 0.0;
        double cntb[PET_MAXV];
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v) cntb[v] = 0.0;
        if (st.has_null && lane == 0) {
            double x = F0 - mx;
            double p = (x > GL_EXP_CUTOFF) ? exp(x) : 0.0;
            denom += p;
            sig += p * yy;
        }
        for (int h = lane; h < H; h += 32) {
            double snew = 0.0, s2new = 0.0;
            const double ywh = row[h], wn2h = (st.n_blocks > 0) ? a.wn2[h] : 0.0;
#pragma unroll
            for (int b = 0; b < PET_MAXV; ++b) {
                if (b < st.n_blocks) {
                    const double v = st.block_val[b];
                    double q = yy + v * (v * wn2h - 2.0 * ywh);
                    double F = rd ? logpj_row[st.has_null + b * H + h] : combine(it, it.prior_block[b], q);
                    double x = F - mx;
                    double p = (x > GL_EXP_CUTOFF) ? exp(x) : 0.0;
                    denom += p;
                    sig += p * q;
                    cntb[b] += p;
                    snew = fma(p, v, snew);
                    s2new = fma(p, v * v, s2new);
                }
            }
            if (do_stats) {
                row[h] = snew;
                if (a.S2) a.S2[r * st.ldH + h] = s2new;   // normalised in phase 6
            }
        }
#pragma unroll 2
        for (int s = lane; s < S; s += 32) {
            double q = qbuf[s];
            double F = rd ? logpj_row[col_states + s] : combine(it, prior_of<BINARY>(a, s), q);
            double x = F - mx;
            double p = (x > GL_EXP_CUTOFF) ? exp(x) : 0.0;
            denom += p;
            sig += p * q;
            qbuf[s] = p;
        }
        if (lane == 0) qbuf[S] = 0.0;   // zero slot read by padding entries of the gather table
        denom = warp_sum(denom);
        const double lse = mx + log(denom);
        if (lane == 0) a.lse[n] = lse;
        if (!do_stats) continue;
        sig = warp_sum(sig);
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v)
            if (v < st.n_blocks) cntb[v] = warp_sum(cntb[v]);
        const double inv = 1.0 / denom;

        // ---- phase 5: pair sums from the shared-memory gather table -----------------
        // every lane runs NCH chunks of CH ids; all chunks of an output belong to one lane
        for (int o = lane; o <= st.n_out; o += 32) mom[o] = 0.0;
        __syncwarp();
        for (int i = lane; i < st.n_direct; i += 32) {
            unsigned d = st.direct[i];
            mom[d >> 16] = qbuf[d & 0xFFFFu];
        }
        {
            double acc = 0.0;
            const unsigned short *ids = ids_s + lane;
            for (int c = 0; c < NCH; ++c) {
                const unsigned co = chunk_s[c * 32 + lane];
                double s0 = 0.0, s1 = 0.0;
                for (int i = 0; i < CH; i += 4) {
                    const unsigned short *e = ids + (c * CH + i) * 32;
                    s0 += qbuf[e[0]] + qbuf[e[32]];
                    s1 += qbuf[e[64]] + qbuf[e[96]];
                }
                acc += s0 + s1;
                if (co & 0x8000u) { mom[co & 0x7FFFu] = acc; acc = 0.0; }
            }
        }
        __syncwarp();

        // ---- phase 6: moments and outputs --------------------------------------------
        // 6a. P[j][a] = posterior mass of (s_j = v_a):  singleton state + size-weighted pair sums
        for (int idx = lane; idx < Hp * n_cnt; idx += 32) {
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
        if (st.n_blocks == 0)
            for (int h = lane; h < H; h += 32) row[h] = 0.0;
        __syncwarp();
        // 6b. <s_h> row: singles already in row[], add candidate marginals, normalise, store
        double cnt_states[PET_MAXV];
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v) cnt_states[v] = 0.0;
        if (lane < Hp) {
            double m1 = 0.0;
#pragma unroll
            for (int v = 0; v < PET_MAXV; ++v)
                if (v < n_cnt) {
                    double P = Pj[lane * n_cnt + v];
                    m1 = fma(BINARY ? 1.0 : st.vals[v], P, m1);
                    cnt_states[v] = P;
                }
            if (live_s[lane]) row[cand_s[lane]] += m1;
        }
        __syncwarp();
        for (int h = lane; h < st.ldH; h += 32) {
            a.S[r * st.ldH + h] = (h < H) ? row[h] * inv : 0.0;
            if (a.S2) a.S2[r * st.ldH + h] = (h < H) ? a.S2[r * st.ldH + h] * inv : 0.0;
        }
        // 6c. second moments scattered into Wq (numpy fancy-index semantics for duplicates)
        for (int idx = lane; idx < Hp * Hp; idx += 32) {
            const int j = idx / Hp, k = idx % Hp;
            if (!(live_s[j] && live_s[k])) continue;
            double m2 = 0.0;
            if (j == k) {
                if (st.diag_from_colsum) continue;
                for (int v = 0; v < n_cnt; ++v) {
                    double vv = BINARY ? 1.0 : st.vals[v];
                    m2 = fma(vv * vv, Pj[j * n_cnt + v], m2);
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
                        m2 = fma(w, sacc, m2);
                    }
            }
            if (m2 != 0.0) atomicAdd(&a.Wq[int64_t(cand_s[j]) * st.ldH + cand_s[k]], m2 * inv);
        }
        // 6d. scalar statistics
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v)
            if (v < n_cnt) cnt_states[v] = warp_sum(cnt_states[v]);
        if (lane == 0) {
            acc_n += 1.0;
            acc_lse += lse;
            acc_sig += sig * inv;
#pragma unroll
            for (int v = 0; v < PET_MAXV; ++v) {
                double c = cnt_states[v];
#pragma unroll
                for (int b = 0; b < PET_MAXV; ++b)
                    if (b < st.n_blocks && st.block_vidx[b] == v) c += cntb[b];
                acc_cnt[v] += c * inv;
            }
        }
        __syncwarp();
    }

    if (do_stats && lane == 0) {
        atomicAdd(&a.scalars[0], acc_n);
        atomicAdd(&a.scalars[1], acc_lse);
        atomicAdd(&a.scalars[2], acc_sig);
        for (int v = 0; v < st.n_cnt; ++v) atomicAdd(&a.scalars[3 + v], acc_cnt[v]);
    }
}

// largest warp count whose shared memory fits one SM (0 = does not fit at all)
int gl_pick_warps(const GLStatic &s) {
    static int cap = []() { const char *e = getenv("PET_GL_WARPS"); int v = e ? atoi(e) : GL_MAX_WARPS; return v < 1 ? 1 : (v > GL_MAX_WARPS ? GL_MAX_WARPS : v); }();
    for (int w = cap; w >= 1; --w)
        if (gl_smem_bytes(s, w) <= 227 * 1024) return w;
    return 0;
}

template <int GMAX, bool BINARY>
static int launch_inst(const GLArgs &a, int sm_count, cudaStream_t stream) {
    auto kern = gl_kernel<GMAX, BINARY>;
    const int warps = gl_pick_warps(a.st);
    if (warps == 0) {
        set_error("posterior kernel needs %zu bytes of shared memory per warp set (H=%d, states=%d): unsupported size",
                  gl_smem_bytes(a.st, 1), a.st.H, a.st.S);
        return PET_EINVAL;
    }
    const int use_warps = warps;
    const size_t smem = gl_smem_bytes(a.st, use_warps);
    static size_t configured = 0;
    if (smem > configured) {
        PET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    int per_sm = int(std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 1024))));
    per_sm = std::min(per_sm, std::max(1, 64 / use_warps));
    int64_t want = ceil_div(a.n_rows, use_warps);
    int64_t grid = std::min<int64_t>(want, int64_t(sm_count) * per_sm);
    if (grid <= 0) return PET_OK;
    kern<<<(unsigned)grid, use_warps * 32, smem, stream>>>(a);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

int launch_gl_kernel(const GLArgs &a, int gamma, bool binary, int sm_count, cudaStream_t stream) {
    if (binary) {
        if (gamma <= 3) return launch_inst<3, true>(a, sm_count, stream);
        if (gamma <= 5) return launch_inst<5, true>(a, sm_count, stream);
        return launch_inst<8, true>(a, sm_count, stream);
    }
    if (gamma <= 4) return launch_inst<4, false>(a, sm_count, stream);
    return launch_inst<8, false>(a, sm_count, stream);
}

}  // namespace pet
