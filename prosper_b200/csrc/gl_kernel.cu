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
//   1 top-H' preselection            bsc_et.py:110-112 / tsc_et.py:198-210 / dsc_et.py:398-408
//   2 gather YW[cand], G[cand,cand]
//   3 log-joint of every column, running max   bsc_et.py:168-190 / tsc_et.py:337-355 / dsc_et.py:558-584
//   4 exp / sum (log-sum-exp), sigma and prior statistics   bsc_et.py:271-272,362,395-415
//   5 first and second posterior moments over the candidates (gather lists, no atomics)
//   6 <s> row for the statistics GEMM, Wq scatter, per-datapoint outputs   bsc_et.py:349-366
#include <math.h>

#include "gl_kernel.cuh"

namespace pet {

constexpr int GL_WARPS = 4;
constexpr double GL_EXP_CUTOFF = -100.0;   // exp(x) for x below this contributes < 4e-44 relative

struct WarpSmem {
    double *row;    // H   : YW row, later un-normalised <s>
    double *qbuf;   // S   : squared errors, later un-normalised posteriors of the states
    double *Gc;     // 16x16 gathered Gram block
    double *ywc;    // 16  gathered scores
    double *mom;    // n_out moment outputs
    int *cand;      // 16
    int *live;      // 16
};

static __host__ __device__ inline int r2(int x) { return (x + 1) & ~1; }

size_t gl_smem_bytes(const GLStatic &s, int warps) {
    size_t per_warp = size_t(r2(s.H)) + r2(s.S) + PET_MAXHP * PET_MAXHP + PET_MAXHP + r2(s.n_out) + PET_MAXHP;
    return (size_t(r2(s.S)) + per_warp * warps) * sizeof(double);
}

__device__ __forceinline__ double combine(const GLIter &it, double prior, double q) {
    return it.anneal_prior ? it.beta * (prior + it.pre1 * q) : prior + it.beta * (it.pre1 * q);
}

// squared error and log-prior of one multi-state record
template <int GMAX, bool BINARY>
__device__ __forceinline__ void eval_state(unsigned long long rec, const GLStatic &st, const GLIter &it,
                                           const double *ywc, const double *Gc, double yy, double &q,
                                           double &prior) {
    int pos[GMAX];
    double val[GMAX];
    double pr = double(st.zbase) * it.lp0;
#pragma unroll
    for (int m = 0; m < GMAX; ++m) {
        unsigned b = unsigned(rec >> (8 * m)) & 0xFFu;
        bool ok = (b != 0xFFu);
        pos[m] = ok ? int(b & 15u) : 0;
        int vi = ok ? int(b >> 4) : 0;
        val[m] = ok ? (BINARY ? 1.0 : st.vals[vi]) : 0.0;
        pr += ok ? (it.lp[vi] - it.lp0) : 0.0;
    }
    double acc = yy;
#pragma unroll
    for (int m = 0; m < GMAX; ++m) {
        double lin = fma(val[m], Gc[pos[m] * PET_MAXHP + pos[m]], -2.0 * ywc[pos[m]]);
        double cross = 0.0;
#pragma unroll
        for (int m2 = 0; m2 < m; ++m2) cross = fma(val[m2], Gc[pos[m] * PET_MAXHP + pos[m2]], cross);
        acc = fma(val[m], fma(2.0, cross, lin), acc);
    }
    q = acc;
    prior = pr;
}

template <int GMAX>
__device__ __forceinline__ double state_prior(unsigned long long rec, const GLStatic &st, const GLIter &it) {
    double pr = double(st.zbase) * it.lp0;
#pragma unroll
    for (int m = 0; m < GMAX; ++m) {
        unsigned b = unsigned(rec >> (8 * m)) & 0xFFu;
        pr += (b != 0xFFu) ? (it.lp[b >> 4] - it.lp0) : 0.0;
    }
    return pr;
}

// selection score of cause h (or signed/valued singleton) from the score row
__device__ __forceinline__ double sel_score(const GLArgs &a, const double *row, int h, double yy) {
    const GLStatic &st = a.st;
    switch (st.select_mode) {
        case SEL_BSC:
            return row[h] * a.invn[h];
        case SEL_NEGDIST:
            return 2.0 * row[h] - a.wn2[h];
        case SEL_GIVEN:
            return -row[h];
        default: {   // SEL_DSC: best valued singleton of this h
            double best = -INFINITY;
            for (int b = 0; b < st.n_blocks; ++b) {
                double v = st.block_val[b];
                double q = yy + v * (v * a.wn2[h] - 2.0 * row[h]);
                double f = a.it.sel_prior[b] + a.it.pre1 * q;
                best = fmax(best, f);
            }
            return best;
        }
    }
}

template <int GMAX, bool BINARY>
__global__ void __launch_bounds__(GL_WARPS * 32) gl_kernel(const GLArgs a) {
    extern __shared__ __align__(16) double smem[];
    const GLStatic &st = a.st;
    const GLIter &it = a.it;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = st.H, Hp = st.Hp, S = st.S;

    unsigned long long *states_s = reinterpret_cast<unsigned long long *>(smem);
    for (int s = threadIdx.x; s < S; s += blockDim.x) states_s[s] = st.states[s];
    const size_t per_warp = size_t(r2(H)) + r2(S) + PET_MAXHP * PET_MAXHP + PET_MAXHP + r2(st.n_out) + PET_MAXHP;
    double *wbase = smem + r2(S) + per_warp * warp;
    WarpSmem ws;
    ws.row = wbase;
    ws.qbuf = ws.row + r2(H);
    ws.Gc = ws.qbuf + r2(S);
    ws.ywc = ws.Gc + PET_MAXHP * PET_MAXHP;
    ws.mom = ws.ywc + PET_MAXHP;
    ws.cand = reinterpret_cast<int *>(ws.mom + r2(st.n_out));
    ws.live = ws.cand + PET_MAXHP;
    __syncthreads();

    const bool do_stats = !(a.flags & (GLF_LSE_ONLY | GLF_SELECT_ONLY));
    const int n_cnt = st.n_cnt;
    const int base2 = Hp * n_cnt;                 // first off-diagonal second-moment output
    const double cut = (a.flags & GLF_USE_CUT) ? *a.cut : 0.0;

    // per-warp running sums over its datapoints (lane 0 holds them)
    double acc_n = 0.0, acc_lse = 0.0, acc_sig = 0.0;
    double acc_cnt[PET_MAXV];
#pragma unroll
    for (int v = 0; v < PET_MAXV; ++v) acc_cnt[v] = 0.0;

    const int64_t wstride = int64_t(gridDim.x) * GL_WARPS;
    for (int64_t r = int64_t(blockIdx.x) * GL_WARPS + warp; r < a.n_rows; r += wstride) {
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
        for (int h = lane; h < H; h += 32) ws.row[h] = yw[h];
        __syncwarp();

        // ---- phase 1: top-H' preselection -------------------------------------------
        if (a.flags & GLF_SELECT) {
            const bool tsc = (st.select_mode == SEL_TSC);
            const int items = tsc ? 2 * H : H;
            double prev_v = INFINITY;
            int prev_i = 0x7fffffff;
            for (int rnd = 0; rnd < Hp; ++rnd) {
                double best_v = -INFINITY;
                int best_i = -1;
                for (int i = lane; i < items; i += 32) {
                    double v;
                    if (tsc) {   // item = (sign block, h): -1 block first (tsc_et.py:54-65)
                        int h = i % H;
                        double sg = (i < H) ? -1.0 : 1.0;
                        v = it.pre1 * (yy + (a.wn2[h] - 2.0 * sg * ws.row[h]));
                    } else {
                        v = sel_score(a, ws.row, i, yy);
                    }
                    bool below = (v < prev_v) || (v == prev_v && i < prev_i);
                    bool better = (v > best_v) || (v == best_v && i > best_i);
                    if (below && better) { best_v = v; best_i = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    double ov = __shfl_xor_sync(0xffffffffu, best_v, o);
                    int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                    if ((ov > best_v) || (ov == best_v && oi > best_i)) { best_v = ov; best_i = oi; }
                }
                if (best_i < 0) best_i = 0;   // all-NaN row: keep indices valid
                prev_v = best_v; prev_i = best_i;
                if (lane == 0) {
                    int h = tsc ? (best_i % H) : best_i;
                    bool ascending = (st.select_mode == SEL_BSC) || tsc;
                    ws.cand[ascending ? (Hp - 1 - rnd) : rnd] = h;
                }
            }
            __syncwarp();
            if (lane < Hp) a.cand[n * Hp + lane] = ws.cand[lane];
        } else {
            if (lane < Hp) ws.cand[lane] = a.cand[n * Hp + lane];
            __syncwarp();
        }
        if (a.flags & GLF_SELECT_ONLY) continue;

        // ---- phase 2: gather scores and Gram block of the candidates ----------------
        if (lane < Hp) {
            int c = ws.cand[lane];
            ws.ywc[lane] = ws.row[c];
            int lv = 1;
            for (int j = lane + 1; j < Hp; ++j) lv &= (ws.cand[j] != c);   // numpy "last write wins"
            ws.live[lane] = lv;
        }
        for (int idx = lane; idx < Hp * Hp; idx += 32) {
            int j = idx / Hp, k = idx % Hp;
            ws.Gc[j * PET_MAXHP + k] = a.G[int64_t(ws.cand[j]) * st.ldH + ws.cand[k]];
        }
        __syncwarp();

        double *logpj_row = a.logpj ? a.logpj + n * a.ld_logpj : nullptr;
        const bool rd = (a.flags & GLF_READ_LOGPJ) != 0, wr = (a.flags & GLF_WRITE_LOGPJ) != 0;
        const int col_states = st.has_null + st.n_blocks * H;

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
                    double q = yy + v * (v * a.wn2[h] - 2.0 * ws.row[h]);
                    F = combine(it, it.prior_block[b], q);
                    if (wr) logpj_row[st.has_null + b * H + h] = F;
                }
                mx = fmax(mx, F);
            }
        }
        for (int s = lane; s < S; s += 32) {
            double q, prior;
            eval_state<GMAX, BINARY>(states_s[s], st, it, ws.ywc, ws.Gc, yy, q, prior);
            ws.qbuf[s] = q;
            double F;
            if (rd) F = logpj_row[col_states + s];
            else {
                F = combine(it, prior, q);
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
            const double ywh = ws.row[h], wn2h = (st.n_blocks > 0) ? a.wn2[h] : 0.0;
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
                ws.row[h] = snew;
                if (a.S2) a.S2[r * st.ldH + h] = s2new;   // scaled below through a second pass
            }
        }
        for (int s = lane; s < S; s += 32) {
            double q = ws.qbuf[s];
            double F;
            if (rd) F = logpj_row[col_states + s];
            else {
                F = combine(it, state_prior<GMAX>(states_s[s], st, it), q);
            }
            double x = F - mx;
            double p = (x > GL_EXP_CUTOFF) ? exp(x) : 0.0;
            denom += p;
            sig += p * q;
            ws.qbuf[s] = p;
        }
        denom = warp_sum(denom);
        const double lse = mx + log(denom);
        if (lane == 0) a.lse[n] = lse;
        if (!do_stats) continue;
        sig = warp_sum(sig);
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v) cntb[v] = warp_sum(cntb[v]);
        const double inv = 1.0 / denom;

        // ---- phase 5: posterior moments over the candidates (gather lists) ----------
        for (int o = lane; o < st.n_out; o += 32) ws.mom[o] = 0.0;
        __syncwarp();
        {
            int cur = -1;
            double acc = 0.0;
            for (int t = 0; t < st.entries_per_lane; ++t) {
                unsigned e = st.entries[t * 32 + lane];
                if (e == 0xFFFFFFFFu) continue;
                int o = int((e >> 16) & 0xFFu);
                if (o != cur) {
                    if (cur >= 0) atomicAdd(&ws.mom[cur], acc);
                    cur = o;
                    acc = 0.0;
                }
                double p = ws.qbuf[e & 0xFFFFu];
                acc += BINARY ? p : p * st.wlut[e >> 24];
            }
            if (cur >= 0) atomicAdd(&ws.mom[cur], acc);
        }
        __syncwarp();

        // ---- phase 6: outputs --------------------------------------------------------
        // 6a. <s_h> row: singles already in row[], add candidate marginals, normalise, store
        if (st.n_blocks == 0)
            for (int h = lane; h < H; h += 32) ws.row[h] = 0.0;
        __syncwarp();
        double cnt_states[PET_MAXV];
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v) cnt_states[v] = 0.0;
        if (lane < Hp) {
            double m1 = 0.0;
#pragma unroll
            for (int v = 0; v < PET_MAXV; ++v)
                if (v < n_cnt) {
                    double P = ws.mom[lane * n_cnt + v];
                    m1 = fma(BINARY ? 1.0 : st.vals[v], P, m1);
                    cnt_states[v] = P;
                }
            if (ws.live[lane]) ws.row[ws.cand[lane]] += m1;
        }
        __syncwarp();
        for (int h = lane; h < st.ldH; h += 32) {
            a.S[r * st.ldH + h] = (h < H) ? ws.row[h] * inv : 0.0;
            if (a.S2) a.S2[r * st.ldH + h] = (h < H) ? a.S2[r * st.ldH + h] * inv : 0.0;
        }
        // 6b. second moments scattered into Wq (numpy fancy-index semantics for duplicates)
        for (int idx = lane; idx < Hp * Hp; idx += 32) {
            int j = idx / Hp, k = idx % Hp;
            if (!(ws.live[j] && ws.live[k])) continue;
            double m2;
            if (j == k) {
                if (st.diag_from_colsum) continue;
                m2 = 0.0;
                for (int v = 0; v < n_cnt; ++v) {
                    double vv = BINARY ? 1.0 : st.vals[v];
                    m2 = fma(vv * vv, ws.mom[j * n_cnt + v], m2);
                }
            } else {
                int lo = min(j, k), hi = max(j, k);
                m2 = ws.mom[base2 + lo * Hp - lo * (lo + 1) / 2 + (hi - lo - 1)];
            }
            if (m2 != 0.0) atomicAdd(&a.Wq[int64_t(ws.cand[j]) * st.ldH + ws.cand[k]], m2 * inv);
        }
        // 6c. scalar statistics
#pragma unroll
        for (int v = 0; v < PET_MAXV; ++v) cnt_states[v] = warp_sum(cnt_states[v]);
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

template <int GMAX, bool BINARY>
static int launch_inst(const GLArgs &a, int sm_count, cudaStream_t stream) {
    auto kern = gl_kernel<GMAX, BINARY>;
    size_t smem = gl_smem_bytes(a.st, GL_WARPS);
    if (smem > 227 * 1024) {
        set_error("posterior kernel needs %zu bytes of shared memory (H=%d, states=%d): unsupported size",
                  smem, a.st.H, a.st.S);
        return PET_EINVAL;
    }
    static size_t configured = 0;
    if (smem > configured) {
        PET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    int per_sm = int(std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 1024))));
    int64_t want = ceil_div(a.n_rows, GL_WARPS);
    int64_t grid = std::min<int64_t>(want, int64_t(sm_count) * per_sm);
    if (grid <= 0) return PET_OK;
    kern<<<(unsigned)grid, GL_WARPS * 32, smem, stream>>>(a);
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
