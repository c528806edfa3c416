// Posterior kernel of Gaussian Sparse Coding (spike-and-slab prior, Gaussian slab) -- gsc_et.py.
//
// For a state s with active causes a (|a| = k <= gamma) the reference builds, per cluster of
// datapoints, W_s (D x k), Lambda_s = W_s^T Sigma^-1 W_s + Psi_s^-1, C^-1 = B - B W_s Lambda_s^-1 W_s^T B
// (a D x D matrix) and evaluates  -(logdet Psi_s + logdet Lambda_s) - (y - W_s mu_s)^T C^-1 (y - W_s mu_s)
// (gsc_et.py:304-346).  With YW' = y^T Sigma^-1 W (score GEMM), G' = W^T Sigma^-1 W and yy' = y^T Sigma^-1 y:
//     A = G'[a,a],  b = YW'[a] - A mu_a,  Lambda = A + Psi_a^-1,  kappa = Lambda^-1 b + mu_a
//     quad = yy' - 2 mu_a.YW'[a] + mu_a^T A mu_a - b^T Lambda^-1 b
// i.e. only k x k algebra per (datapoint, state), no D-dimensional work.  One warp per datapoint:
//   1 singleton log-weights of ALL H causes (scalar formulas, per-cause constants from a table)
//   2 top-H' of the singleton marginal scores, sorted ascending       gsc_et.py:728-730,752-809
//   3 gather G', Psi, YW', mu, logit of the candidates
//   4 log-weights of the multi-cause states (k x k Gauss-Jordan per state), max, normaliser
//   5 weighted moments <s>, <s z>, <s s^T>, <s z z^T s^T>                gsc_et.py:349-391,524-541
//   6 rows for the statistics GEMMs + scatter of the H' x H' blocks (or dense compat output)
// Reference quirks kept: no factor 1/2 in the exponent; every weight except the null state's is
// clamped to >= tiny (gsc_et.py:355-357,520-522 vs :461-463); normaliser 1/(sum + tiny) (:564).
#include <math.h>

#include "gsc_kernel.cuh"

namespace pet {

constexpr int GSC_WARPS = 4;
constexpr int KM = PET_MAXG;                 // max active causes per state
constexpr double LOG_TINY = -708.3964185322641;   // log(np.finfo(float64).tiny)
constexpr double DBL_TINY = 2.2250738585072014e-308;

// In-place Gauss-Jordan inverse with partial pivoting of a k x k matrix (stride KM); returns log|det|.
__device__ __forceinline__ double inv_logdet(int k, double *M) {
    double Iv[KM * KM];
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) Iv[i * KM + j] = (i == j) ? 1.0 : 0.0;
    double logdet = 0.0;
    for (int c = 0; c < k; ++c) {
        int piv = c;
        double best = fabs(M[c * KM + c]);
        for (int r = c + 1; r < k; ++r)
            if (fabs(M[r * KM + c]) > best) { best = fabs(M[r * KM + c]); piv = r; }
        if (piv != c)
            for (int j = 0; j < k; ++j) {
                double t = M[c * KM + j]; M[c * KM + j] = M[piv * KM + j]; M[piv * KM + j] = t;
                t = Iv[c * KM + j]; Iv[c * KM + j] = Iv[piv * KM + j]; Iv[piv * KM + j] = t;
            }
        double d = M[c * KM + c];
        logdet += log(fabs(d));
        double inv = 1.0 / d;
        for (int j = 0; j < k; ++j) { M[c * KM + j] *= inv; Iv[c * KM + j] *= inv; }
        for (int r = 0; r < k; ++r) {
            if (r == c) continue;
            double f = M[r * KM + c];
            if (f == 0.0) continue;
            for (int j = 0; j < k; ++j) { M[r * KM + j] -= f * M[c * KM + j]; Iv[r * KM + j] -= f * Iv[c * KM + j]; }
        }
    }
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) M[i * KM + j] = Iv[i * KM + j];
    return logdet;
}

// top-Hp entries of a shared-memory score row: value descending, ties -> larger index first
__device__ __forceinline__ void select_row_max(const double *buf, int H, int Hp, int *cand) {
    const int lane = threadIdx.x & 31;
    double prev_v = INFINITY;
    int prev_i = 0x7fffffff;
    for (int rnd = 0; rnd < Hp; ++rnd) {
        double best_v = -INFINITY;
        int best_i = -1;
        for (int i = lane; i < H; i += 32) {
            double v = buf[i];
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
        if (best_i < 0) best_i = 0;
        prev_v = best_v;
        prev_i = best_i;
        if (lane == 0) cand[rnd] = best_i;
    }
}

struct StateEval {          // everything phase 5 needs of one multi-cause state
    int k;
    int pos[KM];
    double kappa[KM];
    double lam_inv[KM * KM];
    double lw;              // beta * (post + prior), clamped
    double raw;             // post + prior (compute_lpj, gsc_et.py:926)
};

__device__ __forceinline__ void eval_gsc_state(unsigned long long rec, int Hp, double beta, double yyw,
                                               const double *Gc, const double *Pc, const double *ywc,
                                               const double *muc, const double *logitc, StateEval &e, bool need_moments) {
    int k = 0;
    for (int m = 0; m < KM; ++m) {
        int p = int(unsigned(rec >> (8 * m)) & 0xFFu);
        if (p == Hp) break;
        e.pos[k++] = p;
    }
    e.k = k;
    double P[KM * KM], L[KM * KM], b[KM];
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) P[i * KM + j] = Pc[e.pos[i] * PET_MAXHP + e.pos[j]];
    double logdetP = inv_logdet(k, P);                       // Psi_s^-1, log|det Psi_s|   (gsc_et.py:319,338)
    double prior = 0.0, quad = yyw;
    for (int i = 0; i < k; ++i) {
        double Amu = 0.0;
        for (int j = 0; j < k; ++j) {
            double Aij = Gc[e.pos[i] * PET_MAXHP + e.pos[j]];
            L[i * KM + j] = Aij + P[i * KM + j];             // Lambda_s
            Amu = fma(Aij, muc[e.pos[j]], Amu);
        }
        b[i] = ywc[e.pos[i]] - Amu;                          // W_s^T Sigma^-1 (y - W_s mu_s)
        quad += muc[e.pos[i]] * (Amu - 2.0 * ywc[e.pos[i]]);
        prior += logitc[e.pos[i]];
    }
    double logdetL = inv_logdet(k, L);                       // Lambda_s^-1
    for (int i = 0; i < k; ++i) {
        double v = 0.0;
        for (int j = 0; j < k; ++j) v = fma(L[i * KM + j], b[j], v);
        e.kappa[i] = v + muc[e.pos[i]];                      // gsc_et.py:329-331
        quad -= b[i] * v;
    }
    e.raw = -(logdetP + logdetL) - quad + prior;
    double lw = beta * e.raw;                                // gsc_et.py:338-354
    if (!(lw >= LOG_TINY)) lw = LOG_TINY;                    // NaN or below tiny -> tiny (:355-357)
    e.lw = lw;
    if (need_moments)
        for (int i = 0; i < k * KM; ++i) e.lam_inv[i] = L[i];
}

__global__ void __launch_bounds__(GSC_WARPS * 32) gsc_kernel(const __grid_constant__ GSCArgs a) {
    extern __shared__ __align__(16) double smem[];
    const GLStatic &st = a.st;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = st.H, Hp = st.Hp, S = st.S;
    const int Hr = (H + 1) & ~1, Sr = (S + 1) & ~1;
    unsigned long long *states_s = reinterpret_cast<unsigned long long *>(smem);
    for (int s = threadIdx.x; s < S; s += blockDim.x) states_s[s] = st.states[s];
    const int per_warp = 3 * Hr + Sr + 4 * PET_MAXHP * PET_MAXHP + 8 * PET_MAXHP;
    double *wb = smem + Sr + size_t(per_warp) * warp;
    double *row = wb;                              // H   YW' row
    double *buf = row + Hr;                        // H   scores, then singleton log-weights, then <s> row
    double *szr = buf + Hr;                        // H   <s z> row
    double *lwm = szr + Hr;                        // S   log-weights of the multi-cause states
    double *Gc = lwm + Sr;                         // 16x16 blocks
    double *Pc = Gc + PET_MAXHP * PET_MAXHP;
    double *Mss = Pc + PET_MAXHP * PET_MAXHP;
    double *Mzz = Mss + PET_MAXHP * PET_MAXHP;
    double *ywc = Mzz + PET_MAXHP * PET_MAXHP;     // 16 each
    double *muc = ywc + PET_MAXHP;
    double *lgc = muc + PET_MAXHP;
    double *ms = lgc + PET_MAXHP;                  // candidate marginals
    double *msz = ms + PET_MAXHP;
    double *sz2c = msz + PET_MAXHP;
    int *cand_s = reinterpret_cast<int *>(sz2c + PET_MAXHP);   // 16 ints (+16 spare)
    __syncthreads();

    const int64_t wstride = int64_t(gridDim.x) * GSC_WARPS;
    for (int64_t r = int64_t(blockIdx.x) * GSC_WARPS + warp; r < a.n_rows; r += wstride) {
        const int64_t n = a.row0 + r;
        const double yyw = a.yyw[n];
        const double *yw = a.YW + r * st.ldH;
        // ---- 1: singleton scores of all causes ------------------------------------------------
        for (int h = lane; h < H; h += 32) {
            const double v = yw[h];
            row[h] = v;
            const double g = a.tb.g[h], mu = a.tb.mu[h];
            const double b = v - g * mu;
            const double quad = yyw + mu * (mu * g - 2.0 * v) - b * b * a.tb.ilam[h];
            double post = -a.tb.lcdet[h] - quad;             // gsc_et.py:800-803
            if (post != post || post < -1.7976931348623157e308) post = -1.7976931348623157e308;
            else if (isinf(post)) post = 0.0;                // :805-807
            buf[h] = post;
        }
        __syncwarp();
        // ---- 2: top-H' causes, ascending cause index (gsc_et.py:729-730) ------------------------
        if (a.flags & GSCF_SELECT) {
            select_row_max(buf, H, Hp, cand_s);
            __syncwarp();
            int mine = (lane < Hp) ? cand_s[lane] : 0x7fffffff;
            int rank = 0;
            for (int j = 0; j < Hp; ++j) rank += (cand_s[j] < mine) ? 1 : 0;
            __syncwarp();
            if (lane < Hp) cand_s[rank] = mine;
            __syncwarp();
            if (lane < Hp) a.cand[n * Hp + lane] = cand_s[lane];
        } else {
            if (lane < Hp) cand_s[lane] = a.cand[n * Hp + lane];
            __syncwarp();
        }
        if (a.flags & GSCF_SELECT_ONLY) continue;

        // ---- 3: gather ---------------------------------------------------------------------------
        for (int idx = lane; idx < Hp * Hp; idx += 32) {
            int j = idx / Hp, k = idx % Hp;
            Gc[j * PET_MAXHP + k] = a.G[int64_t(cand_s[j]) * st.ldH + cand_s[k]];
            Pc[j * PET_MAXHP + k] = a.psi[int64_t(cand_s[j]) * st.ldH + cand_s[k]];
            Mss[j * PET_MAXHP + k] = 0.0;
            Mzz[j * PET_MAXHP + k] = 0.0;
        }
        if (lane < Hp) {
            int c = cand_s[lane];
            ywc[lane] = row[c]; muc[lane] = a.tb.mu[c]; lgc[lane] = a.tb.logit[c];
            ms[lane] = 0.0; msz[lane] = 0.0; sz2c[lane] = 0.0;
        }
        __syncwarp();

        if (a.flags & GSCF_LOGPJ) {                              // compute_lpj: gsc_et.py:864 (null), :893 (singletons), :926
            double *out = a.logpj + n * a.ld_logpj;
            if (lane == 0) out[0] = -yyw;
            for (int h = lane; h < H; h += 32) {
                const double v = row[h];
                const double g = a.tb.g[h], mu = a.tb.mu[h];
                const double b = v - g * mu;
                const double quad = yyw + mu * (mu * g - 2.0 * v) - b * b * a.tb.ilam[h];
                out[1 + h] = (-a.tb.lcdet[h] - quad) + a.tb.logit[h];
            }
            for (int s = lane; s < S; s += 32) {
                StateEval e;
                eval_gsc_state(states_s[s], Hp, 1.0, yyw, Gc, Pc, ywc, muc, lgc, e, false);
                out[1 + H + s] = e.raw;
            }
            __syncwarp();
            continue;
        }
        // ---- 4: log-weights -----------------------------------------------------------------------
        const double lw0 = -a.beta * yyw;                    // null state, NOT clamped (gsc_et.py:459-463)
        double mx = lw0;
        for (int h = lane; h < H; h += 32) {
            const double v = row[h];
            const double g = a.tb.g[h], mu = a.tb.mu[h];
            const double b = v - g * mu;
            const double quad = yyw + mu * (mu * g - 2.0 * v) - b * b * a.tb.ilam[h];
            double lw = a.beta * (-a.tb.lcdet[h] - quad + a.tb.logit[h]);     // gsc_et.py:510-518
            if (!(lw >= LOG_TINY)) lw = LOG_TINY;
            buf[h] = lw;
            mx = fmax(mx, lw);
        }
        for (int s = lane; s < S; s += 32) {
            StateEval e;
            eval_gsc_state(states_s[s], Hp, a.beta, yyw, Gc, Pc, ywc, muc, lgc, e, false);
            lwm[s] = e.lw;
            mx = fmax(mx, e.lw);
        }
        mx = warp_max(mx);
        double Z = (lane == 0) ? exp(lw0 - mx) : 0.0;
        for (int h = lane; h < H; h += 32) Z += exp(buf[h] - mx);
        for (int s = lane; s < S; s += 32) Z += exp(lwm[s] - mx);
        Z = warp_sum(Z);
        const double nf = 1.0 / (Z + DBL_TINY * exp(-mx));    // 1/(sum of weights + tiny), shifted (gsc_et.py:564)

        // ---- 5: weighted moments -------------------------------------------------------------------
        for (int h = lane; h < H; h += 32) {                 // singletons: dense rows
            const double w = exp(buf[h] - mx) * nf;
            const double g = a.tb.g[h], mu = a.tb.mu[h];
            const double kap = (row[h] - g * mu) * a.tb.ilam[h] + mu;         // gsc_et.py:495-497
            buf[h] = w;                                                       // <s_h>
            szr[h] = w * kap;                                                 // <s_h z_h>
            if (a.SZ2) a.SZ2[r * st.ldH + h] = w * (kap * kap + a.tb.ilam[h]);   // :500
            if (a.flags & GSCF_DENSE) row[h] = w * (kap * kap + a.tb.ilam[h]);
        }
        __syncwarp();
        for (int s = lane; s < S; s += 32) {                 // multi-cause states: candidate blocks
            const double w = exp(lwm[s] - mx) * nf;
            if (w == 0.0) continue;
            StateEval e;
            eval_gsc_state(states_s[s], Hp, a.beta, yyw, Gc, Pc, ywc, muc, lgc, e, true);
            for (int i = 0; i < e.k; ++i) {
                atomicAdd(&ms[e.pos[i]], w);
                atomicAdd(&msz[e.pos[i]], w * e.kappa[i]);
                for (int j = 0; j < e.k; ++j) {
                    atomicAdd(&Mss[e.pos[i] * PET_MAXHP + e.pos[j]], w);
                    atomicAdd(&Mzz[e.pos[i] * PET_MAXHP + e.pos[j]],
                              w * (e.kappa[i] * e.kappa[j] + e.lam_inv[i * KM + j]));   // gsc_et.py:334,361
                }
            }
        }
        __syncwarp();

        // ---- 6: outputs ----------------------------------------------------------------------------
        if (lane < Hp) {
            int c = cand_s[lane];
            buf[c] += ms[lane];
            szr[c] += msz[lane];
        }
        __syncwarp();
        if (a.flags & GSCF_STATS) {
            for (int h = lane; h < st.ldH; h += 32) {
                a.XS[r * st.ldH + h] = (h < H) ? buf[h] : 0.0;
                a.XSZ[r * st.ldH + h] = (h < H) ? szr[h] : 0.0;
                if (h >= H) a.SZ2[r * st.ldH + h] = 0.0;
            }
            for (int idx = lane; idx < Hp * Hp; idx += 32) {
                int j = idx / Hp, k = idx % Hp;
                double vss = Mss[j * PET_MAXHP + k], vzz = Mzz[j * PET_MAXHP + k];
                // diag(<s s^T>) = <s>: the diagonal of sum_ss is taken from the column sums of XS
                if (j != k && vss != 0.0) atomicAdd(&a.sum_ss[int64_t(cand_s[j]) * st.ldH + cand_s[k]], vss);
                if (vzz != 0.0) atomicAdd(&a.sum_szsz[int64_t(cand_s[j]) * st.ldH + cand_s[k]], vzz);
            }
        }
        if (a.flags & GSCF_DENSE) {
            const int64_t o = a.dst ? a.dst[n] : n;
            double *os = a.xpt_s + o * H, *osz = a.xpt_sz + o * H;
            double *oss = a.xpt_ss + o * H * H, *ozz = a.xpt_szsz + o * H * H;
            for (int h = lane; h < H; h += 32) { os[h] = buf[h]; osz[h] = szr[h]; }
            for (int i = lane; i < H * H; i += 32) { oss[i] = 0.0; ozz[i] = 0.0; }
            __syncwarp();
            // singleton diagonals: <s s> = w_h, <sz sz> = w_h (kappa^2 + 1/Lambda); buf holds singles+marginals,
            // so remove the candidate marginals again for the pure singleton part
            for (int h = lane; h < H; h += 32) {
                double sing = buf[h];
                for (int j = 0; j < Hp; ++j) if (cand_s[j] == h) sing -= ms[j];
                oss[int64_t(h) * H + h] = sing;
                ozz[int64_t(h) * H + h] = row[h];
            }
            __syncwarp();
            for (int idx = lane; idx < Hp * Hp; idx += 32) {
                int j = idx / Hp, k = idx % Hp;
                oss[int64_t(cand_s[j]) * H + cand_s[k]] += Mss[j * PET_MAXHP + k];
                ozz[int64_t(cand_s[j]) * H + cand_s[k]] += Mzz[j * PET_MAXHP + k];
            }
        }
        __syncwarp();
    }
}

size_t gsc_smem_bytes(const GLStatic &s) {
    const int Hr = (s.H + 1) & ~1, Sr = (s.S + 1) & ~1;
    const size_t per_warp = 3 * Hr + Sr + 4 * PET_MAXHP * PET_MAXHP + 8 * PET_MAXHP;
    return (Sr + per_warp * GSC_WARPS) * sizeof(double);
}

int launch_gsc_kernel(const GSCArgs &a, int gamma, int sm_count, cudaStream_t stream) {
    (void)gamma;
    size_t smem = gsc_smem_bytes(a.st);
    if (smem > 227 * 1024) {
        set_error("GSC kernel needs %zu bytes of shared memory (H=%d): unsupported size", smem, a.st.H);
        return PET_EINVAL;
    }
    static size_t configured = 0;
    if (smem > configured) {
        PET_CUDA(cudaFuncSetAttribute(gsc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    int per_sm = int(std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 1024))));
    int64_t grid = std::min<int64_t>(ceil_div(a.n_rows, GSC_WARPS), int64_t(sm_count) * per_sm);
    if (grid <= 0) return PET_OK;
    gsc_kernel<<<(unsigned)grid, GSC_WARPS * 32, smem, stream>>>(a);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// ---- per-iteration tables ------------------------------------------------------------------------
__global__ void gsc_tables_kernel(const double *G, int64_t ldG, const double *psi, int64_t ldpsi, const double *pi,
                                  const double *mu, int H, double *g, double *ilam, double *lcdet, double *logit) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    const double gh = G[int64_t(h) * ldG + h], ps = psi[int64_t(h) * ldpsi + h];
    const double lam = gh + 1.0 / ps;                        // gsc_et.py:486
    g[h] = gh;
    ilam[h] = 1.0 / lam;
    lcdet[h] = log(ps) + log(fabs(lam));                     // :502  (slogdet of a 1x1)
    logit[h] = log(pi[h]) - log(1.0 - pi[h]);                // :457
    (void)mu;
}

int launch_gsc_tables(const double *G, int64_t ldG, const double *psi, int64_t ldpsi, const double *pi, const double *mu,
                      int H, double *g, double *ilam, double *lcdet, double *logit, cudaStream_t st) {
    gsc_tables_kernel<<<(unsigned)ceil_div(H, 128), 128, 0, st>>>(G, ldG, psi, ldpsi, pi, mu, H, g, ilam, lcdet, logit);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// Wt_out[h][d] = Wt[h][d] * B_dd   (rows of W^T scaled by the inverse noise variance)
__global__ void scale_rows_kernel(double *out, const double *in, int64_t ld, int H, int D, const double *bdiag, double bscalar) {
    int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= int64_t(H) * ld) return;
    int d = int(idx % ld);
    out[idx] = (d < D) ? in[idx] * (bdiag ? bdiag[d] : bscalar) : 0.0;
}
int launch_scale_rows(double *out, const double *in, int64_t ld, int H, int D, const double *bdiag, double bscalar, cudaStream_t st) {
    scale_rows_kernel<<<(unsigned)ceil_div(int64_t(H) * ld, 256), 256, 0, st>>>(out, in, ld, H, D, bdiag, bscalar);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// out[n] = sum_d y_nd^2 B_dd
__global__ void weighted_rownorm_kernel(const double *Y, int64_t ldy, int64_t n, int D, const double *bdiag, double bscalar, double *out) {
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *y = Y + row * ldy;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) s = fma(y[d] * (bdiag ? bdiag[d] : bscalar), y[d], s);
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}
int launch_weighted_rownorm(const double *Y, int64_t ldy, int64_t n, int D, const double *bdiag, double bscalar, double *out, cudaStream_t st) {
    if (n <= 0) return PET_OK;
    weighted_rownorm_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(Y, ldy, n, D, bdiag, bscalar, out);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

}  // namespace pet
