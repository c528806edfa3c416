// Posterior / statistics kernel of the max-superposition models MCA-ET and MMCA-ET.
//
// These models combine the active causes with a rho-norm (soft max),
//     MCA :  Wbar_d = ( sum_{h in s} W_hd^rho )^(1/rho)                       mca_et.py:171-173
//     MMCA:  Wbar_d = sign(t) |t|^(1/rho),  t = sum_{h in s} sign(W_hd)|W_hd|^rho   mmca_et.py:191-192
// so ||y - Wbar_s||^2 is NOT a quadratic form in the state and there is no Gram shortcut: every
// (datapoint, state) needs a D-loop with a log/exp pair.  The H' candidate rows of W^rho and log|W|
// are staged in shared memory once per datapoint; one CTA handles one datapoint at a time.
// The all-H singleton block still comes from the score GEMM (||W_h - y||^2 = wn2 - 2 yW + yy).
//
// Phases per datapoint:
//   1 stage y, the score row, W^rho[cand] and log|W|[cand]
//   2 log-joints of null / singles / multi-states              mca_et.py:156-175, mmca_et.py:175-194
//   3 annealed posterior exp(beta*logpj - corr), log-denominators  mca_et.py:237-238,250, :327
//   4 statistics: singles posterior row for the statistics GEMM, Aid scatter, pi / sigma sums
//                                                               mca_et.py:287-325, mmca_et.py:312-352
#include <math.h>

#include "mca_kernel.cuh"

namespace pet {

constexpr int MCA_THREADS = 128;

__device__ __forceinline__ double block_reduce(double v, double *red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) r = is_max ? fmax(r, red[w]) : r + red[w];
    return r;
}

// Wlbar and Wbar of state `rec` at feature d from the staged candidate rows (row Hp is all zero)
__device__ __forceinline__ void wbar_at(const MCAArgs &a, unsigned long long rec, const double *Wr_c, int d,
                                        double &Wlbar, double &Wbar) {
    double t0 = 0.0;
    for (int m = 0; m < a.gamma; ++m) t0 += Wr_c[int(unsigned(rec >> (8 * m)) & 0xFFu) * a.ldc + d];
    if (a.mmca) {
        Wlbar = log(fabs(t0)) / a.rho;                       // -inf for t0 == 0 -> Wbar = 0
        double mag = exp(Wlbar);
        Wbar = (t0 > 0.0) ? mag : ((t0 < 0.0) ? -mag : 0.0);
    } else {
        Wlbar = log(t0) / a.rho;
        Wbar = exp(Wlbar);
    }
}

__global__ void __launch_bounds__(MCA_THREADS) mca_kernel(const __grid_constant__ MCAArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int D = a.D, H = a.H, Hp = a.Hp, S = a.S, ldc = a.ldc;
    double *y_s = smem;                               // D
    double *row = y_s + ((D + 1) & ~1);               // H      score row  y.W
    double *Wr_c = row + ((H + 1) & ~1);              // (Hp+1) x ldc   W^rho of the candidates (+ zero row)
    double *Wl_c = Wr_c + (Hp + 1) * ldc;             // Hp x ldc       log|W| of the candidates
    double *Fb = Wl_c + Hp * ldc;                     // C      log-joints, later annealed posteriors
    double *qb = Fb + ((a.C + 1) & ~1);               // S      squared errors of the multi-states
    double *Aid = qb + ((S + 1) & ~1);                // Hp x ldc
    double *red = Aid + Hp * ldc;                     // 32
    int *cand_s = reinterpret_cast<int *>(red + 32);  // Hp

    const int tid = threadIdx.x;
    const bool rd = (a.flags & GLF_READ_LOGPJ) != 0, wr = (a.flags & GLF_WRITE_LOGPJ) != 0;
    const bool lse_only = (a.flags & GLF_LSE_ONLY) != 0;
    const double cut = (a.flags & GLF_USE_CUT) ? *a.cut : 0.0;
    for (int i = tid; i < ldc; i += blockDim.x) Wr_c[Hp * ldc + i] = 0.0;

    double acc_n = 0.0, acc_ld = 0.0, acc_sig = 0.0, acc_pi = 0.0;   // thread 0 accumulates

    for (int64_t r = blockIdx.x; r < a.n_rows; r += gridDim.x) {
        const int64_t n = a.row0 + r;
        __syncthreads();
        if ((a.flags & GLF_USE_CUT) && !lse_only) {
            if (!(a.lse[n] >= cut)) {            // truncated away (mca_et.py:254-261)
                for (int h = tid; h < a.ldH; h += blockDim.x) a.Spost[r * a.ldH + h] = 0.0;
                continue;
            }
        }
        // ---- phase 1: stage ------------------------------------------------------------------
        const double yy = a.yy[n];
        for (int d = tid; d < D; d += blockDim.x) y_s[d] = a.Y[r * a.ldY + d];
        for (int h = tid; h < H; h += blockDim.x) row[h] = a.YW[r * a.ldH + h];
        if (tid < Hp) cand_s[tid] = a.cand[n * Hp + tid];
        __syncthreads();
        for (int i = tid; i < Hp * D; i += blockDim.x) {
            int j = i / D, d = i % D;
            Wr_c[j * ldc + d] = a.Wr[int64_t(cand_s[j]) * a.ldD + d];
            Wl_c[j * ldc + d] = a.Wl[int64_t(cand_s[j]) * a.ldD + d];
        }
        __syncthreads();
        double *logpj_row = a.logpj ? a.logpj + n * a.ld_logpj : nullptr;

        // ---- phase 2: log-joints (NOT annealed: mca_et.py:160-175) ----------------------------
        if (tid == 0) Fb[0] = rd ? logpj_row[0] : a.pre1 * yy;
        for (int h = tid; h < H; h += blockDim.x) {
            double q = yy - 2.0 * row[h] + a.wn2[h];
            Fb[1 + h] = rd ? logpj_row[1 + h] : a.pil_bar + a.pre1 * q;
        }
        for (int s = tid; s < S; s += blockDim.x) {
            const unsigned long long rec = a.states[s];
            double q = 0.0;
            for (int d = 0; d < D; ++d) {
                double wl, wb;
                wbar_at(a, rec, Wr_c, d, wl, wb);
                double e = wb - y_s[d];
                q = fma(e, e, q);
            }
            qb[s] = q;
            int nmem = 0;
            for (int m = 0; m < a.gamma; ++m) nmem += (int(unsigned(rec >> (8 * m)) & 0xFFu) != Hp);
            Fb[1 + H + s] = rd ? logpj_row[1 + H + s] : a.pil_bar * double(nmem) + a.pre1 * q;
        }
        __syncthreads();
        if (wr) {
            for (int c = tid; c < a.C; c += blockDim.x) logpj_row[c] = Fb[c];
            if (lse_only) continue;              // compat E_step: logpj only
        }

        // ---- phase 3: annealed posterior --------------------------------------------------------
        double mx = -INFINITY, mx1 = -INFINITY;
        for (int c = tid; c < a.C; c += blockDim.x) { mx1 = fmax(mx1, Fb[c]); }
        mx1 = block_reduce(mx1, red, true);
        mx = a.beta * mx1;                                   // corr = beta * max(logpj)   (mca_et.py:237)
        double s1 = 0.0, sb = 0.0;
        for (int c = tid; c < a.C; c += blockDim.x) {
            double f = Fb[c];
            s1 += exp(f - mx1);                              // un-annealed, for Q (mca_et.py:327)
            double pb = exp(a.beta * f - mx);
            sb += pb;
            Fb[c] = pb;                                      // pjb
        }
        s1 = block_reduce(s1, red, false);
        sb = block_reduce(sb, red, false);
        const double lse_b = mx + log(sb);                   // log(pjb.sum) + corr  (mca_et.py:250)
        const double lse_1 = mx1 + log(s1);
        if (tid == 0) a.lse[n] = lse_b;
        if (lse_only) continue;
        const double inv = 1.0 / sb;

        // ---- phase 4: statistics ------------------------------------------------------------------
        double pi_p = 0.0, sg_p = 0.0;
        if (tid == 0) sg_p = Fb[0] * yy;
        for (int h = tid; h < a.ldH; h += blockDim.x) {
            double p = (h < H) ? Fb[1 + h] : 0.0;
            a.Spost[r * a.ldH + h] = p * inv;                    // singles posterior -> statistics GEMM
            if (h < H) {
                pi_p += p;
                sg_p = fma(p, yy - 2.0 * row[h] + a.wn2[h], sg_p);
            }
        }
        for (int s = tid; s < S; s += blockDim.x) {
            const unsigned long long rec = a.states[s];
            int nmem = 0;
            for (int m = 0; m < a.gamma; ++m) nmem += (int(unsigned(rec >> (8 * m)) & 0xFFu) != Hp);
            double p = Fb[1 + H + s];
            pi_p = fma(p, double(nmem), pi_p);
            sg_p = fma(p, qb[s], sg_p);
        }
        // Aid[j][d] = sum_{s contains j} exp(logpjb_s + model term(s,j,d)); one thread per feature d
        for (int d = tid; d < D; d += blockDim.x) {
            for (int j = 0; j < Hp; ++j) Aid[j * ldc + d] = 0.0;
            for (int s = 0; s < S; ++s) {
                const double pb = Fb[1 + H + s];
                if (pb == 0.0) continue;
                const unsigned long long rec = a.states[s];
                double wl, wb;
                wbar_at(a, rec, Wr_c, d, wl, wb);
                const double lpb = log(pb);                  // = beta*logpj_s - corr
                for (int m = 0; m < a.gamma; ++m) {
                    int j = int(unsigned(rec >> (8 * m)) & 0xFFu);
                    if (j == Hp) break;
                    double e;
                    if (a.mmca) e = lpb - (a.rho - 1.0) * fmax(wl - Wl_c[j * ldc + d], 0.0);     // mmca_et.py:338-340
                    else e = lpb + (1.0 - a.rho) * wl + (a.rho - 1.0) * Wl_c[j * ldc + d];        // mca_et.py:309
                    Aid[j * ldc + d] += exp(e);
                }
            }
            const double yd = y_s[d];
            for (int j = 0; j < Hp; ++j) {
                double v = Aid[j * ldc + d] * inv;
                if (v != 0.0) {
                    atomicAdd(&a.Wpm[int64_t(cand_s[j]) * a.ldD + d], v * yd);
                    atomicAdd(&a.Wqm[int64_t(cand_s[j]) * a.ldD + d], v);
                }
            }
        }
        pi_p = block_reduce(pi_p, red, false);
        sg_p = block_reduce(sg_p, red, false);
        if (tid == 0) {
            acc_n += 1.0;
            acc_ld += lse_1;
            acc_sig += sg_p * inv;
            acc_pi += pi_p * inv;
        }
    }
    if (tid == 0 && !(a.flags & GLF_LSE_ONLY)) {
        atomicAdd(&a.scalars[0], acc_n);
        atomicAdd(&a.scalars[1], acc_ld);
        atomicAdd(&a.scalars[2], acc_sig);
        atomicAdd(&a.scalars[3], acc_pi);
    }
}

size_t mca_smem_bytes(const MCAArgs &a) {
    size_t d = ((a.D + 1) & ~1) + ((a.H + 1) & ~1) + size_t(a.Hp + 1) * a.ldc + size_t(a.Hp) * a.ldc +
               ((a.C + 1) & ~1) + ((a.S + 1) & ~1) + size_t(a.Hp) * a.ldc + 32 + 16;
    return d * sizeof(double);
}

int launch_mca_kernel(const MCAArgs &a, int sm_count, cudaStream_t stream) {
    size_t smem = mca_smem_bytes(a);
    if (smem > 227 * 1024) {
        set_error("MCA/MMCA kernel needs %zu bytes of shared memory (D=%d, H=%d, Hprime=%d): unsupported size", smem, a.D, a.H, a.Hp);
        return PET_EINVAL;
    }
    static size_t configured = 0;
    if (smem > configured) {
        PET_CUDA(cudaFuncSetAttribute(mca_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    int per_sm = int(std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 1024))));
    int64_t grid = std::min<int64_t>(a.n_rows, int64_t(sm_count) * per_sm);
    if (grid <= 0) return PET_OK;
    mca_kernel<<<(unsigned)grid, MCA_THREADS, smem, stream>>>(a);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// ---- per-iteration tables and the MCA preselection score ---------------------------------------
// Wl = log|W|, Wr = |W|^rho (MCA, W >= W_tol) or sign(W)|W|^rho (MMCA)      mca_et.py:149-150, mmca_et.py:169-171
__global__ void mca_tables_kernel(const double *Wt, int64_t ldk, int H, int D, double rho, int mmca, double *Wl,
                                  double *Wr, int64_t ldD, double *wn2) {
    int h = blockIdx.x;
    double s = 0.0;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double w = Wt[int64_t(h) * ldk + d];
        double l = log(fabs(w));
        double r = exp(rho * l);
        Wl[int64_t(h) * ldD + d] = l;
        Wr[int64_t(h) * ldD + d] = (mmca && w < 0.0) ? -r : r;
        s = fma(w, w, s);
    }
    __shared__ double red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
        wn2[h] = t;
    }
}

// sim[n][h] = sum_d max(W_hd - y_d, 0)   (= sum |max(W,y) - y|, mca_et.py:104-106); W is (D,H) h-contiguous
__global__ void mca_sim_kernel(const double *Y, int64_t ldY, int64_t rows, const double *W, int64_t ldW, int D, int H,
                               double *sim, int64_t ldH) {
    extern __shared__ double ys[];
    int64_t r = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) ys[d] = Y[r * ldY + d];
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        double s = 0.0;
        for (int d = 0; d < D; ++d) s += fmax(W[int64_t(d) * ldW + h] - ys[d], 0.0);
        sim[r * ldH + h] = s;
    }
}

int launch_mca_tables(const double *Wt, int64_t ldk, int H, int D, double rho, int mmca, double *Wl, double *Wr,
                      int64_t ldD, double *wn2, cudaStream_t st) {
    mca_tables_kernel<<<H, 128, 0, st>>>(Wt, ldk, H, D, rho, mmca, Wl, Wr, ldD, wn2);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

int launch_mca_sim(const double *Y, int64_t ldY, int64_t rows, const double *W, int64_t ldW, int D, int H, double *sim,
                   int64_t ldH, cudaStream_t st) {
    if (rows <= 0) return PET_OK;
    mca_sim_kernel<<<(unsigned)rows, 128, D * sizeof(double), st>>>(Y, ldY, rows, W, ldW, D, H, sim, ldH);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// W_new (D,H) from the all-reduced statistics                                mca_et.py:330-350, mmca_et.py:370-394
//   A(d,h)  = (Y^T . singles posterior)(d,h), colsum(h) = sum_n singles posterior  (row D of the GEMM output)
__global__ void mca_update_kernel(const double *A, int64_t ldA, const double *Wpm, const double *Wqm, int64_t ldD,
                                  const double *Wt, int64_t ldk, int D, int H, int mmca, double tol, double *W_new,
                                  int64_t ldo) {
    int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= int64_t(D) * H) return;
    int d = int(idx / H), h = int(idx % H);
    const double w = Wt[int64_t(h) * ldk + d];
    const double colsum = A[int64_t(D) * ldA + h];
    double wp, wq, out;
    if (mmca) {
        wp = A[int64_t(d) * ldA + h] + Wpm[int64_t(h) * ldD + d];                 // mmca_et.py:318
        wq = colsum + Wqm[int64_t(h) * ldD + d];                                  // :319
        wq = fmax(wq, tol);                                                       // :384-385
        double wn = wp / wq;
        double inertia = fmax(1.0 - exp(-wq / 2.5), 0.2);                         // :391-392
        out = inertia * wn + (1.0 - inertia) * w;
    } else {
        const double w2 = w * w;
        wp = w2 * A[int64_t(d) * ldA + h] + Wpm[int64_t(h) * ldD + d];            // mca_et.py:294
        wq = w2 * colsum + Wqm[int64_t(h) * ldD + d];                             // :295
        const double tiny = 2.2250738585072014e-308;
        if (wq < tiny) { wp = 0.0; wq = tiny; }                                   // :344-346
        out = wp / wq;
    }
    W_new[int64_t(d) * ldo + h] = out;
}

int launch_mca_update(const double *A, int64_t ldA, const double *Wpm, const double *Wqm, int64_t ldD, const double *Wt,
                      int64_t ldk, int D, int H, int mmca, double tol, double *W_new, int64_t ldo, cudaStream_t st) {
    int64_t tot = int64_t(D) * H;
    mca_update_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(A, ldA, Wpm, Wqm, ldD, Wt, ldk, D, H, mmca, tol, W_new, ldo);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

}  // namespace pet
