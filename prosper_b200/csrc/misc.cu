// Small data-movement kernels and the radix k-select of the truncation rule.
#include <math.h>

#include "common.cuh"

namespace pet {

// Wt(H, ldk) = W(D, ldw)^T, zero padded in k  (bsc_et.py:142 `W = model_params['W'].T`)
__global__ void transpose_w_kernel(double *Wt, int64_t ldk, const double *W, int64_t ldw, int D, int H) {
    __shared__ double t[32][33];
    int h0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int d = d0 + i, h = h0 + threadIdx.x;
        t[i][threadIdx.x] = (d < D && h < H) ? W[int64_t(d) * ldw + h] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int h = h0 + i, d = d0 + threadIdx.x;
        if (h < H && d < ldk) Wt[int64_t(h) * ldk + d] = (d < D) ? t[threadIdx.x][i] : 0.0;
    }
}

__global__ void gram_diag_kernel(const double *G, int64_t ldg, int H, double *wn2, double *invn) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    double g = G[int64_t(h) * ldg + h];
    wn2[h] = g;
    invn[h] = 1.0 / sqrt(g);
}

// wmu[h] = W_h . mu (one warp per h): the BSC selection scores the RAW datapoint (bsc_et.py:110-112) while the shard
// is stored shifted by mu, so (W_h . y) = (W_h . (y - mu)) + (W_h . mu)
__global__ void wdotmu_kernel(const double *Wt, int64_t ldk, int H, int D, const double *mu, double *out) {
    int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (h >= H) return;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) s = fma(Wt[int64_t(h) * ldk + d], mu[d], s);
    s = warp_sum(s);
    if (lane == 0) out[h] = s;
}

// y <- y - mu (BSC offset, bsc_et.py:169); rare path, only when mu is non-zero
__global__ void subtract_mu_kernel(double *Y, int64_t ldy, int64_t n, int D, const double *mu) {
    int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n * D) return;
    int64_t r = idx / D;
    int d = int(idx % D);
    Y[r * ldy + d] -= mu[d];
}

// one warp per row: yy[n] = ||y_n||^2, and the engine's layout columns: Y[n][D] = 1, rest 0
// src != Y: the rows are first copied from a (contiguous) upload staging buffer
__global__ void rownorm_pad_kernel(const double *src, int64_t ld_src, double *Y, int64_t ldy, int64_t n, int D, double *yy) {
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    double *y = Y + row * ldy;
    const double *x = src + row * ld_src;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) {
        const double v = x[d];
        if (x != y) y[d] = v;
        s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) yy[row] = s;
    for (int d = D + lane; d < ldy; d += 32) y[d] = (d == D) ? 1.0 : 0.0;
}

__global__ void cand_to_i64_kernel(int64_t *out, const int *in, int64_t count) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) out[i] = in[i];
}
__global__ void cand_from_i64_kernel(int *out, const int64_t *in, int64_t count, int H) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) {
        int64_t v = in[i];
        out[i] = int(v < 0 ? 0 : (v >= H ? H - 1 : v));
    }
}

// binary models: Wq[h][h] += sum_n <s_h>  (row D of the statistics GEMM output)
__global__ void add_diag_kernel(double *Wq, int64_t ld, const double *colsum, int H) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < H) Wq[int64_t(h) * ld + h] += colsum[h];
}

// column sums of a (rows, ld) matrix accumulated into out[0..cols) -- DSC second moments
__global__ void colsum_kernel(double *out, const double *M, int64_t ld, int64_t rows, int cols) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    int64_t r0 = int64_t(blockIdx.y) * 256, r1 = (rows < r0 + 256) ? rows : r0 + 256;
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) s += M[r * ld + c];
    atomicAdd(&out[c], s);
}
// out[c] += sum over the rows kept by the truncation rule lse >= cut (> cut if strict) of M[r][c]: data_sum of the BSC
// mu update (bsc_et.py:282 after the cut of :250-257)
__global__ void colsum_kept_kernel(double *out, const double *M, int64_t ld, int64_t rows, int cols, const double *lse,
                                   const double *cut, int strict) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    int64_t r0 = int64_t(blockIdx.y) * 256, r1 = (rows < r0 + 256) ? rows : r0 + 256;
    const double cv = cut ? *cut : 0.0;
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        const bool keep = !cut || (strict ? lse[r] > cv : lse[r] >= cv);
        if (keep) s += M[r * ld + c];
    }
    atomicAdd(&out[c], s);
}
// out[c] += sum_r M[r][c]^2   (GSC sigma update: sum_n y_nd^2, gsc_et.py:696,710)
__global__ void colsumsq_kernel(double *out, const double *M, int64_t ld, int64_t rows, int cols) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    int64_t r0 = int64_t(blockIdx.y) * 256, r1 = (rows < r0 + 256) ? rows : r0 + 256;
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) { double v = M[r * ld + c]; s = fma(v, v, s); }
    atomicAdd(&out[c], s);
}
__global__ void add_diag_from_vec_kernel(double *Wq, int64_t ld, const double *v, int H) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < H) Wq[int64_t(h) * ld + h] += v[h];
}

// ---- k-th largest of n doubles: 8-bit MSD radix select ----------------------------------
// Replaces `parallel.allsort(all_denoms)[-N_use]` (bsc_et.py:252, parallel.py:87-110): only one
// order statistic of the sorted array is ever used.
__device__ __forceinline__ unsigned long long key_of(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);   // ascending order-preserving
}
__device__ __forceinline__ double val_of(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// state: [0]=prefix key bits decided so far, [1]=remaining rank k (1-based among matching), hist[256]
__global__ void ksel_init_kernel(unsigned long long *state, unsigned long long k) {
    if (threadIdx.x == 0) { state[0] = 0ull; state[1] = k; }
    state[2 + threadIdx.x] = 0ull;
}
__global__ void ksel_hist_kernel(const double *v, int64_t n, int pass, unsigned long long *state) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    const unsigned long long prefix = state[0];
    const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        unsigned long long k = key_of(v[i]);
        if ((k & mask) == prefix) atomicAdd(&h[(k >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&state[2 + threadIdx.x], (unsigned long long)h[threadIdx.x]);
}
__global__ void ksel_pick_kernel(unsigned long long *state, int pass, double *out) {
    if (threadIdx.x == 0) {
        const int shift = 56 - 8 * pass;
        unsigned long long k = state[1], cum = 0;
        int d = 255;
        for (; d > 0; --d) {
            unsigned long long c = state[2 + d];
            if (cum + c >= k) break;
            cum += c;
        }
        state[1] = k - cum;
        state[0] |= (unsigned long long)d << shift;
        if (pass == 7) *out = val_of(state[0]);
    }
    __syncthreads();
    state[2 + threadIdx.x] = 0ull;
}

// the same select for small arrays in ONE launch (the bars configurations are launch bound: 17 launches -> 1)
constexpr int KSEL_SMALL_N = 65536;
__global__ void __launch_bounds__(1024) ksel_small_kernel(const double *v, int n, unsigned long long k, double *out) {
    __shared__ unsigned int h[256];
    __shared__ unsigned long long s_prefix, s_k;
    if (threadIdx.x == 0) { s_prefix = 0ull; s_k = k; }
    for (int pass = 0; pass < 8; ++pass) {
        if (threadIdx.x < 256) h[threadIdx.x] = 0u;
        __syncthreads();
        const int shift = 56 - 8 * pass;
        const unsigned long long prefix = s_prefix;
        const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = key_of(v[i]);
            if ((key & mask) == prefix) atomicAdd(&h[(key >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long kk = s_k, cum = 0;
            int d = 255;
            for (; d > 0; --d) {
                const unsigned long long c = h[d];
                if (cum + c >= kk) break;
                cum += c;
            }
            s_k = kk - cum;
            s_prefix = prefix | ((unsigned long long)d << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = val_of(s_prefix);
}

int kth_largest(const double *vals, int64_t n, int64_t k, double *out, unsigned long long *state, int sm_count,
                cudaStream_t st) {
    if (n <= 0 || k < 1 || k > n) {
        set_error("kth_largest: need 1 <= k <= n (k=%lld, n=%lld)", (long long)k, (long long)n);
        return PET_EINVAL;
    }
    if (n <= KSEL_SMALL_N) {
        ksel_small_kernel<<<1, 1024, 0, st>>>(vals, int(n), (unsigned long long)k, out);
        PET_LAUNCH_CHECK();
        return PET_OK;
    }
    ksel_init_kernel<<<1, 256, 0, st>>>(state, (unsigned long long)k);
    PET_LAUNCH_CHECK();
    int blocks = int(std::min<int64_t>(ceil_div(n, 256), int64_t(sm_count) * 8));
    for (int pass = 0; pass < 8; ++pass) {
        ksel_hist_kernel<<<blocks, 256, 0, st>>>(vals, n, pass, state);
        PET_LAUNCH_CHECK();
        ksel_pick_kernel<<<1, 256, 0, st>>>(state, pass, out);
        PET_LAUNCH_CHECK();
    }
    return PET_OK;
}

// ---- launch wrappers ----------------------------------------------------------------------
int launch_transpose_w(double *Wt, int64_t ldk, const double *W, int64_t ldw, int D, int H, cudaStream_t st) {
    dim3 g((unsigned)ceil_div(H, 32), (unsigned)ceil_div(ldk, 32)), b(32, 8);
    transpose_w_kernel<<<g, b, 0, st>>>(Wt, ldk, W, ldw, D, H);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
int launch_gram_diag(const double *G, int64_t ldg, int H, double *wn2, double *invn, cudaStream_t st) {
    gram_diag_kernel<<<(unsigned)ceil_div(H, 256), 256, 0, st>>>(G, ldg, H, wn2, invn);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
int launch_wdotmu(const double *Wt, int64_t ldk, int H, int D, const double *mu, double *out, cudaStream_t st) {
    wdotmu_kernel<<<(unsigned)ceil_div(int64_t(H) * 32, 256), 256, 0, st>>>(Wt, ldk, H, D, mu, out);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
int launch_subtract_mu(double *Y, int64_t ldy, int64_t n, int D, const double *mu, cudaStream_t st) {
    if (n <= 0) return PET_OK;
    subtract_mu_kernel<<<(unsigned)ceil_div(n * D, 256), 256, 0, st>>>(Y, ldy, n, D, mu);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
int launch_rownorm_pad(const double *src, int64_t ld_src, double *Y, int64_t ldy, int64_t n, int D, double *yy, cudaStream_t st) {
    if (n <= 0) return PET_OK;
    rownorm_pad_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(src, ld_src, Y, ldy, n, D, yy);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
int launch_cand_to_i64(int64_t *out, const int *in, int64_t count, cudaStream_t st) {
    if (count <= 0) return PET_OK;
    cand_to_i64_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(out, in, count);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
int launch_cand_from_i64(int *out, const int64_t *in, int64_t count, int H, cudaStream_t st) {
    if (count <= 0) return PET_OK;
    cand_from_i64_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(out, in, count, H);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
int launch_add_diag(double *Wq, int64_t ld, const double *colsum, int H, cudaStream_t st) {
    add_diag_kernel<<<(unsigned)ceil_div(H, 256), 256, 0, st>>>(Wq, ld, colsum, H);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
int launch_colsumsq(double *out, const double *M, int64_t ld, int64_t rows, int cols, cudaStream_t st) {
    if (rows <= 0) return PET_OK;
    dim3 g((unsigned)ceil_div(cols, 128), (unsigned)ceil_div(rows, 256));
    colsumsq_kernel<<<g, 128, 0, st>>>(out, M, ld, rows, cols);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
int launch_colsum(double *out, const double *M, int64_t ld, int64_t rows, int cols, cudaStream_t st) {
    if (rows <= 0) return PET_OK;
    dim3 g((unsigned)ceil_div(cols, 128), (unsigned)ceil_div(rows, 256));
    colsum_kernel<<<g, 128, 0, st>>>(out, M, ld, rows, cols);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

int launch_colsum_kept(double *out, const double *M, int64_t ld, int64_t rows, int cols, const double *lse, const double *cut,
                       int strict, cudaStream_t st) {
    if (rows <= 0) return PET_OK;
    dim3 g((unsigned)ceil_div(cols, 128), (unsigned)ceil_div(rows, 256));
    colsum_kept_kernel<<<g, 128, 0, st>>>(out, M, ld, rows, cols, lse, cut, strict);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

}  // namespace pet
