// Synthetic data and initialisation on the device, so that large benchmarks never
// touch the host: model data generation (camodels/__init__.py:104-122, bsc_et.py:67-95, tsc_et.py:214-275,
// dsc_et.py:238-300, mca_et.py:65-91) and the pieces of standard_init (camodels/__init__.py:196-235).
// Random numbers: counter-based Philox4x32-10 keyed by (seed, stream) with the counter = (row, column), so
// every rank / chunk / launch geometry draws the same values for the same global element.  Parity with
// np.random's Mersenne Twister streams is neither possible nor required (distributional tests only).
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace pet {

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t a, uint32_t b) const {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ a, n1 = lo1, n2 = hi0 ^ c[3] ^ b, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // 128 random bits for counter (i, j)
    __device__ __forceinline__ void draw(uint64_t i, uint64_t j, uint32_t (&c)[4]) const {
        c[0] = uint32_t(i); c[1] = uint32_t(i >> 32); c[2] = uint32_t(j); c[3] = uint32_t(j >> 32);
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) { round(c, a, b); a += 0x9E3779B9u; b += 0xBB67AE85u; }
    }
};

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {      // uniform in (0, 1)
    const uint64_t m = (uint64_t(hi) << 21) ^ uint64_t(lo >> 11);       // 53 bits
    return (double(m) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double normal_from(const uint32_t (&c)[4]) {  // Box-Muller, one value per counter
    const double u1 = u53(c[0], c[1]), u2 = u53(c[2], c[3]);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

constexpr int GEN_MAXK = 16;
struct GenArgs {
    int64_t n, row0, ldy, ldw, lds;
    int D, H, K, combine;                  // combine: 0 = sum_h s_h W_h, 1 = entry of largest magnitude (mca_et.py:83-85)
    double values[GEN_MAXK], cum[GEN_MAXK];
    double sigma;
    const double *W;                       // (D, H) row-major, as model_params['W']
    double *y; int8_t *s_idx;              // s_idx (n, H): index k of the drawn value, optional
    uint64_t seed;
};

// one warp per datapoint
__global__ void __launch_bounds__(256) generate_kernel(const GenArgs a) {
    __shared__ int act_h[8][64];
    __shared__ double act_v[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Philox lat{uint32_t(a.seed), uint32_t(a.seed >> 32) ^ 0x1234567u};
    const Philox noi{uint32_t(a.seed) ^ 0x9E3779B9u, uint32_t(a.seed >> 32) ^ 0x89ABCDEFu};
    for (int64_t r = int64_t(blockIdx.x) * 8 + warp; r < a.n; r += int64_t(gridDim.x) * 8) {
        const uint64_t grow = uint64_t(a.row0 + r);
        double *y = a.y + r * a.ldy;
        for (int d = lane; d < a.D; d += 32) y[d] = 0.0;
        // latents in chunks of 32 causes; non-zero ones are applied in batches of up to 64
        int n_act = 0;
        for (int h0 = 0; h0 < a.H || n_act > 0; h0 += 32) {
            if (h0 < a.H) {
                const int h = h0 + lane;
                int k = -1;
                double v = 0.0;
                if (h < a.H) {
                    uint32_t c[4];
                    lat.draw(grow, uint64_t(h), c);
                    const double u = u53(c[0], c[1]);
                    k = a.K - 1;
                    for (int t = 0; t < a.K - 1; ++t) if (u < a.cum[t]) { k = t; break; }
                    v = a.values[k];
                    if (a.s_idx) a.s_idx[r * a.lds + h] = (int8_t)k;
                }
                const unsigned m = __ballot_sync(0xffffffffu, v != 0.0);
                if (v != 0.0) {
                    const int slot = n_act + __popc(m & ((1u << lane) - 1u));
                    act_h[warp][slot] = h;
                    act_v[warp][slot] = v;
                }
                n_act += __popc(m);
                __syncwarp();
            }
            if (n_act > 32 || (h0 + 32 >= a.H && n_act > 0)) {
                for (int d = lane; d < a.D; d += 32) {
                    double acc = y[d];
                    const double *wr = a.W + int64_t(d) * a.ldw;
                    for (int i = 0; i < n_act; ++i) {
                        const double t = act_v[warp][i] * wr[act_h[warp][i]];
                        if (a.combine == 0) acc += t;
                        else if (fabs(t) > fabs(acc)) acc = t;          // first entry of largest magnitude wins
                    }
                    y[d] = acc;
                }
                n_act = 0;
                __syncwarp();
            }
        }
        for (int d = lane; d < a.D; d += 32) {
            uint32_t c[4];
            noi.draw(grow, uint64_t(d), c);
            y[d] += a.sigma * normal_from(c);
        }
    }
}

// X[i][j] = base[i] (or 0) + scale * N(0,1), counter (i, j): W_init = W_mean[:, None] + noise (camodels/__init__.py:226)
__global__ void normal_fill_kernel(double *X, int64_t ld, int64_t rows, int64_t cols, const double *row_base, double scale,
                                   uint64_t seed) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t i = idx / cols, j = idx % cols;
    const Philox g{uint32_t(seed) ^ 0x5bd1e995u, uint32_t(seed >> 32) ^ 0x1b873593u};
    uint32_t c[4];
    g.draw(uint64_t(i), uint64_t(j), c);
    X[i * ld + j] = (row_base ? row_base[i] : 0.0) + scale * normal_from(c);
}

// out[c] += sum_r (M[r][c] - mean[c])^2  (second pass of the data variance, camodels/__init__.py:220)
__global__ void col_centered_sumsq_kernel(double *out, const double *M, int64_t ld, int64_t rows, int cols, const double *mean) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int64_t r0 = int64_t(blockIdx.y) * 256, r1 = min(rows, r0 + 256);
    const double mu = mean[c];
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) { const double t = M[r * ld + c] - mu; s = fma(t, t, s); }
    atomicAdd(out + c, s);
}

}  // namespace pet

using namespace pet;

extern "C" int pet_generate_data(int32_t combine, int64_t n, int64_t row0, int32_t D, int32_t H, const double *W_dev, int64_t ldW,
                                 int32_t K, const double *values_host, const double *probs_host, double sigma, uint64_t seed,
                                 double *y_dev, int64_t ldy, int8_t *s_idx_dev, int64_t lds, void *stream) {
    if (n < 0 || D < 1 || H < 1 || !W_dev || !y_dev || ldW < H || ldy < D || K < 2 || K > GEN_MAXK || !values_host || !probs_host ||
        (s_idx_dev && lds < H) || combine < 0 || combine > 1 || !(sigma >= 0.0)) {
        set_error("pet_generate_data: bad arguments");
        return PET_EINVAL;
    }
    if (n == 0) return PET_OK;
    GenArgs a{};
    a.n = n; a.row0 = row0; a.ldy = ldy; a.ldw = ldW; a.lds = lds; a.D = D; a.H = H; a.K = K; a.combine = combine;
    double cum = 0.0;
    for (int k = 0; k < K; ++k) {
        if (!(probs_host[k] >= 0.0)) { set_error("pet_generate_data: negative probability"); return PET_EINVAL; }
        cum += probs_host[k];
        a.values[k] = values_host[k];
        a.cum[k] = cum;
    }
    if (fabs(cum - 1.0) > 1e-9) { set_error("pet_generate_data: probabilities sum to %.12g, not 1", cum); return PET_EINVAL; }
    a.sigma = sigma; a.W = W_dev; a.y = y_dev; a.s_idx = s_idx_dev; a.seed = seed;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 8), 148 * 16);
    generate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

extern "C" int pet_normal_fill(double *X_dev, int64_t ld, int64_t rows, int64_t cols, const double *row_base_dev, double scale,
                               uint64_t seed, void *stream) {
    if (!X_dev || rows < 0 || cols < 0 || ld < cols) { set_error("pet_normal_fill: bad arguments"); return PET_EINVAL; }
    if (rows * cols == 0) return PET_OK;
    normal_fill_kernel<<<(unsigned)ceil_div(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(X_dev, ld, rows, cols, row_base_dev,
                                                                                              scale, seed);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// dst[i][:] = src[idx[i]][:]: the datapoint subset of select_partial_data (camodels/__init__.py:125-152) taken from the
// device-resident shard, one warp per destination row
__global__ void gather_rows_kernel(double *dst, int64_t ld_dst, const double *src, int64_t ld_src, const int64_t *idx,
                                   int64_t n_sel, int64_t n_src, int cols) {
    const int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_sel) return;
    int64_t r = idx[i];
    r = r < 0 ? 0 : (r >= n_src ? n_src - 1 : r);
    for (int c = lane; c < cols; c += 32) dst[i * ld_dst + c] = src[r * ld_src + c];
}

extern "C" int pet_gather_rows(int64_t n_sel, int64_t n_src, int64_t cols, const double *src_dev, int64_t ld_src,
                               const int64_t *idx_dev, double *dst_dev, int64_t ld_dst, void *stream) {
    if (!src_dev || !idx_dev || !dst_dev || n_sel < 0 || n_src < 1 || cols < 1 || ld_src < cols || ld_dst < cols) {
        set_error("pet_gather_rows: bad arguments");
        return PET_EINVAL;
    }
    if (n_sel == 0) return PET_OK;
    gather_rows_kernel<<<(unsigned)ceil_div(n_sel * 32, 256), 256, 0, (cudaStream_t)stream>>>(dst_dev, ld_dst, src_dev, ld_src, idx_dev,
                                                                                             n_sel, n_src, (int)cols);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

extern "C" int pet_col_centered_sumsq(int64_t rows, int64_t cols, const double *M_dev, int64_t ld, const double *mean_dev,
                                      double *out_dev, void *stream) {
    if (!M_dev || !mean_dev || !out_dev || rows < 0 || cols < 1 || ld < cols) { set_error("pet_col_centered_sumsq: bad arguments"); return PET_EINVAL; }
    if (rows == 0) return PET_OK;
    dim3 g((unsigned)ceil_div(cols, 128), (unsigned)ceil_div(rows, 256));
    col_centered_sumsq_kernel<<<g, 128, 0, (cudaStream_t)stream>>>(out_dev, M_dev, ld, rows, (int)cols, mean_dev);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
