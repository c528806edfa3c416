// Mixture models (prosper/em/mixturemodels/{__init__,MoG,MoP}.py): the dense (n x H)
// posterior and the element-wise helpers around the engine's GEMMs.  The contractions themselves are
// pet_dgemm_kk / pet_dgemm_mn calls issued by the host mirror (prosper_b200/em/mixturemodels).
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace pet {

// lp[n][h] = beta * (s1 T1[n][h] + s2 T2[n][h] + k[h])  (the log-joint, returned as 'logpj');
// post = exp(lp) with the reference's clamps (NaN -> tiny, < tiny -> tiny, inf -> max/H), rows normalised
// (MoG.py:208-218, MoP.py:165-175: NOT max-shifted upstream, so rows that underflow become uniform; kept).
__global__ void mix_posterior_kernel(int64_t n, int H, const double *T1, const double *T2, int64_t ldt, double s1, double s2,
                                     const double *k, double beta, double *lp, int64_t ldl, double *post, int64_t ldp) {
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const double mxv = DBL_MAX / double(H);
    double sum = 0.0;
    for (int h = lane; h < H; h += 32) {
        const double t2 = T2 ? T2[r * ldt + h] : 0.0;
        const double v = beta * (s1 * T1[r * ldt + h] + s2 * t2 + k[h]);
        lp[r * ldl + h] = v;
        double p = exp(v);
        if (p != p || p < DBL_MIN) p = DBL_MIN;
        if (isinf(p)) p = mxv;
        post[r * ldp + h] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    for (int h = lane; h < H; h += 32) post[r * ldp + h] /= sum;
}

// out[n * stride] = sum_d A[n][d] B[n][d]
__global__ void rowdot_kernel(int64_t n, int D, const double *A, int64_t lda, const double *B, int64_t ldb, double *out,
                              int64_t stride) {
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) s = fma(A[r * lda + d], B[r * ldb + d], s);
    s = warp_sum(s);
    if (lane == 0) out[r * stride] = s;
}

// out[n][d] = op(X[n][d]):  0: X^2   1: X * w[n * wstride]   2: X - v[d]   3: (a / (sum_d X[n][:] + eps)) * X + 1 (MoP.py:236-244)
__global__ void rowop_kernel(int op, int64_t n, int D, const double *X, int64_t ldx, const double *w, int64_t wstride, double a,
                             double *out, int64_t ldo) {
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    double f = 0.0;
    if (op == 1) f = w[r * wstride];
    if (op == 3) {
        double s = 0.0;
        for (int d = lane; d < D; d += 32) s += X[r * ldx + d];
        f = a / (warp_sum(s) + DBL_EPSILON);
    }
    for (int d = lane; d < ldo; d += 32) {
        double v = 0.0;
        if (d < D) {
            const double x = X[r * ldx + d];
            v = (op == 0) ? x * x : (op == 1) ? x * f : (op == 2) ? x - w[d] : f * x + 1.0;
        }
        out[r * ldo + d] = v;
    }
}

}  // namespace pet

using namespace pet;

extern "C" int pet_mix_posterior(int64_t n, int32_t H, const double *T1_dev, const double *T2_dev, int64_t ldt, double s1, double s2,
                                 const double *k_dev, double beta, double *logpj_dev, int64_t ld_logpj, double *post_dev,
                                 int64_t ld_post, void *stream) {
    if (n < 0 || H < 1 || !T1_dev || !k_dev || !logpj_dev || !post_dev || ldt < H || ld_logpj < H || ld_post < H) {
        set_error("pet_mix_posterior: bad arguments");
        return PET_EINVAL;
    }
    if (n == 0) return PET_OK;
    mix_posterior_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(n, H, T1_dev, T2_dev, ldt, s1, s2, k_dev, beta,
                                                                                          logpj_dev, ld_logpj, post_dev, ld_post);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

extern "C" int pet_rowdot(int64_t n, int32_t D, const double *A_dev, int64_t lda, const double *B_dev, int64_t ldb, double *out_dev,
                          int64_t out_stride, void *stream) {
    if (n < 0 || D < 1 || !A_dev || !B_dev || !out_dev || lda < D || ldb < D || out_stride < 1) { set_error("pet_rowdot: bad arguments"); return PET_EINVAL; }
    if (n == 0) return PET_OK;
    rowdot_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(n, D, A_dev, lda, B_dev, ldb, out_dev, out_stride);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

extern "C" int pet_rowop(int32_t op, int64_t n, int32_t D, const double *X_dev, int64_t ldx, const double *w_dev, int64_t w_stride,
                         double a, double *out_dev, int64_t ldo, void *stream) {
    if (op < 0 || op > 3 || n < 0 || D < 1 || !X_dev || !out_dev || ldx < D || ldo < D || ((op == 1 || op == 2) && !w_dev)) {
        set_error("pet_rowop: bad arguments");
        return PET_EINVAL;
    }
    if (n == 0) return PET_OK;
    rowop_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(op, n, D, X_dev, ldx, w_dev, w_stride, a, out_dev, ldo);
    PET_LAUNCH_CHECK();
    return PET_OK;
}
