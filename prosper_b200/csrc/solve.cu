// On-device parameter solve of the M-step:  X . A = B  with A = Wq (H,H) symmetric PSD.
//
// Replaces `np.linalg.lstsq(Wq, Wp)` (bsc_et.py:377-380, dsc_et.py:732-735) and
// `np.dot(np.linalg.pinv(Wq), Wp)` (tsc_et.py:493).  Wq = sum_n <s s^T> is symmetric
// positive semi-definite, so a blocked right-looking Cholesky A = L L^T is used; a pivot
// below eps*n*max_diag is dropped (its row/column of L and the matching column of X become
// zero), which reproduces the minimum-norm answer lstsq/pinv give for a dead unit whose
// row/column of Wq is zero.  B is held transposed, (D,H), so X = W_new comes out directly
// in the reference's (D,H) layout and every rank-k update is a K-contiguous DMMA GEMM.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace pet {

int dgemm_kk(int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B,
             int64_t ldb, double *C, int64_t ldc, double alpha, int accumulate, cudaStream_t st);

constexpr int NB = 64;

// scal[0] = max diagonal, scal[1] = dropped-pivot counter
__global__ void maxdiag_kernel(const double *A, int64_t lda, int n, double *scal, double tol_scale) {
    __shared__ double red[32];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(A[int64_t(i) * lda + i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        m = warp_max(m);
        if (threadIdx.x == 0) { scal[0] = m; scal[1] = 0.0; scal[2] = tol_scale; }
    }
}

// Unblocked Cholesky of the nb x nb diagonal block at (j0,j0); writes L (lower, zero upper)
// and invd[j0+c] = 1/L_cc (0 for dropped pivots).
__global__ void __launch_bounds__(256) potrf_block_kernel(double *A, int64_t lda, int j0, int nb, int n,
                                                          double *invd, double *scal) {
    __shared__ double T[NB][NB + 1];
    __shared__ double s_inv;
    // LAPACK gelsd with rcond < 0 (what bsc_et.py:377-380 passes on NumPy 2.x) keeps singular values above
    // eps * sigma_max with eps = 2^-53; the largest diagonal entry stands in for sigma_max here
    const double tol = scal[0] * scal[2];
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
        int r = idx / nb, c = idx % nb;
        T[r][c] = (c <= r) ? A[int64_t(j0 + r) * lda + j0 + c] : 0.0;
    }
    __syncthreads();
    for (int c = 0; c < nb; ++c) {
        if (threadIdx.x == 0) {
            double d = T[c][c];
            if (d > tol) {
                double l = sqrt(d);
                T[c][c] = l;
                s_inv = 1.0 / l;
            } else {
                T[c][c] = 0.0;
                s_inv = 0.0;
                atomicAdd(&scal[1], 1.0);
            }
            invd[j0 + c] = s_inv;
        }
        __syncthreads();
        const double inv = s_inv;
        for (int r = c + 1 + threadIdx.x; r < nb; r += blockDim.x) T[r][c] *= inv;
        __syncthreads();
        // trailing update of the lower triangle: T[r][cc] -= L[r][c] * L[cc][c], c < cc <= r
        int rem = nb - c - 1;
        for (int idx = threadIdx.x; idx < rem * rem; idx += blockDim.x) {
            int r = c + 1 + idx / rem, cc = c + 1 + idx % rem;
            if (cc <= r) T[r][cc] -= T[r][c] * T[cc][c];
        }
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
        int r = idx / nb, c = idx % nb;
        A[int64_t(j0 + r) * lda + j0 + c] = T[r][c];
    }
}

// Row-wise triangular solves against the nb x nb diagonal block L_jj (at (j0,j0) of L):
//   forward  (backward=0): x . L_jj^T = r   ->  x_c = (r_c - sum_{t<c} x_t L[c][t]) * invd[c]
//   backward (backward=1): x . L_jj   = r   ->  x_c = (r_c - sum_{t>c} x_t L[t][c]) * invd[c]
// applied in place to rows [0,m) of R (ldr), columns [j0, j0+nb).
__global__ void __launch_bounds__(NB) trsm_rows_kernel(double *R, int64_t ldr, int64_t m, const double *L,
                                                        int64_t ldl, int j0, int nb, const double *invd,
                                                        int backward) {
    extern __shared__ __align__(16) double trsm_smem[];
    double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(trsm_smem);
    double (*Xs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(trsm_smem + NB * (NB + 1));
    double *inv_s = trsm_smem + 2 * NB * (NB + 1);
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
        int r = idx / nb, c = idx % nb;
        Ls[r][c] = L[int64_t(j0 + r) * ldl + j0 + c];
    }
    if (threadIdx.x < nb) inv_s[threadIdx.x] = invd[j0 + threadIdx.x];
    const int64_t row0 = int64_t(blockIdx.x) * NB;
    const int rows = (m - row0 < NB) ? int(m - row0) : NB;
    // coalesced load of the row slab
    for (int idx = threadIdx.x; idx < rows * nb; idx += blockDim.x) {
        int r = idx / nb, c = idx % nb;
        Xs[r][c] = R[(row0 + r) * ldr + j0 + c];
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < rows) {
        if (!backward) {
            for (int c = 0; c < nb; ++c) {
                double s = Xs[r][c];
                for (int t = 0; t < c; ++t) s = fma(-Xs[r][t], Ls[c][t], s);
                Xs[r][c] = s * inv_s[c];
            }
        } else {
            for (int c = nb - 1; c >= 0; --c) {
                double s = Xs[r][c];
                for (int t = c + 1; t < nb; ++t) s = fma(-Xs[r][t], Ls[t][c], s);
                Xs[r][c] = s * inv_s[c];
            }
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows * nb; idx += blockDim.x) {
        int rr = idx / nb, c = idx % nb;
        R[(row0 + rr) * ldr + j0 + c] = Xs[rr][c];
    }
}

// out(n,n) = in^T (lower factor -> its transpose), tiled through shared memory
__global__ void transpose_sq_kernel(double *out, const double *in, int64_t ld, int n) {
    __shared__ double t[32][33];
    int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = by + i, c = bx + threadIdx.x;
        t[i][threadIdx.x] = (r < n && c < n) ? in[int64_t(r) * ld + c] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = bx + i, c = by + threadIdx.x;
        if (r < n && c < n) out[int64_t(r) * ld + c] = t[threadIdx.x][i];
    }
}

// zero the strictly-upper part of the factor (the trailing GEMM updates the full square)
__global__ void zero_upper_kernel(double *A, int64_t lda, int n) {
    int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= int64_t(n) * n) return;
    int r = int(idx / n), c = int(idx % n);
    if (c > r) A[int64_t(r) * lda + c] = 0.0;
}

// A (n,n) lda : overwritten by L.  B (m,n) ldb : overwritten by X with X.A = B.
// work: n*lda doubles (L^T) + n (invd) + 2 (scalars).
int spd_solve_right(int64_t n64, int64_t m, double *A, int64_t lda, double *B, int64_t ldb, double *work,
                    cudaStream_t st) {
    const int n = int(n64);
    constexpr size_t TRSM_SMEM = (2 * NB * (NB + 1) + NB) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        PET_CUDA(cudaFuncSetAttribute(trsm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TRSM_SMEM)));
        configured = true;
    }
    double *Lt = work;
    double *invd = work + n64 * lda;
    double *scal = invd + round_up(n64, 2);
    static double tol_scale = []() { const char *e = getenv("PET_PIVOT_TOL"); return e ? atof(e) : 1.1102230246251565e-16; }();
    maxdiag_kernel<<<1, 256, 0, st>>>(A, lda, n, scal, tol_scale);
    PET_LAUNCH_CHECK();
    // ---- factor ----
    for (int j0 = 0; j0 < n; j0 += NB) {
        int nb = std::min(NB, n - j0);
        potrf_block_kernel<<<1, 256, 0, st>>>(A, lda, j0, nb, n, invd, scal);
        PET_LAUNCH_CHECK();
        int j1 = j0 + nb;
        if (j1 < n) {
            int64_t rows = n - j1;
            // panel: L[j1:, j0:j1] = A[j1:, j0:j1] . L_jj^{-T}
            trsm_rows_kernel<<<(unsigned)ceil_div(rows, NB), NB, TRSM_SMEM, st>>>(A + int64_t(j1) * lda, lda, rows, A,
                                                                         lda, j0, nb, invd, 0);
            PET_LAUNCH_CHECK();
            // trailing: A[j1:, j1:] -= L[j1:, j0:j1] . L[j1:, j0:j1]^T
            PET_CHECK(dgemm_kk(rows, rows, nb, A + int64_t(j1) * lda + j0, lda, A + int64_t(j1) * lda + j0,
                               lda, A + int64_t(j1) * lda + j1, lda, -1.0, 1, st));
        }
    }
    {
        int64_t tot = n64 * n64;
        zero_upper_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(A, lda, n);
        PET_LAUNCH_CHECK();
        dim3 g((unsigned)ceil_div(n, 32), (unsigned)ceil_div(n, 32)), b(32, 8);
        transpose_sq_kernel<<<g, b, 0, st>>>(Lt, A, lda, n);
        PET_LAUNCH_CHECK();
    }
    if (m <= 0) return PET_OK;
    const unsigned row_blocks = (unsigned)ceil_div(m, NB);
    // ---- Z . L^T = B, column blocks ascending ----
    for (int j0 = 0; j0 < n; j0 += NB) {
        int nb = std::min(NB, n - j0);
        if (j0 > 0)
            PET_CHECK(dgemm_kk(m, nb, j0, B, ldb, A + int64_t(j0) * lda, lda, B + j0, ldb, -1.0, 1, st));
        trsm_rows_kernel<<<row_blocks, NB, TRSM_SMEM, st>>>(B, ldb, m, A, lda, j0, nb, invd, 0);
        PET_LAUNCH_CHECK();
    }
    // ---- X . L = Z, column blocks descending ----
    int last = ((n - 1) / NB) * NB;
    for (int j0 = last; j0 >= 0; j0 -= NB) {
        int nb = std::min(NB, n - j0);
        int j1 = j0 + nb;
        if (j1 < n)
            PET_CHECK(dgemm_kk(m, nb, n - j1, B + j1, ldb, Lt + int64_t(j0) * lda + j1, lda, B + j0, ldb, -1.0,
                               1, st));
        trsm_rows_kernel<<<row_blocks, NB, TRSM_SMEM, st>>>(B, ldb, m, A, lda, j0, nb, invd, 1);
        PET_LAUNCH_CHECK();
    }
    return PET_OK;
}

int64_t spd_solve_work_doubles(int64_t n, int64_t lda) { return n * lda + round_up(n, 2) + 4; }

}  // namespace pet
