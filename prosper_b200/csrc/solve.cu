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

// Unblocked Cholesky of the nb x nb diagonal block at (j0,j0); writes L (lower, zero upper) back, and the
// inverse of the triangular block both plain and transposed (NB x NB, ld NB) so that every triangular solve of
// the blocked algorithm becomes a DMMA GEMM:   x . L_jj^T = r  <=>  x = r . inv(L_jj)^T .
// A pivot below tol is dropped: its row/column of L and of inv(L) are zero (minimum-norm behaviour for dead units).
__global__ void __launch_bounds__(256) potrf_block_kernel(double *A, int64_t lda, int j0, int nb, int n,
                                                          double *invd, double *scal, double *Linv, double *LinvT) {
    extern __shared__ __align__(16) double potrf_smem[];
    double (*T)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potrf_smem);
    double (*X)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potrf_smem + NB * (NB + 1));
    double *dinv = potrf_smem + 2 * NB * (NB + 1);
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
    const double tol = scal[0] * scal[2];
    for (int r = ty; r < NB; r += 4) {
        T[r][tx] = (r < nb && tx < nb && tx <= r) ? A[int64_t(j0 + r) * lda + j0 + tx] : 0.0;
        X[r][tx] = 0.0;
    }
    __syncthreads();
    // X starts as the identity and receives the same row operations that reduce L to the identity (forward elimination
    // of [L | I]), one column of L per step: row c is scaled by 1/L_cc, then rows r > c lose L[r][c] times row c.  The
    // inverse is complete when the factorisation is; its updates ride on the trailing update's barriers.
    if (threadIdx.x < NB) X[threadIdx.x][threadIdx.x] = 1.0;
    __syncthreads();
    for (int c = 0; c < nb; ++c) {
        // every thread derives 1/L_cc from the pivot itself (one barrier less than broadcasting it through shared memory);
        // the diagonal entry is overwritten only after the barrier, nobody reads it in the second phase
        const double d = T[c][c];
        const bool keep = d > tol;
        const double inv = keep ? rsqrt(d) : 0.0;
        if (threadIdx.x == 0) {
            dinv[c] = inv;
            invd[j0 + c] = inv;
            if (!keep) atomicAdd(&scal[1], 1.0);
        }
        if (ty == 0 && tx > c && tx < nb) T[tx][c] *= inv;
        if (ty == 1 && tx <= c) X[c][tx] *= inv;              // row c of the inverse is final
        __syncthreads();
        if (threadIdx.x == 0) T[c][c] = d * inv;
        // trailing update of the lower triangle: T[r][cc] -= L[r][c] * L[cc][c], c < cc <= r
        // (all loads first, then the stores: a load after a possibly aliasing shared store would serialise the loop)
        const int cc = c + 1 + tx;
        const double lcc = (cc < nb) ? T[cc][c] : 0.0;
        const double xck = (tx <= c) ? X[c][tx] : 0.0;         // inverse: X[r][k] -= L[r][c] * X[c][k] for r > c, k <= c
        double lr[NB / 4], tv[NB / 4], xv[NB / 4];
#pragma unroll
        for (int i = 0; i < NB / 4; ++i) {
            const int r = c + 1 + ty + 4 * i;
            const bool in = r < nb;
            lr[i] = in ? T[r][c] : 0.0;
            tv[i] = (in && cc <= r) ? T[r][cc] : 0.0;
            xv[i] = (in && tx <= c) ? X[r][tx] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < NB / 4; ++i) {
            const int r = c + 1 + ty + 4 * i;
            if (r < nb) {
                if (cc <= r) T[r][cc] = tv[i] - lr[i] * lcc;
                if (tx <= c) X[r][tx] = xv[i] - lr[i] * xck;
            }
        }
        __syncthreads();
    }
    __syncthreads();
    for (int r = ty; r < NB; r += 4) {
        if (r < nb && tx < nb) A[int64_t(j0 + r) * lda + j0 + tx] = T[r][tx];
        Linv[r * NB + tx] = X[r][tx];
        LinvT[r * NB + tx] = X[tx][r];
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
// work: n*lda (L^T) + n (invd) + 4 (scalars) + 2 * nblocks * NB*NB (inverse diagonal blocks, plain and transposed).
int spd_solve_right(int64_t n64, int64_t m, double *A, int64_t lda, double *B, int64_t ldb, double *work,
                    cudaStream_t st) {
    const int n = int(n64);
    const int nblk = (n + NB - 1) / NB;
    static double tol_scale = []() { const char *e = getenv("PET_PIVOT_TOL"); return e ? atof(e) : 1.1102230246251565e-16; }();
    double *Lt = work;
    double *invd = work + n64 * lda;
    double *scal = invd + round_up(n64, 2);
    double *Linv = scal + 4;
    double *LinvT = Linv + int64_t(nblk) * NB * NB;
    constexpr size_t POTRF_SMEM = (2 * NB * (NB + 1) + NB) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        PET_CUDA(cudaFuncSetAttribute(potrf_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(POTRF_SMEM)));
        configured = true;
    }
    maxdiag_kernel<<<1, 256, 0, st>>>(A, lda, n, scal, tol_scale);
    PET_LAUNCH_CHECK();
    // ---- factor ----
    for (int j0 = 0, jb = 0; j0 < n; j0 += NB, ++jb) {
        int nb = std::min(NB, n - j0);
        double *Li = Linv + int64_t(jb) * NB * NB;
        potrf_block_kernel<<<1, 256, POTRF_SMEM, st>>>(A, lda, j0, nb, n, invd, scal, Li, LinvT + int64_t(jb) * NB * NB);
        PET_LAUNCH_CHECK();
        int j1 = j0 + nb;
        if (j1 < n) {
            int64_t rows = n - j1;
            double *panel = A + int64_t(j1) * lda + j0;
            // panel: L[j1:, j0:j1] = A[j1:, j0:j1] . inv(L_jj)^T   (in place: one column tile, rows are CTA-private)
            PET_CHECK(dgemm_kk(rows, nb, nb, panel, lda, Li, NB, panel, lda, 1.0, 0, st));
            // trailing: A[j1:, j1:] -= L[j1:, j0:j1] . L[j1:, j0:j1]^T
            PET_CHECK(dgemm_kk(rows, rows, nb, panel, lda, panel, lda, A + int64_t(j1) * lda + j1, lda, -1.0, 1, st));
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
    // Both sweeps are RIGHT-LOOKING: as soon as a block column of the solution is known, all the remaining columns are
    // updated with it in ONE wide GEMM (K = 64, hundreds of tiles) -- the left-looking form has a growing K on a dozen tiles
    // and is latency bound.
    // ---- Z . L^T = B, column blocks ascending ----
    for (int j0 = 0, jb = 0; j0 < n; j0 += NB, ++jb) {
        int nb = std::min(NB, n - j0);
        int j1 = j0 + nb;
        PET_CHECK(dgemm_kk(m, nb, nb, B + j0, ldb, Linv + int64_t(jb) * NB * NB, NB, B + j0, ldb, 1.0, 0, st));
        if (j1 < n)      // B[:, j1:] -= Z_j . L[j1:, j0:j1]^T
            PET_CHECK(dgemm_kk(m, n - j1, nb, B + j0, ldb, A + int64_t(j1) * lda + j0, lda, B + j1, ldb, -1.0, 1, st));
    }
    // ---- X . L = Z, column blocks descending ----
    for (int jb = nblk - 1; jb >= 0; --jb) {
        int j0 = jb * NB;
        int nb = std::min(NB, n - j0);
        PET_CHECK(dgemm_kk(m, nb, nb, B + j0, ldb, LinvT + int64_t(jb) * NB * NB, NB, B + j0, ldb, 1.0, 0, st));
        if (j0 > 0)      // B[:, :j0] -= X_j . L[j0:j1, :j0]   (L^T rows are K-contiguous)
            PET_CHECK(dgemm_kk(m, j0, nb, B + j0, ldb, Lt + j0, lda, B, ldb, -1.0, 1, st));
    }
    return PET_OK;
}

int64_t spd_solve_work_doubles(int64_t n, int64_t lda) {
    return n * lda + round_up(n, 2) + 4 + 2 * ceil_div(n, NB) * NB * NB;
}

}  // namespace pet
