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

// Cholesky of the nb x nb diagonal block at (j0,j0); writes L (lower, zero upper) back, and the inverse of the
// triangular block both plain and transposed (NB x NB, ld NB) so that every triangular solve of the blocked algorithm
// becomes a DMMA GEMM:   x . L_jj^T = r  <=>  x = r . inv(L_jj)^T .
// A pivot below tol is dropped: its row/column of L and of inv(L) are zero (minimum-norm behaviour for dead units).
//
// The block T and its inverse X live in REGISTERS: 256 threads, thread (pi, pj) owns the 4 x 4 patch at rows 4pi..,
// columns 4pj.. of both.  The kernel is a chain of dependent instructions (two warps per scheduler), so it is written for
// few instructions and few barriers: FOUR columns per step and two barriers per step --
//   1. the owner of the diagonal patch factors it and inverts the 4 x 4 factor (M) in registers;
//   2. the patches below it become L_p = T_p . M^T, the patches of X's row block become M . X_C; both are published;
//   3. every remaining patch takes its rank-4 update  T -= L_p(rows) . L_p(cols)^T ,  X -= L_p(rows) . X_C .
// X starts as the identity and receives the row operations that reduce L to the identity (forward elimination of [L | I]).
__global__ void __launch_bounds__(256) potrf_block_kernel(double *A, int64_t lda, int j0, int nb, int n,
                                                          double *invd, double *scal, double *Linv, double *LinvT) {
    __shared__ __align__(16) double dinv_s[2][16];          // M = inverse of the 4 x 4 diagonal factor, row-major
    __shared__ __align__(16) double lp_s[2][4][NB];          // L_p: the four new columns of L (column-major: conflict-free rows)
    __shared__ __align__(16) double xc_s[2][4][NB];          // X_C: the four finished rows of the inverse
    const int pi = threadIdx.x & 15, pj = threadIdx.x >> 4;  // a warp holds two patch columns: conditions on pj are warp-uniform
    const int r0 = 4 * pi, c0 = 4 * pj;
    const double tol = scal[0] * scal[2];
    double t[4][4], x[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + i, cc = c0 + j;
            t[i][j] = (r < nb && cc <= r) ? A[int64_t(j0 + r) * lda + j0 + cc] : 0.0;
            x[i][j] = (r == cc) ? 1.0 : 0.0;
        }
    const int nsteps = (nb + 3) >> 2;
    for (int C = 0; C < nsteps; ++C) {
        const int b = C & 1;
        // ---- 1. diagonal patch: 4 x 4 Cholesky and the inverse of its factor ----
        if (pi == C && pj == C) {
            double iv[4];
            int ndrop = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double d = t[c][c];
                const bool valid = 4 * C + c < nb;
                const bool keep = valid && d > tol;
                const double inv = keep ? rsqrt(d) : 0.0;
                iv[c] = inv;
                if (valid) {
                    invd[j0 + 4 * C + c] = inv;
                    ndrop += keep ? 0 : 1;
                }
                t[c][c] = d * inv;
#pragma unroll
                for (int r = c + 1; r < 4; ++r) t[r][c] *= inv;
#pragma unroll
                for (int cc = c + 1; cc < 4; ++cc)
#pragma unroll
                    for (int r = cc; r < 4; ++r) t[r][cc] -= t[r][c] * t[cc][c];
            }
            if (ndrop) atomicAdd(&scal[1], double(ndrop));
            double M[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int r = 0; r < 4; ++r) M[r][c] = 0.0;
                M[c][c] = iv[c];
#pragma unroll
                for (int r = c + 1; r < 4; ++r) {
                    double sum = 0.0;
#pragma unroll
                    for (int k = c; k < r; ++k) sum += t[r][k] * M[k][c];
                    M[r][c] = -iv[r] * sum;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) dinv_s[b][4 * r + c] = M[r][c];
        }
        __syncthreads();
        // ---- 2. the new columns of L below the diagonal patch, the finished rows of the inverse ----
        if ((pj == C && pi > C) || (pi == C && pj <= C)) {
            double M[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) M[r][c] = dinv_s[b][4 * r + c];
            if (pi > C) {                                    // L_p = T_p . M^T
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double l[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        double sum = t[i][0] * M[k][0];
#pragma unroll
                        for (int q = 1; q <= k; ++q) sum += t[i][q] * M[k][q];
                        l[k] = sum;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) { t[i][k] = l[k]; lp_s[b][k][r0 + i] = l[k]; }
                }
            } else {                                         // X_C = M . X_C
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double xn[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        double sum = M[i][0] * x[0][j];
#pragma unroll
                        for (int q = 1; q <= i; ++q) sum += M[i][q] * x[q][j];
                        xn[i] = sum;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) { x[i][j] = xn[i]; xc_s[b][i][c0 + j] = xn[i]; }
                }
            }
        }
        __syncthreads();
        // ---- 3. rank-4 updates of everything below ----
        if (pi > C && pj <= pi) {
            double lr[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) lr[i][k] = lp_s[b][k][r0 + i];
            if (pj > C) {                                    // T[r][cc] -= sum_k L[r][k] L[cc][k]
                double lc[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int k = 0; k < 4; ++k) lc[j][k] = lp_s[b][k][c0 + j];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) t[i][j] -= lr[i][k] * lc[j][k];
            } else {                                         // X[r][q] -= sum_k L[r][k] X_C[k][q]
                double xr[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int j = 0; j < 4; ++j) xr[k][j] = xc_s[b][k][c0 + j];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[i][j] -= lr[i][k] * xr[k][j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + i, cc = c0 + j;
            if (r < nb && cc < nb) A[int64_t(j0 + r) * lda + j0 + cc] = (cc <= r) ? t[i][j] : 0.0;
            const double xv = (cc <= r) ? x[i][j] : 0.0;
            Linv[r * NB + cc] = xv;
            LinvT[cc * NB + r] = xv;
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
    maxdiag_kernel<<<1, 256, 0, st>>>(A, lda, n, scal, tol_scale);
    PET_LAUNCH_CHECK();
    // B stacked under A with the same leading dimension: the rows of B ride along as extra rows of every panel, which IS
    // the forward sweep Z . L^T = B (the panel product gives Z_j, the trailing update removes Z_j . L[j1:, j0:j1]^T from the
    // remaining columns) -- 32 launches of the dependent chain less than a separate sweep.
    const bool stacked = m > 0 && B == A + n64 * lda && ldb == lda;
    const int64_t extra = stacked ? m : 0;
    // ---- factor ----
    for (int j0 = 0, jb = 0; j0 < n; j0 += NB, ++jb) {
        int nb = std::min(NB, n - j0);
        double *Li = Linv + int64_t(jb) * NB * NB;
        potrf_block_kernel<<<1, 256, 0, st>>>(A, lda, j0, nb, n, invd, scal, Li, LinvT + int64_t(jb) * NB * NB);
        PET_LAUNCH_CHECK();
        int j1 = j0 + nb;
        int64_t rows = n - j1;
        double *panel = A + int64_t(j1) * lda + j0;
        // panel: L[j1:, j0:j1] = A[j1:, j0:j1] . inv(L_jj)^T   (in place: one column tile, rows are CTA-private)
        if (rows + extra > 0) PET_CHECK(dgemm_kk(rows + extra, nb, nb, panel, lda, Li, NB, panel, lda, 1.0, 0, st));
        // trailing: A[j1:, j1:] -= L[j1:, j0:j1] . L[j1:, j0:j1]^T
        if (rows > 0)
            PET_CHECK(dgemm_kk(rows + extra, rows, nb, panel, lda, panel, lda, A + int64_t(j1) * lda + j1, lda, -1.0, 1, st));
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
    for (int j0 = 0, jb = 0; j0 < n && !stacked; j0 += NB, ++jb) {
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
