// Declarations of the MCA-ET / MMCA-ET kernels (see mca_kernel.cu).
#pragma once
#include <algorithm>

#include "gl_kernel.cuh"

namespace pet {

struct MCAArgs {
    int mmca;                      // 0 = MCA, 1 = MMCA
    int D, H, Hp, S, C, gamma;
    int ldH, ldY, ldD, ldc;        // ldc: stride of the staged candidate rows in shared memory
    const unsigned long long *states;   // binary records, unused member = Hp
    double rho, beta, pre1, pil_bar;
    int flags;                     // GLF_WRITE_LOGPJ | GLF_READ_LOGPJ | GLF_LSE_ONLY | GLF_USE_CUT
    int64_t n_rows, row0;
    const double *Y;               // chunk base (n_rows, ldY)
    const double *YW;              // chunk scores (n_rows, ldH)
    const double *yy;              // (n,) global index
    const double *wn2;             // (H,)
    const double *Wl, *Wr;         // (H, ldD) log|W| and (signed) |W|^rho
    const int *cand;               // (n, Hp) global index
    double *logpj; int64_t ld_logpj;
    double *lse;                   // (n,) annealed log-denominators
    const double *cut;
    double *Spost;                 // (n_rows, ldH) singles posterior of the chunk
    double *Wpm, *Wqm;             // (H, ldD) multi-cause numerators / denominators (atomics)
    double *scalars;               // [0]=n_used [1]=sum log sum exp(logpj) [2]=sigma stat [3]=pi stat
};

int launch_mca_kernel(const MCAArgs &a, int sm_count, cudaStream_t stream);
size_t mca_smem_bytes(const MCAArgs &a);
int launch_mca_tables(const double *Wt, int64_t ldk, int H, int D, double rho, int mmca, double *Wl, double *Wr,
                      int64_t ldD, double *wn2, cudaStream_t st);
int launch_mca_sim(const double *Y, int64_t ldY, int64_t rows, const double *W, int64_t ldW, int D, int H, double *sim,
                   int64_t ldH, cudaStream_t st);
int launch_mca_update(const double *A, int64_t ldA, const double *Wpm, const double *Wqm, int64_t ldD, const double *Wt,
                      int64_t ldk, int D, int H, int mmca, double tol, double *W_new, int64_t ldo, cudaStream_t st);

}  // namespace pet
