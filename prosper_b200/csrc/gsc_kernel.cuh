// Declarations of the GSC (spike-and-slab) kernels, see gsc_kernel.cu.
#pragma once
#include "gl_kernel.cuh"

namespace pet {

// per-cause constants of one iteration, each (H,)
struct GSCTables {
    const double *g;        // G'_hh = W_h^T Sigma^-1 W_h
    const double *ilam;     // 1 / (g + 1/psi_hh)          Lambda_h^-1 of the singleton state
    const double *lcdet;    // log psi_hh + log Lambda_h    (C_det of the singleton, gsc_et.py:502)
    const double *logit;    // log pi_h - log(1 - pi_h)
    const double *mu;       // mu_h
};

enum {
    GSCF_SELECT = 1 << 0,       // run the preselection (else candidates are given)
    GSCF_SELECT_ONLY = 1 << 1,
    GSCF_DENSE = 1 << 2,        // compat E_step: write dense xpt_* rows (at row dst[n])
    GSCF_STATS = 1 << 3,        // accumulate the M-step statistics
    GSCF_LOGPJ = 1 << 4,        // compute_lpj (gsc_et.py:811-944): write the un-annealed, un-clamped log-joints and stop
};

struct GSCArgs {
    GLStatic st;                 // binary state space + select tables (select_mode = SEL_ROW_MAX)
    GSCTables tb;
    int flags;
    double beta;
    int64_t n_rows, row0;
    const double *YW;            // (n_rows, ldH)  y^T Sigma^-1 W of the chunk
    const double *yyw;           // (n,) y^T Sigma^-1 y, global index
    const double *G;             // (H, ldH) W^T Sigma^-1 W
    const double *psi;           // (H, ldH) psi_sq
    int *cand;                   // (n, Hp) global index, sorted ascending
    // chunk outputs for the statistics GEMMs
    double *XS, *XSZ, *SZ2;      // (n_rows, ldH): <s>, <s z>, singleton part of diag <s z z>
    double *sum_ss, *sum_szsz;   // (H, ldH) atomic scatter targets (multi-cause blocks)
    // dense compat outputs (row index dst[n], or n when dst == NULL)
    const int64_t *dst;
    double *xpt_s, *xpt_sz;      // (n, H)
    double *xpt_ss, *xpt_szsz;   // (n, H, H)
    double *logpj; int64_t ld_logpj;   // GSCF_LOGPJ: (n, 1 + H + S), global row index
};

int launch_gsc_kernel(const GSCArgs &a, int gamma, int sm_count, cudaStream_t stream);
int launch_gsc_tables(const double *G, int64_t ldG, const double *psi, int64_t ldpsi, const double *pi, const double *mu,
                      int H, double *g, double *ilam, double *lcdet, double *logit, cudaStream_t st);
int launch_scale_rows(double *Wt_out, const double *Wt, int64_t ld, int H, int D, const double *bdiag, double bscalar,
                      cudaStream_t st);
int launch_weighted_rownorm(const double *Y, int64_t ldy, int64_t n, int D, const double *bdiag, double bscalar,
                            double *out, cudaStream_t st);

}  // namespace pet
