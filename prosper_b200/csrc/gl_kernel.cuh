// Shared declarations of the "Gaussian-linear" ET posterior kernel (BSC / TSC / DSC).
#pragma once
#include "common.cuh"

namespace pet {

constexpr int PET_MAXV = 6;        // max number of non-zero latent values (K-1)
constexpr int PET_MAXHP = 16;      // max H'
constexpr int PET_MAXG = 8;
constexpr int PET_GRP_LANES = 128;  // lanes of the state kernel that share one datapoint (also the owner count of the gather tables)        // max gamma (members per state record)

// selection rules of the reference's select_Hprimes variants
enum SelectMode {
    SEL_BSC = 0,    // cosine score, ascending output order            (bsc_et.py:110-112)
    SEL_NEGDIST,    // smallest ||W_h - y||^2, best first               (mmca_et.py:115-119)
    SEL_TSC,        // top-H' of the 2H signed singletons, duplicates   (tsc_et.py:198-210)
    SEL_DSC,        // top-H' distinct h over (K-1)H singletons         (dsc_et.py:398-408)
    SEL_GIVEN,      // scores precomputed per (n,h) in the row buffer, smallest first (MCA)
};

// kernel behaviour flags
enum {
    GLF_SELECT      = 1 << 0,   // run the top-H' selection (else use cand_io as input)
    GLF_WRITE_LOGPJ = 1 << 1,   // materialise logpj (compat E_step)
    GLF_READ_LOGPJ  = 1 << 2,   // take logpj from the caller (compat M_step)
    GLF_LSE_ONLY    = 1 << 3,   // stop after the log-denominator
    GLF_USE_CUT     = 1 << 4,   // skip datapoints below the truncation cut
    GLF_CUT_STRICT  = 1 << 5,   // '>' instead of '>=' (dsc_et.py:832)
    GLF_SELECT_ONLY = 1 << 6,   // stop after selection
    GLF_NO_SROW     = 1 << 8,   // the <s> chunk is never materialised: the singleton kernel only reduces, the state kernel
                                // leaves scl[n] = {e^(m1-m)/Z, candidate marginals} and gl_post_slice recomputes the
                                // singleton posteriors from the score row while it cuts them into int8 slices
    GLF_DEFER_STATS = 1 << 9,   // truncated iteration, first sweep: the state kernel evaluates the posterior once and parks the
                                // per-datapoint statistics (pair sums in `pairs`, scalar contributions in rs[5], rs[6]); once
                                // the cut is known gl_finalize_cut adds up the datapoints that stay and zeroes scl of the rest
    GLF_FOLD_SCALE  = 1 << 7,   // no scale kernel: the state kernel folds the candidate marginals into the un-normalised
                                // <s> row (divided by the row's scale); consumers multiply rows by scl[n][0] on load
};

struct GLStatic {                 // fixed per engine
    int H, Hp, S, C;
    int ldH;                      // leading dimension of YW / S / G / Wq rows
    int has_null;                 // column 0 = null state
    int n_blocks;                 // all-H singleton blocks (BSC 1, DSC K-1, TSC 0)
    double block_val[PET_MAXV];   // latent value of each block
    int block_vidx[PET_MAXV];     // which count accumulator it feeds
    int n_cnt;                    // number of non-zero latent values
    double vals[PET_MAXV];        // value LUT by vidx
    int zbase;                    // zero-entry count base for the state prior (H' TSC, H DSC, 0 BSC)
    int select_mode;
    int diag_from_colsum;         // binary: Wq diagonal = column sums of <s> (not scattered)
    const unsigned long long *states;   // S records: 8 x (pos | vidx<<4), 0xFF = unused
    const unsigned long long *inc_states;   // binary, <= 5 members: incremental records (statespace.cpp.inc) or null
    const unsigned short *entries;   // [(chunk*CH+i)*32 + lane] state ids of the pair-sum gather lists
    const unsigned short *chunk_tab; // [chunk*32 + lane] output id | 0x8000 on the last chunk of an output
    const unsigned int *direct;      // (output<<16 | state id) of single-entry lists
    int n_direct;
    int n_chunks, chunk_len;         // per lane; chunk_len is a multiple of 4
    int n_out;                       // pair-sum outputs: pairs * n_cnt^2 * n_g
    int n_g;                         // state sizes 2..n_g+1 carry pairs
    const int *single_idx;           // [Hp*n_cnt] in-table singleton state ids (TSC) or -1
    int binary;                      // BSC-style state space: values {0,1}, states ordered by size
    int size_start[PET_MAXG + 2];    // binary: index of the first state with g members (g = 2..gamma), then S
};

struct GLIter {                   // per call
    double beta, pre1;
    int anneal_prior;
    double prior_null;
    double prior_block[PET_MAXV];
    double lp[PET_MAXV];          // log-prior weight per non-zero value
    double lp0;                   // log-prior weight of a zero entry
    double sel_prior[PET_MAXV];   // TSC/DSC selection: per-block log prior
};

struct GLArgs {
    GLStatic st;
    GLIter it;
    int flags;
    int64_t n_rows;               // rows in this chunk
    int64_t row0;                 // global index of the first row (for per-datapoint arrays)
    const double *YW;             // (n_rows, ldH) scores of this chunk
    const double *yy;             // (n,) squared norms (global index)
    const double *wn2;            // (H,)  ||W_h||^2
    const double *invn;           // (H,)  1/||W_h||
    const double *wmu;            // (H,)  W_h . mu (BSC selection scores the un-shifted datapoint); NULL when mu = 0
    const double *G;              // (H, ldH) Gram matrix
    const double *state_prior;    // (S,) log-prior of every multi-state for this iteration
    int *cand;                    // (n, Hp) global index
    double *logpj; int64_t ld_logpj;   // (n, C) global index (read or write)
    double *lse;                  // (n,) global index (written, or read when GLF_USE_CUT)
    const double *cut;            // device scalar
    double *S;                    // (n_rows, ldH) chunk posterior matrix <s>
    double *S2;                   // (n_rows, ldH) second moments of the singleton blocks (DSC) or NULL
    double *Wq;                   // (H, ldH) atomic scatter target
    double *scalars;              // [0]=n_used [1]=sum lse [2]=sigma stat [3..]=counts
    double *rs;                   // (n, 4+PET_MAXV) row-kernel partial sums {m1, Z1, sig1, -, cnt1[..]}, global index
    double *ywc;                  // (n, Hp) scores of the candidates, global index
    double *scl;                  // (n, 1+PET_MAXHP) {scale of the singleton row, candidate marginals}, global index
    int *tile_counter;            // tensor-core state kernel: next 128-datapoint tile to hand out (zeroed before the launch)
    double *pairs;                // GLF_DEFER_STATS: (H'(H'-1)/2, pairs_ld) normalised pair sums <s_j s_k>, feature major, global index
    int64_t pairs_ld;
};

int launch_gl_kernel(const GLArgs &a, int gamma, bool binary, int sm_count, cudaStream_t st);
int launch_gl_row(const GLArgs &a, int sm_count, cudaStream_t st);
int launch_gl_state(const GLArgs &a, int gamma, bool binary, int sm_count, cudaStream_t st);
int launch_gl_scale(const GLArgs &a, cudaStream_t st);
// second half of a GLF_DEFER_STATS evaluation (bsc_et.py:250-257, 349-366): datapoints at or above the cut scatter their parked
// pair sums into Wq and add their scalars; the others get scl = 0, so the slicer gives them an all-zero <s> row
int launch_gl_finalize_cut(const GLArgs &a, cudaStream_t st);
// <s>[n,h] = exp(F_h - m1[n]) scl[n][0] (+ scl[n][1+j] if h = cand[n][j]) cut into ns signed 7-bit slices relative to the
// fixed scale 1 (a posterior mean lies in [0, 1]: slice 0 <= 64, the others are UNSIGNED 7-bit digits 0..127, so the
// int32 accumulators of ozaki_gemm hold K <= 37 000 terms per split) and stored transposed, out[t][h][r] (r < Kp, zero
// beyond n_rows): the B operand of the statistics GEMM straight from the score rows.  scale_out[h] = 1.
int launch_gl_post_slice(const GLArgs &a, int ns, int Kp, int8_t *out, int64_t row_stride, int64_t slice_stride, double *scale_out,
                         cudaStream_t st);
size_t gl_smem_bytes(const GLStatic &s, int warps);
int gl_pick_warps(const GLStatic &s);
int launch_state_prior(const GLStatic &st, const GLIter &it, double *out, cudaStream_t stream);

}  // namespace pet
