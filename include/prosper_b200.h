/*
 * prosper_b200.h -- C ABI of the B200-native truncated-EM (Expectation Truncation) engine.
 *
 * This is the drop-in boundary for ONE hot path of ml-uol/prosper: the per-iteration
 *     select_Hprimes -> E_step -> M_step
 * of the sparse-coding models in prosper/em/camodels/*_et.py.  The reference has no
 * native boundary (it is pure Python/NumPy behind the CAModel operator API), so the entry
 * points below are what a ctypes binding of that operator API needs; each one cites the
 * reference interface it replaces.  The reference-side stub is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success or a negative
 *     PET_E* code, with the text available from pet_last_error() (thread-local).
 *   - all floating point is IEEE float64, all matrices row-major; `ld*` are leading
 *     dimensions in ELEMENTS.
 *   - pointers named *_dev must be device pointers on the engine's device, pointers
 *     named *_host must be host pointers (pinned memory makes the copies asynchronous);
 *     pointers with neither suffix may be either (resolved with cudaPointerGetAttributes).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls
 *     enqueue work on it and return without synchronising unless they return host
 *     scalars/arrays (stated per function).
 *   - no CPU fallback exists: pet_create fails if no sm_100 device is present.
 */
#ifndef PROSPER_B200_H
#define PROSPER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PET_ABI_VERSION 1

/* error codes */
#define PET_OK          0
#define PET_EINVAL     -1   /* bad argument / unsupported configuration           */
#define PET_ECUDA      -2   /* CUDA runtime error (see pet_last_error)             */
#define PET_ENOMEM     -3   /* device allocation failed                            */
#define PET_ESTATE     -4   /* call order violated (e.g. E-step before set_data)   */
#define PET_ENUMERIC   -5   /* non-finite value where the reference asserts finite */

/* model kinds: the concrete CAModel subclasses (prosper/em/camodels/*_et.py) */
#define PET_MODEL_BSC   0   /* bsc_et.py  BSC_ET  */
#define PET_MODEL_MCA   1   /* mca_et.py  MCA_ET  */
#define PET_MODEL_MMCA  2   /* mmca_et.py MMCA_ET */
#define PET_MODEL_TSC   3   /* tsc_et.py  TSC_ET  */
#define PET_MODEL_DSC   4   /* dsc_et.py  DSC_ET  */
#define PET_MODEL_GSC   5   /* gsc_et.py  GSC     */

typedef struct pet_engine pet_engine;

/* Constructor arguments of CAModel.__init__(D, H, Hprime, gamma, ...)
 * (camodels/__init__.py:60-102); `states` is DSC_ET's `states=` (dsc_et.py:135) and is
 * fixed to {-1,0,1} for TSC (tsc_et.py:125). */
typedef struct pet_config {
    int32_t model;          /* PET_MODEL_*                                         */
    int32_t device;         /* CUDA device ordinal                                  */
    int64_t D, H, Hprime, gamma;
    int32_t n_states;       /* K: number of latent values (TSC/DSC), else 0         */
    const double *states;   /* K values, exactly one of them 0 (host pointer)       */
    int64_t chunk_rows;     /* datapoints per pipeline chunk; 0 = choose            */
} pet_config;

/* model_params dict of the reference: W is (D,H) row-major exactly like the NumPy array
 * (bsc_et.py:142 transposes internally, so do we); pi is a scalar (BSC/MCA/MMCA/TSC) or a
 * K-vector (DSC, dsc_et.py:503); mu is BSC's optional offset (bsc_et.py:145-149). */
typedef struct pet_params {
    const double *W;        /* (D,H), host or device                                */
    int64_t ldW;            /* >= H                                                 */
    const double *pi_host;  /* n_pi values                                          */
    int32_t n_pi;
    double sigma;
    const double *mu;       /* (D,) or NULL (= zeros), host or device               */
} pet_params;

/* The annealing values the hot path reads through anneal['...'] (bsc_et.py:152,187,247). */
typedef struct pet_anneal {
    double T;               /* anneal['T']                                          */
    double Ncut_factor;     /* anneal['Ncut_factor']                                */
    int32_t anneal_prior;   /* anneal['anneal_prior'] != 0                          */
} pet_anneal;

/* Everything M_step reduces over datapoints, laid out as ONE contiguous float64 buffer so
 * that the reference's 7-9 MPI collectives per iteration (bsc_et.py:225,258,266,373-374,
 * 387,417) become one all-reduce.  Offsets/sizes are returned by pet_stats_layout. */
typedef struct pet_stats_layout {
    int64_t total;          /* number of doubles                                    */
    int64_t off_Wp, rows_Wp, cols_Wp, ld_Wp;   /* numerator, stored (D+1, H): row D = sum_n <s> */
    int64_t off_Wq, rows_Wq, cols_Wq, ld_Wq;   /* (H,H) for BSC/TSC/DSC, (D,H) for MCA/MMCA   */
    int64_t off_scalars;    /* [0]=n_used [1]=sum_n log sum_c exp(logpj) [2]=sigma stat
                               [3..3+n_counts) = per-value activity counts                  */
    int64_t n_scalars;
} pet_stats_layout;

/* ---- lifecycle ------------------------------------------------------------------- */
int  pet_abi_version(void);
const char *pet_last_error(void);
int  pet_create(const pet_config *cfg, pet_engine **out);
void pet_destroy(pet_engine *e);

/* ---- state space (camodels/__init__.py:21-47, tsc_et.py:23-80, dsc_et.py:56-63) ---- */
int64_t pet_num_states(const pet_engine *e);    /* rows of state_matrix                    */
int64_t pet_num_columns(const pet_engine *e);   /* columns of logpj                        */
/* state_matrix as float64 (S,Hprime) into host memory, reference row order */
int  pet_state_matrix(const pet_engine *e, double *out_host);

/* ---- data ------------------------------------------------------------------------ */
/* Bind my_data['y'] (n,D).  The engine keeps its own padded device copy (ld rounded up,
 * one extra all-ones column used by the statistics GEMM).  With a host pointer (pinned for
 * full speed) nothing blocks the caller: the shard is uploaded chunk-wise on an internal
 * copy stream, a few chunks AHEAD of the first pass that consumes it, so the upload
 * overlaps that pass -- `y` must therefore stay valid and unchanged until the first
 * operator call after pet_set_data has completed.  Contiguous sources (ld == D) travel as
 * 1-D copies through staging slots and are padded on the device. */
int  pet_set_data(pet_engine *e, const double *y, int64_t n, int64_t ld, void *stream);
/* Chunk length policy of the NEXT pet_set_data (when pet_config.chunk_rows is 0): bytes of the
 * posterior matrix <S> per chunk.  Default 600 MB: long chunks amortise the wave tails of every
 * kernel and suit a shard that stays resident; a caller that re-uploads the shard for every
 * pass is bound by the upload and wants short chunks (128 MB) for a fine-grained overlap.
 * 0 restores the default. */
int  pet_set_chunk_target(pet_engine *e, int64_t bytes_of_posterior_per_chunk);
int64_t pet_num_data(const pet_engine *e);

/* ---- the three operators --------------------------------------------------------- */
/* select_Hprimes(model_params, data) -> data['candidates']  (bsc_et.py:98-115,
 * mca_et.py:88-111, mmca_et.py:95-124, tsc_et.py:142-212, dsc_et.py:347-410).
 * cand_out (n,Hprime) int64, host or device, may be NULL (kept internally). */
int  pet_select_hprimes(pet_engine *e, const pet_params *p, int64_t *cand_out, void *stream);

/* Override the internally kept candidates (my_data['candidates'] supplied by the caller). */
int  pet_set_candidates(pet_engine *e, const int64_t *cand, void *stream);

/* E_step(anneal, model_params, my_data) -> {'logpj': (n,C)}  (bsc_et.py:119-192, ...).
 * logpj_out (n,C) host or device. */
int  pet_e_step(pet_engine *e, const pet_anneal *a, const pet_params *p,
                double *logpj_out, int64_t ld_logpj, void *stream);

/* CAModel.inference (camodels/__init__.py:255-375) on the device, for one E-step's logpj (n,C)
 * in device memory: per datapoint the K most probable columns (idx_out (n,topK) int32, value
 * descending) with p_out (n,topK) = normalised log-posterior if logprob, else
 * exp(logpj - rowmax) exactly as the reference returns it (:309,:324), and -- binary layout
 * [null | singletons | states] only, m_out_dev may be NULL -- m_out (n,H) = marginal
 * LOG-probability of every cause (:336-342; the caller exponentiates, :371).  Needs bound data and candidates. */
int  pet_posterior_topk(pet_engine *e, const double *logpj_dev, int64_t ld_logpj, int32_t topK,
                        int32_t logprob, int32_t *idx_out_dev, double *p_out_dev,
                        double *m_out_dev, void *stream);

/* First half of M_step: per-datapoint log-denominators log sum_c exp(logpj) used by the
 * truncation rule (bsc_et.py:222,247-258) and by L (bsc_et.py:265).  `logpj` NULL = evaluate
 * the E-step on the fly (fused path, nothing of size n*C is materialised).
 * logdenom_out_dev (n,) device, may be NULL (kept internally). */
int  pet_log_denominators(pet_engine *e, const pet_anneal *a, const pet_params *p,
                          const double *logpj, int64_t ld_logpj, int32_t flags,
                          double *logdenom_out_dev, void *stream);
/* engine-owned (n,) device array the call above fills (valid until the next pet_set_data) */
const double *pet_log_denominators_ptr(const pet_engine *e);

/* `flags` of pet_log_denominators / pet_m_step_stats */
#define PET_PASS_SELECT        1  /* (re)run select_Hprimes inside this pass (fused step)      */
#define PET_PASS_REUSE_SCORES  2  /* same params as the previous pass: reuse its score matrix  */
#define PET_PASS_DEFER_STATS   4  /* pet_log_denominators: a pet_m_step_stats(use_cut, REUSE_SCORES) with the same params
                                   * follows -- evaluate the posterior once and park the per-datapoint statistics   */

/* k-th largest of n device doubles (replaces parallel.allsort(...)[-k], parallel.py:87-110).
 * Result written to *out_dev (device). */
int  pet_kth_largest(pet_engine *e, const double *vals_dev, int64_t n, int64_t k,
                     double *out_dev, void *stream);

/* Second half of M_step, local part: accumulate the packed sufficient statistics of this
 * rank's datapoints (bsc_et.py:334-366,395-415 and the model-specific equivalents).
 * use_cut != 0 keeps only datapoints whose log-denominator is >= *cut_dev (strict > for DSC,
 * dsc_et.py:832); requires pet_log_denominators to have run when the data is not re-evaluated.
 * `logpj` NULL = fused path.  stats_dev: pet_stats_layout.total doubles, overwritten. */
int  pet_m_step_stats(pet_engine *e, const pet_anneal *a, const pet_params *p,
                      const double *logpj, int64_t ld_logpj, int32_t flags,
                      int32_t use_cut, const double *cut_dev,
                      double *stats_dev, void *stream);
int  pet_stats_layout_get(const pet_engine *e, pet_stats_layout *out);

/* Second half of M_step, replicated part: from the (all-reduced) statistics produce the
 * new W (D,H) on the device (lstsq / pinv / element-wise update: bsc_et.py:377-380,
 * tsc_et.py:493, dsc_et.py:732-735, mca_et.py:343-348, mmca_et.py:383-394).  The scalar
 * updates (pi, sigma, L) are O(1) host arithmetic on stats[off_scalars..] and are done by
 * the caller exactly as the reference writes them.  W_new_dev (D,H) ld=H, device.
 * info_host[0] = number of pivots treated as zero (rank deficiency). Synchronises. */
int  pet_m_step_solve(pet_engine *e, const pet_params *p, const double *stats_dev,
                      double *W_new_dev, int32_t *info_host, void *stream);

/* ---- GSC (gsc_et.py): spike-and-slab model with its own parameter set and statistics --------- */
/* model_params of GSC: W (D,H), pi (H,), mu (H,), psi_sq (H,H), sigma_sq scalar / (D,) / (D,D)
 * (gsc_et.py:60-110).  sigma_sq_type: 0 'scalar', 1 'diagonal', 2 'full' (2 is rejected: not built). */
typedef struct pet_gsc_params {
    const double *W;            /* (D,H) host or device                                    */
    int64_t ldW;
    const double *pi_host;      /* (H,)                                                    */
    const double *mu_host;      /* (H,)                                                    */
    const double *psi_sq_host;  /* (H,H) row-major                                         */
    const double *sigma_sq_host;/* 1 or D values                                           */
    int32_t sigma_sq_type;
} pet_gsc_params;

/* Packed statistics of GSC.M_step (gsc_et.py:608-620,662-671,683-713), all-reduced as ONE buffer. */
typedef struct pet_gsc_layout {
    int64_t total, ld;
    int64_t off_A;        /* (D+1, H): Y^T.<sz>; row D = sum_n <sz>                        */
    int64_t off_Mssz;     /* (H,H): sum_n <s> (x) <sz>                                     */
    int64_t off_Mout;     /* (H,H): sum_n <sz> (x) <sz>                                    */
    int64_t off_ss;       /* (H,H): sum_n <s s^T>, off-diagonal part (diagonal = sum_s)    */
    int64_t off_szsz;     /* (H,H): sum_n <sz sz^T>, multi-cause part                      */
    int64_t off_sum_s;    /* (H,)                                                          */
    int64_t off_sum_sz2;  /* (H,): singleton part of the diagonal of sum_n <sz sz^T>       */
    int64_t off_ysq;      /* (D,): sum_n y_nd^2                                            */
    int64_t off_scalars;  /* [0] = datapoints of this rank                                 */
    int64_t off_yyT;      /* (D, ldY): sum_n y_n y_n^T, filled for sigma_sq_type 'full' only; ldY = ld_yyT */
    int64_t ld_yyT;
} pet_gsc_layout;
int  pet_gsc_layout_get(const pet_engine *e, pet_gsc_layout *out);

/* GSC.select_Hprimes (gsc_et.py:721-749): top-H' singleton marginal scores, sorted ascending.
 * cand_out (n,Hprime) int64, host or device, may be NULL. */
int  pet_gsc_select(pet_engine *e, const pet_gsc_params *p, int64_t *cand_out, void *stream);
/* GSC.E_step (gsc_et.py:401-580): dense posterior moments, written at row dst_rows[n] (the
 * reference returns them grouped by candidate set; NULL = identity).  Device outputs:
 * xpt_s, xpt_sz (n,H); xpt_ss, xpt_szsz (n,H,H). */
/* GSC.compute_lpj (gsc_et.py:811-944): preselection, then the log-joint of the null state, the H
 * singletons and the S multi-cause states WITHOUT annealing or the `tiny` clamp of the
 * E-step: logpj_dev (n, 1+H+S) device.  cand_out (n,Hprime) int64 host or device, may be NULL.
 * Feeds pet_posterior_topk for GSC's inference. */
int  pet_gsc_compute_lpj(pet_engine *e, const pet_gsc_params *p, double *logpj_dev, int64_t ld_logpj,
                         int64_t *cand_out, void *stream);
int  pet_gsc_e_step(pet_engine *e, const pet_anneal *a, const pet_gsc_params *p, const int64_t *dst_rows,
                    double *xpt_s_dev, double *xpt_ss_dev, double *xpt_sz_dev, double *xpt_szsz_dev, void *stream);
/* Fused select (flags & PET_PASS_SELECT) + E-step + local statistics; nothing of size n*H*H is
 * materialised.  stats_dev: pet_gsc_layout.total doubles, overwritten. */
int  pet_gsc_stats(pet_engine *e, const pet_anneal *a, const pet_gsc_params *p, int32_t flags,
                   double *stats_dev, void *stream);
/* out_dev[d] += sum over the datapoints kept by the truncation rule (all if !use_cut) of the
 * engine's copy of y[n,d] -- i.e. of y - mu for BSC, the shard is stored shifted by the mu in
 * force: `data_sum` of the mu update, bsc_et.py:282,422-428.  Uses the log-denominators of
 * the last pass. */
int  pet_data_sum(pet_engine *e, int32_t use_cut, const double *cut_dev, double *out_dev, void *stream);
/* out[c] += sum_r M[r][c] (used by the compat GSC M-step on caller-supplied moment tensors) */
int  pet_colsum(int64_t rows, int64_t cols, const double *M_dev, int64_t ld, double *out_dev, void *stream);

/* ---- synthetic data and initialisation on the device (engine-independent) ------------ */
/* model.generate_data (camodels/__init__.py:104-122, bsc_et.py:67-95, tsc_et.py:214-275,
 * dsc_et.py, mca_et.py:65-91): every latent s[n,h] takes values_host[k] with probability
 * probs_host[k] (K <= 16); y[n,:] = sum_h s_h W[:,h] (combine 0) or, per pixel, the entry
 * s_h W[d,h] of largest magnitude (combine 1, MCA/MMCA), plus N(0, sigma^2).  Counter-based
 * Philox RNG keyed by `seed` and the GLOBAL row index row0 + n, so shards generated by
 * different ranks or in pieces are identical to one big call.  s_idx_dev (n,H) int8
 * receives the index k of each latent's value, may be NULL.  W_dev is (D,H) row-major. */
int  pet_generate_data(int32_t combine, int64_t n, int64_t row0, int32_t D, int32_t H,
                       const double *W_dev, int64_t ldW, int32_t K, const double *values_host,
                       const double *probs_host, double sigma, uint64_t seed,
                       double *y_dev, int64_t ldy, int8_t *s_idx_dev, int64_t lds, void *stream);
/* X[i][j] = (row_base[i] or 0) + scale * N(0,1): W_init of standard_init
 * (camodels/__init__.py:224-226) and parameter noise (em/__init__.py:63-107) */
int  pet_normal_fill(double *X_dev, int64_t ld, int64_t rows, int64_t cols,
                     const double *row_base_dev, double scale, uint64_t seed, void *stream);
/* dst[i][:] = src[idx[i]][:] for i < n_sel: the random datapoint subset of CAModel.select_partial_data
 * (camodels/__init__.py:125-152) taken from a device-resident shard, no host round trip. */
int  pet_gather_rows(int64_t n_sel, int64_t n_src, int64_t cols, const double *src_dev, int64_t ld_src,
                     const int64_t *idx_dev, double *dst_dev, int64_t ld_dst, void *stream);
/* out[c] += sum_r (M[r][c] - mean[c])^2: data variance of standard_init (:220) */
int  pet_col_centered_sumsq(int64_t rows, int64_t cols, const double *M_dev, int64_t ld,
                            const double *mean_dev, double *out_dev, void *stream);

/* ---- mixture models (prosper/em/mixturemodels; contractions are pet_dgemm_kk / pet_dgemm_mn) ---- */
/* logpj[n][h] = beta (s1 T1[n][h] + s2 T2[n][h] + k[h]); post = exp(logpj) with the reference's clamps
 * (NaN -> tiny, < tiny -> tiny, inf -> max/H), rows normalised: MoG.py:208-218, MoP.py:165-175.
 * T2_dev may be NULL.  All device pointers. */
int  pet_mix_posterior(int64_t n, int32_t H, const double *T1_dev, const double *T2_dev, int64_t ldt,
                       double s1, double s2, const double *k_dev, double beta, double *logpj_dev,
                       int64_t ld_logpj, double *post_dev, int64_t ld_post, void *stream);
/* out[n * out_stride] = sum_d A[n][d] B[n][d] */
int  pet_rowdot(int64_t n, int32_t D, const double *A_dev, int64_t lda, const double *B_dev, int64_t ldb,
                double *out_dev, int64_t out_stride, void *stream);
/* row-wise element operations into a zero-padded (n, ldo) array: op 0: X^2; 1: X * w[n * w_stride];
 * 2: X - w[d]; 3: a / (sum_d X[n][d] + eps) * X + 1 (MoP.normalize, MoP.py:236-244) */
int  pet_rowop(int32_t op, int64_t n, int32_t D, const double *X_dev, int64_t ldx, const double *w_dev,
               int64_t w_stride, double a, double *out_dev, int64_t ldo, void *stream);

/* ---- building blocks exported for tests / benchmarks ------------------------------ */
/* C(M,N) = A.B with both operands K-contiguous: A(M,K) lda, B(N,K) ldb (FP64 tensor-core
 * tiles).  Device pointers, 16-byte aligned, even lda/ldb. */
int  pet_dgemm_kk(int64_t M, int64_t N, int64_t K, const double *A_dev, int64_t lda,
                  const double *B_dev, int64_t ldb, double *C_dev, int64_t ldc,
                  double alpha, double beta, void *stream);
/* C(M,N) = A^T.B with A(K,M) lda, B(K,N) ldb (reduction over rows, split-K inside;
 * workspace_dev holds splits*M*ldc doubles, query with C_dev == NULL -> returns splits). */
int  pet_dgemm_mn(int64_t M, int64_t N, int64_t K, const double *A_dev, int64_t lda,
                  const double *B_dev, int64_t ldb, double *C_dev, int64_t ldc,
                  int32_t accumulate, double *workspace_dev, int64_t workspace_doubles,
                  void *stream);
/* C(M,N) = A.B^T as pet_dgemm_kk (alpha 1, beta 0), computed on the INT8 tcgen05 tensor
 * cores by error-free slicing of both operands into `nslices` (6 or 7) int8 slices with
 * exact int32 accumulation in TMEM (Ozaki scheme; 7 slices: |err| <= ~1e-13 max|C|).
 * Slices both operands, then runs the GEMM `repeat` times (>= 1; for benchmarks).
 * Synchronises the stream. */
int  pet_ozaki_gemm_kk(int64_t M, int64_t N, int64_t K, const double *A_dev, int64_t lda,
                       const double *B_dev, int64_t ldb, double *C_dev, int64_t ldc,
                       int32_t nslices, int32_t repeat, void *stream);
/* Same arithmetic for C(M,N) = A^T.B with A (K,M) lda, B (K,N) ldb (reduction over rows,
 * as pet_dgemm_mn): column-wise slicing with a transposing store, split-K inside. */
int  pet_ozaki_gemm_mn(int64_t M, int64_t N, int64_t K, const double *A_dev, int64_t lda,
                       const double *B_dev, int64_t ldb, double *C_dev, int64_t ldc,
                       int32_t nslices, int32_t repeat, void *stream);
/* milliseconds per product of the last pet_ozaki_gemm_* call (events around the repeat loop) */
double pet_ozaki_last_ms(void);
/* Solve X.A = B for X with A (n,n) symmetric positive semi-definite (pivots below
 * tol are dropped, giving the minimum-norm behaviour of lstsq for dead units).
 * A is overwritten by its Cholesky factor; B (m,n) overwritten by X. Synchronises.
 * When B lies directly under A in one buffer (B_dev == A_dev + n*lda, ldb == lda) the
 * forward sweep is fused into the factorisation (32 launches less at n = 1000). */
int  pet_spd_solve_right(int64_t n, int64_t m, double *A_dev, int64_t lda,
                         double *B_dev, int64_t ldb, double *work_dev,
                         int32_t *info_host, void *stream);
int64_t pet_spd_solve_work_doubles(int64_t n, int64_t lda);   /* size of work_dev */

/* Device time per stage since pet_enable_timing(e,1), measured with CUDA events on the
 * caller's stream around every launch group: out[0..PET_N_STAGES-1] = total ms of
 * [0]=prepare (transpose+Gram+slicing of W) [1]=score GEMM [2]=state kernel (or the whole
 * posterior kernel for MCA/MMCA/GSC) [3]=statistics GEMM [4]=solve [5]=kth-largest
 * [6]=row kernel [7]=scale kernel [8]=int8 slicing of <s> for the statistics GEMM [9]=spare;
 * out[PET_N_STAGES..2*PET_N_STAGES-1] = how many spans each total sums.  Synchronises. */
#define PET_N_STAGES 10
/* which kernels run the score / statistics GEMMs for the bound shard: 0 = FP64 DMMA
 * (dgemm.cu), n > 0 = int8 tcgen05 with n slices per operand of the statistics GEMM (ozaki.cu) */
int32_t pet_gemm_path(const pet_engine *e);
/* slices per operand of the score GEMM y.W (bsc_et.py:107, default 6: 42 bits) and of the statistics GEMM
 * sum_n y <s>^T (bsc_et.py:366, default 7: 49 bits); 0, 0 on the FP64 DMMA path.  Environment: PET_OZAKI_SLICES=6|7
 * (both), PET_OZAKI_SLICES_SCORE, PET_OZAKI_SLICES_STATS. */
int32_t pet_gemm_slices(const pet_engine *e, int32_t *score, int32_t *stats);
/* Which kernel evaluates the multi-cause states of the fused BSC path (bsc_et.py:180-185, 349-366): mode 0 = automatic
 * (tensor cores once the state space has >= 256 states and a chunk of the shard fills half a wave of 128-datapoint
 * tiles), 1 = the scalar FP64 kernel, 2 = the int8 tensor-core kernel (binary states, H' <= 12, gamma <= 5; PET_EINVAL
 * otherwise).  pet_state_kernel_path reports the choice for the first chunk of the bound shard (1 or 2; before data
 * is bound: for a large shard).  The compat E_step / M_step that materialise logpj always use the scalar kernel. */
int  pet_set_state_kernel(pet_engine *e, int32_t mode);
int32_t pet_state_kernel_path(const pet_engine *e);
int  pet_stage_times_ms(pet_engine *e, double *out_host);
int  pet_enable_timing(pet_engine *e, int32_t on);
/* number of kernels launched by the engine since creation (bench.py's gpu_launches) */
int64_t pet_launch_count(const pet_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* PROSPER_B200_H */
