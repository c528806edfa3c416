// C-ABI engine: owns the device-resident layout of the training shard and orchestrates the
// per-iteration pipeline  prepare -> [chunk: score GEMM -> posterior kernel -> statistics GEMM]
// -> (all-reduce by the caller) -> solve.   See include/prosper_b200.h for the contract.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <vector>

#include "common.cuh"
#include "gl_kernel.cuh"
#include "gl_state_tc.cuh"
#include "mca_kernel.cuh"
#include "gsc_kernel.cuh"
#include "ozaki.cuh"

namespace pet {

thread_local std::string g_error;
thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

// kernels / launchers implemented in the other translation units
int dgemm_kk(int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B, int64_t ldb,
             double *C, int64_t ldc, double alpha, int accumulate, cudaStream_t st);
int dgemm_mn(int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B, int64_t ldb,
             double *C, int64_t ldc, int accumulate, double *work, int64_t work_doubles, int sm_count,
             cudaStream_t st);
int dgemm_mn_splits(int64_t M, int64_t N, int64_t K, int sm_count);
int spd_solve_right(int64_t n, int64_t m, double *A, int64_t lda, double *B, int64_t ldb, double *work,
                    cudaStream_t st);
int64_t spd_solve_work_doubles(int64_t n, int64_t lda);
int kth_largest(const double *vals, int64_t n, int64_t k, double *out, unsigned long long *state, int sm_count,
                cudaStream_t st);
int launch_transpose_w(double *Wt, int64_t ldk, const double *W, int64_t ldw, int D, int H, cudaStream_t st);
int launch_gram_diag(const double *G, int64_t ldg, int H, double *wn2, double *invn, cudaStream_t st);
int launch_subtract_mu(double *Y, int64_t ldy, int64_t n, int D, const double *mu, cudaStream_t st);
int launch_wdotmu(const double *Wt, int64_t ldk, int H, int D, const double *mu, double *out, cudaStream_t st);
int launch_rownorm_pad(const double *src, int64_t ld_src, double *Y, int64_t ldy, int64_t n, int D, double *yy, cudaStream_t st);
int launch_cand_to_i64(int64_t *out, const int *in, int64_t count, cudaStream_t st);
int launch_cand_from_i64(int *out, const int64_t *in, int64_t count, int H, cudaStream_t st);
int launch_add_diag(double *Wq, int64_t ld, const double *colsum, int H, cudaStream_t st);
int launch_colsum(double *out, const double *M, int64_t ld, int64_t rows, int cols, cudaStream_t st);
int launch_colsumsq(double *out, const double *M, int64_t ld, int64_t rows, int cols, cudaStream_t st);
int launch_colsum_kept(double *out, const double *M, int64_t ld, int64_t rows, int cols, const double *lse, const double *cut,
                       int strict, cudaStream_t st);

int launch_infer(const GLStatic &st, int C, int binary_layout, const int *cand, const double *logpj, int64_t ld, int64_t n,
                 int topK, int logprob, int *idx_out, double *p_out, double *m_out, int sm_count, cudaStream_t stream);

#include "statespace.cpp.inc"

static bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

enum Stage { ST_PREPARE = 0, ST_SCORE, ST_POST, ST_STATS, ST_SOLVE, ST_KSEL, ST_ROW, ST_SCALE, ST_SLICE, ST_SPARE, ST_COUNT };
static_assert(ST_COUNT == PET_N_STAGES, "stage table");

struct StageTimer {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Span { int stage; cudaEvent_t a, b; };
    std::vector<Span> spans;
    double totals[ST_COUNT] = {};
    int counts[ST_COUNT] = {};
    cudaEvent_t get() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
    void begin(int stage, cudaStream_t st) {
        if (!on) return;
        Span s{stage, get(), get()};
        cudaEventRecord(s.a, st);
        spans.push_back(s);
    }
    void end(cudaStream_t st) {
        if (!on) return;
        cudaEventRecord(spans.back().b, st);
    }
    void collect() {
        for (auto &s : spans) {
            cudaEventSynchronize(s.b);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, s.a, s.b);
            totals[s.stage] += ms;
            counts[s.stage] += 1;
        }
        spans.clear();
        used = 0;
    }
    void reset() { collect(); for (int i = 0; i < ST_COUNT; ++i) { totals[i] = 0; counts[i] = 0; } }
};

}  // namespace pet

using namespace pet;

struct pet_engine {
    int model = 0, device = 0, sm_count = 148;
    int D = 0, H = 0, Hp = 0, gamma = 0, K = 0, k0 = 0;
    std::vector<double> values;     // latent values (K), binary: {0,1}
    bool binary = true;
    StateSpace ss;
    int64_t C = 0;
    GLStatic gls{};
    int64_t ldY = 0, ldH = 0, chunk_rows = 0, chunk_cfg = 0, chunk_target_bytes = int64_t(600) << 20;
    std::vector<int64_t> chunk_start;        // rows [chunk_start[c], chunk_start[c+1]) form chunk c (each at most chunk_rows long)

    // device-resident shard
    double *Y = nullptr; int64_t n = 0, n_cap = 0;
    double *yy = nullptr; int *cand = nullptr; double *lse = nullptr;
    double *rs = nullptr, *ywc = nullptr, *scl = nullptr;      // per-datapoint records exchanged by the posterior kernels
    // single-evaluation truncated iteration (GLF_DEFER_STATS): parked pair sums (H'(H'-1)/2, pairs_ld), which chunks of the
    // last log-denominator sweep parked their statistics, and whether that record is still current
    double *pairs = nullptr; int64_t pairs_ld = 0;
    int *tile_counter = nullptr;                               // dynamic tile hand-out of the tensor-core state kernel
    std::vector<char> chunk_deferred; bool defer_valid = false;
    bool yy_valid = false, wmu_nonzero = false; int cand_state = 0;
    std::vector<double> mu_applied;

    // per-iteration
    double *Wt = nullptr, *G = nullptr, *wn2 = nullptr, *invn = nullptr, *Wtmp = nullptr, *mu_dev = nullptr, *wmu = nullptr, *mu_full = nullptr;
    double *YW = nullptr; int64_t yw_rows = 0; bool yw_all = false;
    double *Sbuf = nullptr, *S2buf = nullptr;
    double *gemm_work = nullptr; int64_t gemm_work_doubles = 0;
    double *solveA = nullptr, *solveB = nullptr, *solve_work = nullptr;
    double *s2sum = nullptr;
    double *Wl = nullptr, *Wr = nullptr, *simbuf = nullptr; int64_t ldD = 0;   // MCA/MMCA tables
    double *Wt2 = nullptr, *gsc_tab = nullptr, *psi_dev = nullptr, *bdiag = nullptr, *Bfull = nullptr, *gsc_T = nullptr, *XSZ = nullptr, *SZ2 = nullptr, *yyw = nullptr;   // GSC
    int64_t *dst_dev = nullptr; int64_t dst_cap = 0;
    const double *Wsrc = nullptr; int64_t Wsrc_ld = 0;                          // W (D,H) on the device for this call
    double *stage_logpj = nullptr; int64_t stage_logpj_doubles = 0;
    int64_t *stage_i64 = nullptr; int64_t stage_i64_count = 0;
    unsigned long long *ksel_state = nullptr;
    unsigned long long *d_states = nullptr, *d_inc = nullptr; unsigned short *d_entries = nullptr, *d_chunk = nullptr; unsigned int *d_direct = nullptr;
    int *d_single = nullptr; double *d_state_prior = nullptr;
    // multi-cause states on the tensor cores (gl_state_tc.cu): constant membership tables; tc_mode 0 = automatic
    // (large state spaces), 1 = never, 2 = whenever the state space is supported
    GLTcHost tc_host; uint8_t *d_tc_fwd = nullptr, *d_tc_rev = nullptr; bool tc_ok = false; int tc_mode = 0;

    // int8-sliced operands of the two large GEMMs (ozaki.cu); oz_on = buffers present and the path selected
    bool oz_want = false, oz_on = false; int oz_ns = 7, oz_ns2 = 7, oz_kpd = 0, oz_splits = 1; int64_t oz_rows = 0, oz_ldT = 0;
    int8_t *ozY = nullptr, *ozYT = nullptr, *ozW = nullptr, *ozS = nullptr;
    double *ozYs = nullptr, *ozYTs = nullptr, *ozWs = nullptr, *ozSs = nullptr, *oz_slabs = nullptr;
    unsigned long long *oz_colmax = nullptr;

    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> chunk_ready; bool upload_pending = false;
    // host shards are uploaded lazily, a few chunks ahead of the sweep that consumes them (so that the small
    // per-iteration uploads are not queued behind the whole shard on the copy engine); contiguous sources go
    // through 1-D copies into staging slots (pitched 2-D DMA is ~35% slower) and are expanded on the device
    const double *up_src = nullptr; int64_t up_ld = 0, up_copied = 0, up_expanded = 0, up_gran = 16384; bool up_staged = false;
    std::vector<double *> up_slots; std::vector<cudaEvent_t> up_frees;     // ring of staging slots, one event per slot
    cudaEvent_t compute_done = nullptr; bool compute_done_valid = false;
    StageTimer timer;
    int64_t launches0 = 0;
};

static int free_dev(void *p) { if (p) cudaFree(p); return 0; }

template <typename T>
static int dev_alloc(T **p, int64_t count) {
    *p = nullptr;
    if (count <= 0) count = 1;
    cudaError_t e = cudaMalloc((void **)p, size_t(count) * sizeof(T));
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of %lld bytes failed: %s", (long long)(count * (int64_t)sizeof(T)), cudaGetErrorString(e));
        return PET_ENOMEM;
    }
    return PET_OK;
}

extern "C" int pet_abi_version(void) { return PET_ABI_VERSION; }
extern "C" const char *pet_last_error(void) { return g_error.c_str(); }

extern "C" void pet_destroy(pet_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    free_dev(e->Y); free_dev(e->yy); free_dev(e->cand); free_dev(e->lse); free_dev(e->rs); free_dev(e->ywc); free_dev(e->scl);
    free_dev(e->pairs);
    if (e->tile_counter) cudaFree(e->tile_counter);
    free_dev(e->Wt); free_dev(e->G); free_dev(e->wn2); free_dev(e->invn); free_dev(e->Wtmp); free_dev(e->mu_dev); free_dev(e->wmu); free_dev(e->mu_full);
    free_dev(e->YW); free_dev(e->Sbuf); free_dev(e->S2buf); free_dev(e->gemm_work);
    free_dev(e->solveA); free_dev(e->solve_work); free_dev(e->s2sum);
    free_dev(e->stage_logpj); free_dev(e->stage_i64); free_dev(e->ksel_state);
    free_dev(e->Wl); free_dev(e->Wr); free_dev(e->simbuf);
    free_dev(e->Bfull); free_dev(e->gsc_T); free_dev(e->Wt2); free_dev(e->gsc_tab); free_dev(e->psi_dev); free_dev(e->bdiag); free_dev(e->XSZ); free_dev(e->SZ2); free_dev(e->yyw); free_dev(e->dst_dev);
    free_dev(e->ozY); free_dev(e->ozYT); free_dev(e->ozW); free_dev(e->ozS); free_dev(e->ozYs); free_dev(e->ozYTs);
    free_dev(e->ozWs); free_dev(e->ozSs); free_dev(e->oz_slabs); free_dev(e->oz_colmax);
    free_dev(e->d_tc_fwd); free_dev(e->d_tc_rev);
    free_dev(e->d_inc); free_dev(e->d_states); free_dev(e->d_entries); free_dev(e->d_chunk); free_dev(e->d_direct); free_dev(e->d_single); free_dev(e->d_state_prior);
    for (auto ev : e->chunk_ready) cudaEventDestroy(ev);
    for (auto p : e->up_slots) free_dev(p);
    for (auto ev : e->up_frees) cudaEventDestroy(ev);
    for (auto ev : e->timer.pool) cudaEventDestroy(ev);
    if (e->compute_done) cudaEventDestroy(e->compute_done);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    delete e;
}

extern "C" int pet_create(const pet_config *cfg, pet_engine **out) {
    if (!cfg || !out) { set_error("pet_create: null argument"); return PET_EINVAL; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: prosper_b200 has no CPU path");
        return PET_ECUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("device %d out of range", cfg->device); return PET_EINVAL; }
    PET_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    PET_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
        return PET_EINVAL;
    }
    if (cfg->model != PET_MODEL_BSC && cfg->model != PET_MODEL_TSC && cfg->model != PET_MODEL_DSC &&
        cfg->model != PET_MODEL_MCA && cfg->model != PET_MODEL_MMCA && cfg->model != PET_MODEL_GSC) {
        set_error("model kind %d is not handled by this engine entry point", cfg->model);
        return PET_EINVAL;
    }
    if (cfg->D < 1 || cfg->H < 1 || cfg->Hprime < 1 || cfg->gamma < 1 || cfg->Hprime > cfg->H ||
        cfg->gamma > cfg->Hprime) {   // camodels/__init__.py:90-91
        set_error("need 1 <= gamma <= Hprime <= H and D >= 1");
        return PET_EINVAL;
    }
    if (cfg->Hprime > PET_MAXHP || cfg->gamma > PET_MAXG) {
        set_error("Hprime <= %d and gamma <= %d supported", PET_MAXHP, PET_MAXG);
        return PET_EINVAL;
    }
    pet_engine *e = new pet_engine();
    e->model = cfg->model; e->device = cfg->device; e->sm_count = prop.multiProcessorCount;
    e->D = (int)cfg->D; e->H = (int)cfg->H; e->Hp = (int)cfg->Hprime; e->gamma = (int)cfg->gamma;
    e->ldY = round_up(e->D + 1, 8);
    e->ldH = round_up(e->H, 8);

    std::vector<std::vector<int>> rows;
    GLStatic &g = e->gls;
    memset(&g, 0, sizeof(g));
    const bool maxmodel = (e->model == PET_MODEL_MCA || e->model == PET_MODEL_MMCA);
    if (e->model == PET_MODEL_BSC || maxmodel || e->model == PET_MODEL_GSC) {
        e->values = {0.0, 1.0}; e->k0 = 0; e->K = 2; e->binary = true;
        enum_binary(e->Hp, e->gamma, rows);
        g.has_null = 1; g.n_blocks = 1; g.block_val[0] = 1.0; g.block_vidx[0] = 0;
        g.zbase = 0; g.diag_from_colsum = 1;
        g.select_mode = (e->model == PET_MODEL_BSC) ? SEL_BSC : (e->model == PET_MODEL_MCA ? SEL_GIVEN : SEL_NEGDIST);
    } else {
        std::vector<double> vals;
        if (e->model == PET_MODEL_TSC) vals = {-1.0, 0.0, 1.0};
        else {
            if (!cfg->states || cfg->n_states < 2) { delete e; set_error("DSC needs the latent values (states)"); return PET_EINVAL; }
            vals.assign(cfg->states, cfg->states + cfg->n_states);
        }
        int k0 = -1;
        for (size_t k = 0; k < vals.size(); ++k) if (vals[k] == 0.0) { if (k0 >= 0) k0 = -2; else k0 = (int)k; }
        if (k0 < 0) { delete e; set_error("states must contain exactly one 0"); return PET_EINVAL; }
        if ((int)vals.size() - 1 > PET_MAXV) { delete e; set_error("at most %d non-zero latent values", PET_MAXV); return PET_EINVAL; }
        e->values = vals; e->k0 = k0; e->K = (int)vals.size(); e->binary = false;
        if (e->model == PET_MODEL_TSC) {
            enum_product(e->Hp, e->K, k0, e->gamma, 0, rows);     // null and singletons included (tsc_et.py:76)
            g.has_null = 0; g.n_blocks = 0; g.zbase = e->Hp; g.select_mode = SEL_TSC;
        } else {
            enum_product(e->Hp, e->K, k0, e->gamma, 2, rows);     // dsc_et.py:59-61
            g.has_null = 1; g.zbase = e->H; g.select_mode = SEL_DSC;
            int b = 0;
            for (int k = 0; k < e->K; ++k) if (k != k0) { g.block_val[b] = vals[k]; g.block_vidx[b] = b; ++b; }
            g.n_blocks = b;
        }
        g.diag_from_colsum = 0;
    }
    int rc = build_state_space(e->ss, e->Hp, rows, e->values, e->k0, e->binary);
    if (rc != PET_OK) { delete e; return rc; }
    g.H = e->H; g.Hp = e->Hp; g.S = (int)e->ss.S; g.ldH = (int)e->ldH;
    g.n_cnt = e->K - 1;
    { int v = 0; for (int k = 0; k < e->K; ++k) if (k != e->k0) g.vals[v++] = e->values[k]; }
    e->C = g.has_null + (int64_t)g.n_blocks * e->H + e->ss.S;
    g.C = (int)e->C;
    g.binary = e->binary ? 1 : 0;
    for (int gsz = 0; gsz < PET_MAXG + 2; ++gsz) g.size_start[gsz] = (int)e->ss.S;
    if (e->binary) {   // states are enumerated by size 2..gamma
        int64_t idx = 0;
        for (int gsz = 2; gsz <= e->gamma; ++gsz) {
            g.size_start[gsz] = (int)idx;
            double c = 1.0;
            for (int t = 0; t < gsz; ++t) c = c * (e->Hp - t) / (t + 1);
            idx += (int64_t)(c + 0.5);
        }
    }
    g.n_chunks = e->ss.n_chunks; g.chunk_len = e->ss.chunk_len;
    g.n_direct = (int)e->ss.direct.size();
    g.n_out = e->ss.n_out;
    g.n_g = e->ss.n_g;
    if (gl_pick_warps(g) == 0) {
        delete e;
        set_error("H=%d with %lld states needs more shared memory than one SM has", e->H, (long long)e->ss.S);
        return PET_EINVAL;
    }

#define TRY(x) do { rc = (x); if (rc != PET_OK) { pet_destroy(e); return rc; } } while (0)
#define TRYC(x) do { cudaError_t _c = (x); if (_c != cudaSuccess) { set_error("%s: %s", #x, cudaGetErrorString(_c)); pet_destroy(e); return PET_ECUDA; } } while (0)
    TRY(dev_alloc(&e->d_states, std::max<int64_t>(1, e->ss.S)));
    TRY(dev_alloc(&e->d_entries, (int64_t)e->ss.entries.size()));
    TRY(dev_alloc(&e->d_chunk, (int64_t)e->ss.chunk_tab.size()));
    TRY(dev_alloc(&e->d_direct, (int64_t)e->ss.direct.size()));
    TRY(dev_alloc(&e->d_single, (int64_t)e->ss.single_idx.size()));
    TRY(dev_alloc(&e->d_state_prior, e->ss.S));
    if (e->ss.S) TRYC(cudaMemcpy(e->d_states, e->ss.records.data(), e->ss.S * 8, cudaMemcpyHostToDevice));
    TRYC(cudaMemcpy(e->d_entries, e->ss.entries.data(), e->ss.entries.size() * 2, cudaMemcpyHostToDevice));
    TRYC(cudaMemcpy(e->d_chunk, e->ss.chunk_tab.data(), e->ss.chunk_tab.size() * 2, cudaMemcpyHostToDevice));
    if (!e->ss.direct.empty()) TRYC(cudaMemcpy(e->d_direct, e->ss.direct.data(), e->ss.direct.size() * 4, cudaMemcpyHostToDevice));
    TRYC(cudaMemcpy(e->d_single, e->ss.single_idx.data(), e->ss.single_idx.size() * 4, cudaMemcpyHostToDevice));
    g.inc_states = nullptr;
    if (!e->ss.inc_records.empty() && !getenv("PET_GL_NO_INC")) {
        TRY(dev_alloc(&e->d_inc, e->ss.S));
        TRYC(cudaMemcpy(e->d_inc, e->ss.inc_records.data(), e->ss.S * 8, cudaMemcpyHostToDevice));
        g.inc_states = e->d_inc;
    }
    if (e->model == PET_MODEL_BSC && gl_tc_supported(g, e->gamma, e->binary)) {
        TRY(gl_tc_build_tables(g, e->gamma, e->ss.matrix, e->tc_host));
        TRY(dev_alloc(&e->d_tc_fwd, (int64_t)e->tc_host.bfwd.size()));
        TRY(dev_alloc(&e->d_tc_rev, (int64_t)e->tc_host.brev.size()));
        TRYC(cudaMemcpy(e->d_tc_fwd, e->tc_host.bfwd.data(), e->tc_host.bfwd.size(), cudaMemcpyHostToDevice));
        TRYC(cudaMemcpy(e->d_tc_rev, e->tc_host.brev.data(), e->tc_host.brev.size(), cudaMemcpyHostToDevice));
        e->tc_host.dev.bfwd = e->d_tc_fwd; e->tc_host.dev.brev = e->d_tc_rev;
        e->tc_ok = true;
        if (const char *env = getenv("PET_GL_TC")) e->tc_mode = atoi(env) == 0 ? 1 : 2;
    }
    g.states = e->d_states; g.entries = e->d_entries; g.chunk_tab = e->d_chunk; g.direct = e->d_direct; g.single_idx = e->d_single;

    // chunk-sized buffers are allocated when the shard is bound (size_chunks): the chunk length depends on its size
    e->chunk_cfg = cfg->chunk_rows;
    if (const char *env = getenv("PET_CHUNK_ROWS")) { if (atoll(env) > 0) e->chunk_cfg = atoll(env); }
    e->chunk_rows = 0;

    TRY(dev_alloc(&e->Wt, e->ldH * e->ldY));
    TRY(dev_alloc(&e->G, e->ldH * e->ldH));
    TRY(dev_alloc(&e->wn2, e->ldH)); TRY(dev_alloc(&e->invn, e->ldH));
    TRY(dev_alloc(&e->wmu, e->ldH)); TRY(dev_alloc(&e->mu_full, e->ldY));
    TRYC(cudaMemset(e->wmu, 0, e->ldH * 8));
    TRY(dev_alloc(&e->Wtmp, (int64_t)e->D * e->ldH));
    TRY(dev_alloc(&e->mu_dev, e->ldY));
    if (e->model == PET_MODEL_DSC) TRY(dev_alloc(&e->s2sum, e->ldH));
    if (e->model == PET_MODEL_GSC) {
        TRY(dev_alloc(&e->Wt2, e->ldH * e->ldY));
        TRY(dev_alloc(&e->gsc_tab, 8 * e->ldH));          // g, ilam, lcdet, logit, mu, pi, (spare)
        TRY(dev_alloc(&e->psi_dev, e->ldH * e->ldH));
        TRY(dev_alloc(&e->bdiag, e->ldY));
        TRYC(cudaMemset(e->Wt2, 0, e->ldH * e->ldY * 8));
        TRYC(cudaMemset(e->psi_dev, 0, e->ldH * e->ldH * 8));
    }
    if (maxmodel) {
        e->ldD = round_up(e->D, 2);
        TRY(dev_alloc(&e->Wl, (int64_t)e->H * e->ldD)); TRY(dev_alloc(&e->Wr, (int64_t)e->H * e->ldD));
    }
    if (e->model == PET_MODEL_BSC || e->model == PET_MODEL_TSC || e->model == PET_MODEL_DSC) {
        // score and statistics GEMMs on the int8 tensor cores (PET_OZAKI=0 keeps the FP64 DMMA kernels)
        const char *env = getenv("PET_OZAKI"), *envs = getenv("PET_OZAKI_SLICES");
        e->oz_want = !(env && atoi(env) == 0);
        // Slices per operand.  The score GEMM y.W feeds log-joints: 6 slices (42 bits below the row / column maxima, an
        // absolute error ~1e-10 in F and so ~1e-10 relative in every posterior) leave the parity unchanged and cost 21
        // slice products with a three-stage pipeline instead of 28 with two.  The statistics GEMM feeds the H x H solve,
        // which amplifies its rounding by the condition number of sum <s s^T>, and its <s> operand is fixed point
        // (scale 1): it keeps all 7 slices (49 bits).  PET_OZAKI_SLICES=6|7 sets both, _SCORE / _STATS one of them.
        e->oz_ns = 6;
        e->oz_ns2 = 7;
        if (envs) e->oz_ns = e->oz_ns2 = (atoi(envs) == 6) ? 6 : 7;
        if (const char *v = getenv("PET_OZAKI_SLICES_SCORE")) e->oz_ns = (atoi(v) == 6) ? 6 : 7;
        if (const char *v = getenv("PET_OZAKI_SLICES_STATS")) e->oz_ns2 = (atoi(v) == 6) ? 6 : 7;
        if (e->oz_want) {
            e->oz_kpd = ozaki_kp(e->D);
            TRY(dev_alloc(&e->ozW, (int64_t)e->oz_ns * e->H * e->oz_kpd));
            TRY(dev_alloc(&e->ozWs, e->ldH));
            TRY(dev_alloc(&e->ozSs, e->ldH));
            TRY(dev_alloc(&e->oz_colmax, std::max<int64_t>(e->ldH, e->ldY)));
        }
    }
    TRY(dev_alloc(&e->solveA, (int64_t)(e->H + e->D) * e->ldH));     // B stacked under A: the solve fuses its forward sweep
    e->solveB = e->solveA + (int64_t)e->H * e->ldH;
    TRY(dev_alloc(&e->solve_work, spd_solve_work_doubles(e->H, e->ldH)));
    TRY(dev_alloc(&e->ksel_state, 2 + 256));
    TRYC(cudaMemset(e->Wt, 0, e->ldH * e->ldY * 8));
    TRYC(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    TRYC(cudaEventCreateWithFlags(&e->compute_done, cudaEventDisableTiming));
#undef TRY
#undef TRYC
    e->launches0 = g_launches;
    *out = e;
    return PET_OK;
}

extern "C" int64_t pet_num_states(const pet_engine *e) { return e ? e->ss.S : -1; }
extern "C" int64_t pet_num_columns(const pet_engine *e) { return e ? e->C : -1; }
extern "C" int64_t pet_num_data(const pet_engine *e) { return e ? e->n : -1; }
extern "C" int64_t pet_launch_count(const pet_engine *e) { return e ? (g_launches - e->launches0) : -1; }
extern "C" int pet_state_matrix(const pet_engine *e, double *out_host) {
    if (!e || !out_host) { set_error("pet_state_matrix: null argument"); return PET_EINVAL; }
    memcpy(out_host, e->ss.matrix.data(), e->ss.matrix.size() * sizeof(double));
    return PET_OK;
}
extern "C" int32_t pet_gemm_path(const pet_engine *e) { return (e && e->oz_on) ? e->oz_ns2 : 0; }
extern "C" int32_t pet_gemm_slices(const pet_engine *e, int32_t *score, int32_t *stats) {
    if (!e || !score || !stats) return PET_EINVAL;
    *score = e->oz_on ? e->oz_ns : 0;
    *stats = e->oz_on ? e->oz_ns2 : 0;
    return PET_OK;
}
// the tensor-core state kernel pays off once the state space fills a few 64-state chunks
// (a 128-datapoint tile is one CTA's sequential work: ~250 us at 1573 states) and the chunk fills at least half a wave of
// tiles; below that the scalar kernel (8 datapoints per CTA pass) finishes sooner
static bool use_state_tc(const pet_engine *e, int kflags, int64_t rows) {
    if (!e->tc_ok || e->tc_mode == 1 || (kflags & (GLF_READ_LOGPJ | GLF_WRITE_LOGPJ))) return false;
    return e->tc_mode == 2 || (e->ss.S >= 256 && rows >= int64_t(64) * e->sm_count);
}
extern "C" int pet_set_state_kernel(pet_engine *e, int32_t mode) {
    if (!e || mode < 0 || mode > 2) { set_error("pet_set_state_kernel: mode must be 0 (auto), 1 (scalar) or 2 (tensor cores)"); return PET_EINVAL; }
    if (mode == 2 && !e->tc_ok) { set_error("pet_set_state_kernel: this model / state space has no tensor-core state kernel"); return PET_EINVAL; }
    e->tc_mode = mode;
    return PET_OK;
}
extern "C" int32_t pet_state_kernel_path(const pet_engine *e) {
    return (e && use_state_tc(e, 0, e->chunk_start.size() > 1 ? e->chunk_start[1] : (e->n > 0 ? e->n : (int64_t(1) << 40)))) ? 2 : 1;
}
extern "C" int pet_enable_timing(pet_engine *e, int32_t on) {
    if (!e) return PET_EINVAL;
    e->timer.reset();
    e->timer.on = on != 0;
    return PET_OK;
}
extern "C" int pet_stage_times_ms(pet_engine *e, double *out) {
    if (!e || !out) return PET_EINVAL;
    e->timer.collect();
    for (int i = 0; i < ST_COUNT; ++i) { out[i] = e->timer.totals[i]; out[ST_COUNT + i] = (double)e->timer.counts[i]; }
    return PET_OK;
}

// ---- data ------------------------------------------------------------------------------
// Chunk length for a shard of n datapoints and the buffers that scale with it.  Long chunks amortise the wave tails of
// every kernel of the pipeline (measured at the north-star shape: 83 ms per iteration with 128 MB chunks of <S>, 73 ms
// with 600 MB in round 1), so the default aims at ~600 MB of <S> per chunk (75 776 datapoints at H = 1000), bounded by
// the shard itself.  (Round 2: 1200 MB chunks looked 2 ms faster; the cause was the wave quantisation of the state
// kernel's 128-datapoint tiles at the chunk length the tuner had picked, see below -- with that fixed 600 and 1200 MB are
// equal, 45.0 / 44.8 ms, and the shorter chunk overlaps a host upload better.)  The length is then tuned
// so that the 128 x 64 tiles of the score GEMM fill whole waves of sm_count persistent CTAs.
static int size_chunks(pet_engine *e, int64_t n) {
    int64_t cr;
    if (e->chunk_cfg > 0) cr = round_up(std::max<int64_t>(128, std::min<int64_t>(e->chunk_cfg, 1 << 20)), 128);
    else {
        const int64_t target = std::max<int64_t>(2048, e->chunk_target_bytes / (e->ldH * 8));
        if (n <= target * 5 / 4) cr = round_up(std::max<int64_t>(n, 128), 128);          // one chunk
        else {
            cr = round_up(target, 128);
            const int64_t ntile = ceil_div(e->H, 64);
            double best_eff = 0.0;
            int64_t best = cr;
            for (int64_t c = round_up(target * 3 / 4, 128); c <= target * 5 / 4; c += 128) {
                const int64_t tiles = (c / 128) * ntile, waves = ceil_div(tiles, e->sm_count);
                // ... and the 128-datapoint tiles of the tensor-core state kernel (one CTA per SM) whole waves as well:
                // 80 512 rows = 629 tiles = 4.25 waves cost that kernel 10.2 ms per iteration, 75 776 = 4 waves 8.9 ms
                const int64_t stiles = c / 128, swaves = ceil_div(stiles, e->sm_count);
                const double eff = double(tiles) / double(waves * e->sm_count) *
                                   (e->tc_ok ? double(stiles) / double(swaves * e->sm_count) : 1.0);
                if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && llabs(c - target) < llabs(best - target))) { best_eff = eff; best = c; }
            }
            cr = best;
        }
    }
    if (cr == e->chunk_rows && e->Sbuf) return PET_OK;
    cudaDeviceSynchronize();
    free_dev(e->Sbuf); free_dev(e->S2buf); free_dev(e->XSZ); free_dev(e->SZ2); free_dev(e->simbuf); free_dev(e->gemm_work);
    free_dev(e->ozS); free_dev(e->oz_slabs); free_dev(e->gsc_T); free_dev(e->stage_logpj);
    e->Sbuf = e->S2buf = e->XSZ = e->SZ2 = e->simbuf = e->gemm_work = e->oz_slabs = e->gsc_T = e->stage_logpj = nullptr;
    e->ozS = nullptr;
    e->stage_logpj_doubles = 0;
    for (auto p : e->up_slots) free_dev(p);
    e->up_slots.clear();
    e->chunk_rows = cr;
    e->n_cap = 0;                                     // per-shard buffers are laid out by chunk: reallocate them too
    PET_CHECK(dev_alloc(&e->Sbuf, cr * e->ldH));
    if (e->model == PET_MODEL_DSC) PET_CHECK(dev_alloc(&e->S2buf, cr * e->ldH));
    if (e->model == PET_MODEL_GSC) { PET_CHECK(dev_alloc(&e->XSZ, cr * e->ldH)); PET_CHECK(dev_alloc(&e->SZ2, cr * e->ldH)); }
    if (e->model == PET_MODEL_MCA) PET_CHECK(dev_alloc(&e->simbuf, cr * e->ldH));
    {
        int splits = dgemm_mn_splits(e->D + 1, e->H, cr, e->sm_count);
        e->gemm_work_doubles = int64_t(splits) * (e->D + 1) * e->ldH;
        if (e->model == PET_MODEL_GSC)
            e->gemm_work_doubles = std::max<int64_t>(e->gemm_work_doubles, int64_t(dgemm_mn_splits(e->H, e->H, cr, e->sm_count)) * e->H * e->ldH);
        PET_CHECK(dev_alloc(&e->gemm_work, e->gemm_work_doubles));
    }
    if (e->oz_want) {
        // The statistics GEMM is computed TRANSPOSED, (H, D+1) = <S>^T . Y with <S> as the 128-row operand: 8 x 11 tiles of
        // 1024 x 704 instead of 6 x 16 tiles of 768 x 1024 for the 677 x 1000 result (8 % less padding) and a split count
        // that fills whole waves (5 x 88 = 440 units = 2.97 waves instead of 3 x 96 = 1.95); the slabs are summed into
        // the (D+1, H) layout of the packed statistics by ozaki_add_slabs_t
        e->oz_ldT = round_up((int64_t)e->D + 1, 8);
        e->oz_splits = ozaki_splits(e->H, e->D + 1, ozaki_kp(cr), e->sm_count, 512);      // <s> may arrive as unsigned digits
        PET_CHECK(dev_alloc(&e->ozS, (int64_t)e->oz_ns2 * e->H * cr));
        PET_CHECK(dev_alloc(&e->oz_slabs, (int64_t)e->oz_splits * e->H * e->oz_ldT));
    }
    return PET_OK;
}

static int ensure_rows(pet_engine *e, int64_t n) {
    if (n <= e->n_cap) return PET_OK;
    cudaDeviceSynchronize();
    free_dev(e->Y); free_dev(e->yy); free_dev(e->cand); free_dev(e->lse); free_dev(e->YW);
    free_dev(e->rs); free_dev(e->ywc); free_dev(e->scl);
    free_dev(e->pairs); e->pairs = nullptr; e->pairs_ld = 0; e->defer_valid = false;
    e->Y = nullptr; e->yy = nullptr; e->cand = nullptr; e->lse = nullptr; e->YW = nullptr;
    e->rs = nullptr; e->ywc = nullptr; e->scl = nullptr;
    e->n_cap = 0;
    PET_CHECK(dev_alloc(&e->Y, n * e->ldY));
    PET_CHECK(dev_alloc(&e->yy, n));
    PET_CHECK(dev_alloc(&e->cand, n * e->Hp));
    PET_CHECK(dev_alloc(&e->lse, n));
    PET_CHECK(dev_alloc(&e->rs, n * (4 + PET_MAXV)));
    PET_CHECK(dev_alloc(&e->ywc, n * e->Hp));
    PET_CHECK(dev_alloc(&e->scl, n * (1 + PET_MAXHP)));
    if (e->model == PET_MODEL_GSC) { free_dev(e->yyw); e->yyw = nullptr; PET_CHECK(dev_alloc(&e->yyw, n)); }
    // cache the whole score matrix when it is affordable (lets the truncated M-step skip the
    // second score GEMM); otherwise one chunk
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    int64_t all_bytes = n * e->ldH * 8;
    e->yw_all = all_bytes <= (int64_t)(free_b / 3);
    e->yw_rows = e->yw_all ? n : std::min<int64_t>(n, e->chunk_rows);
    PET_CHECK(dev_alloc(&e->YW, e->yw_rows * e->ldH));
    e->n_cap = n;
    free_dev(e->ozY); free_dev(e->ozYT); free_dev(e->ozYs); free_dev(e->ozYTs);
    e->ozY = e->ozYT = nullptr; e->ozYs = e->ozYTs = nullptr;
    e->oz_on = false;
    if (e->oz_want) {
        // row slices (score GEMM) and per-chunk transposed column slices (statistics GEMM) of the shard;
        // without room for them the FP64 kernels take over
        const int64_t nchunks = ceil_div(n, e->chunk_rows) + 4;      // + the short chunks of a ramped chunk table
        const int64_t b1 = (int64_t)e->oz_ns * n * e->oz_kpd, b2 = (int64_t)e->oz_ns2 * nchunks * (e->D + 1) * e->chunk_rows;
        cudaMemGetInfo(&free_b, &total_b);
        if (b1 + b2 + (int64_t(1) << 30) < (int64_t)free_b &&
            dev_alloc(&e->ozY, b1) == PET_OK && dev_alloc(&e->ozYT, b2) == PET_OK &&
            dev_alloc(&e->ozYs, n) == PET_OK && dev_alloc(&e->ozYTs, nchunks * e->ldY) == PET_OK) {
            e->oz_on = true;
            e->oz_rows = n;
        } else {
            free_dev(e->ozY); free_dev(e->ozYT); free_dev(e->ozYs); free_dev(e->ozYTs);
            e->ozY = e->ozYT = nullptr; e->ozYs = e->ozYTs = nullptr;
        }
    }
    return PET_OK;
}

// Uniform chunks for a resident shard.  A shard that is being uploaded is consumed at PCIe speed, so what counts is when
// the first chunk can start and how much work is left after the last byte arrived: short chunks at both ends, long ones
// in between.
static void build_chunk_table(pet_engine *e, int64_t n, bool ramp) {
    std::vector<int64_t> &t = e->chunk_start;
    t.assign(1, 0);
    const int64_t cr = e->chunk_rows;
    const int64_t s1 = round_up(std::max<int64_t>(cr / 4, 128), 128), s2 = std::min(cr, 2 * s1);
    if (ramp && e->chunk_cfg <= 0 && n >= 3 * cr && 2 * (s1 + s2) < n) {
        t.push_back(s1);
        t.push_back(s1 + s2);
        const int64_t mid = n - 2 * (s1 + s2), k = ceil_div(mid, cr), len = round_up(ceil_div(mid, k), 128);
        for (int64_t i = 0; i < k; ++i) t.push_back(std::min(t.back() + len, s1 + s2 + mid));
        t.push_back(n - s1);
        t.push_back(n);
    } else {
        for (int64_t r = cr; r < n; r += cr) t.push_back(r);
        t.push_back(n);
    }
}

extern "C" int pet_set_chunk_target(pet_engine *e, int64_t bytes_of_posterior_per_chunk) {
    if (!e) { set_error("pet_set_chunk_target: null engine"); return PET_EINVAL; }
    e->chunk_target_bytes = bytes_of_posterior_per_chunk > 0 ? bytes_of_posterior_per_chunk : (int64_t(600) << 20);
    return PET_OK;
}

extern "C" int pet_set_data(pet_engine *e, const double *y, int64_t n, int64_t ld, void *stream) {
    if (!e || !y || n < 0 || ld < e->D) { set_error("pet_set_data: bad arguments"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    PET_CHECK(size_chunks(e, n));
    PET_CHECK(ensure_rows(e, n));
    e->n = n;
    e->yy_valid = false;
    e->cand_state = 0;
    e->mu_applied.assign(e->D, 0.0);
    e->chunk_start.assign(1, 0);
    if (n == 0) return PET_OK;
    const bool on_device = is_device_ptr(y);
    build_chunk_table(e, n, !on_device);
    if (on_device) {
        PET_CUDA(cudaMemcpy2DAsync(e->Y, e->ldY * 8, y, ld * 8, size_t(e->D) * 8, n, cudaMemcpyDeviceToDevice, st));
        e->upload_pending = false;
    } else {
        // the copy stream must not overwrite Y / the staging slots while earlier kernels still read them
        if (e->compute_done_valid) PET_CUDA(cudaStreamWaitEvent(e->copy_stream, e->compute_done, 0));
        e->up_gran = std::min<int64_t>(e->chunk_rows, 16384);
        const int64_t ngran = ceil_div(n, e->up_gran);
        while ((int64_t)e->chunk_ready.size() < ngran) {
            cudaEvent_t ev;
            PET_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e->chunk_ready.push_back(ev);
        }
        e->up_src = y; e->up_ld = ld; e->up_copied = 0; e->up_expanded = 0;
        e->up_staged = (ld == e->D) && ngran > 1;
        if (e->up_staged && e->up_slots.empty()) {
            const int want = int(std::min<int64_t>(ngran, 2 * ceil_div(e->chunk_rows, e->up_gran) + 2));
            for (int i = 0; i < want; ++i) {
                double *p = nullptr;
                if (dev_alloc(&p, e->up_gran * e->D) != PET_OK) { cudaGetLastError(); break; }
                e->up_slots.push_back(p);
            }
            while (e->up_frees.size() < e->up_slots.size()) {
                cudaEvent_t ev;
                PET_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                e->up_frees.push_back(ev);
            }
            if (e->up_slots.size() < 3) {          // no room for staging: direct pitched copies
                for (auto p : e->up_slots) free_dev(p);
                e->up_slots.clear();
                e->up_staged = false;
            }
        }
        e->upload_pending = true;
    }
    return PET_OK;
}

// ---- per-iteration preparation ---------------------------------------------------------
static int load_W(pet_engine *e, const pet_params *p, cudaStream_t st) {
    if (!p || !p->W || p->ldW < e->H || !p->pi_host || p->n_pi < 1) { set_error("bad model parameters"); return PET_EINVAL; }
    const double *Wsrc = p->W;
    int64_t ldw = p->ldW;
    if (!is_device_ptr(p->W)) {
        PET_CUDA(cudaMemcpy2DAsync(e->Wtmp, e->ldH * 8, p->W, p->ldW * 8, size_t(e->H) * 8, e->D, cudaMemcpyHostToDevice, st));
        Wsrc = e->Wtmp; ldw = e->ldH;
    }
    e->Wsrc = Wsrc; e->Wsrc_ld = ldw;
    PET_CHECK(launch_transpose_w(e->Wt, e->ldY, Wsrc, ldw, e->D, e->H, st));
    return PET_OK;
}

static int flush_upload(pet_engine *e, cudaStream_t st);

static int apply_mu(pet_engine *e, const pet_params *p, cudaStream_t st) {
    // BSC only: y - mu (bsc_et.py:169,335,396).  The shard is stored shifted by the mu in force.
    std::vector<double> mu(e->D, 0.0);
    if (p->mu) {
        if (is_device_ptr(p->mu)) PET_CUDA(cudaMemcpyAsync(mu.data(), p->mu, e->D * 8, cudaMemcpyDeviceToHost, st));
        else memcpy(mu.data(), p->mu, e->D * 8);
        PET_CUDA(cudaStreamSynchronize(st));
    }
    if (e->mu_applied.size() != (size_t)e->D) e->mu_applied.assign(e->D, 0.0);
    bool same = true;
    std::vector<double> delta(e->D);
    for (int d = 0; d < e->D; ++d) { delta[d] = mu[d] - e->mu_applied[d]; same &= (delta[d] == 0.0); }
    if (same) return PET_OK;
    PET_CHECK(flush_upload(e, st));   // all chunks must have landed before shifting in place
    PET_CUDA(cudaMemcpyAsync(e->mu_dev, delta.data(), e->D * 8, cudaMemcpyHostToDevice, st));
    PET_CUDA(cudaStreamSynchronize(st));
    PET_CHECK(launch_subtract_mu(e->Y, e->ldY, e->n, e->D, e->mu_dev, st));
    e->mu_applied = mu;
    e->yy_valid = false;
    return PET_OK;
}

static int prepare(pet_engine *e, const pet_params *p, cudaStream_t st) {
    e->timer.begin(ST_PREPARE, st);
    PET_CHECK(load_W(e, p, st));
    if (e->model == PET_MODEL_BSC) PET_CHECK(apply_mu(e, p, st));
    // G = W^T W  (H,H); its diagonal gives ||W_h||^2 (bsc_et.py:111 recomputes this per datapoint)
    PET_CHECK(dgemm_kk(e->H, e->H, e->D, e->Wt, e->ldY, e->Wt, e->ldY, e->G, e->ldH, 1.0, 0, st));
    PET_CHECK(launch_gram_diag(e->G, e->ldH, e->H, e->wn2, e->invn, st));
    if (e->model == PET_MODEL_BSC) {          // W_h . mu for the selection scores of the un-shifted datapoints
        bool nonzero = false;
        for (double v : e->mu_applied) nonzero |= (v != 0.0);
        e->wmu_nonzero = nonzero;
        if (nonzero) {
            PET_CUDA(cudaMemcpyAsync(e->mu_full, e->mu_applied.data(), e->D * 8, cudaMemcpyHostToDevice, st));
            PET_CUDA(cudaStreamSynchronize(st));      // mu_applied is host memory that may change before the copy runs
            PET_CHECK(launch_wdotmu(e->Wt, e->ldY, e->H, e->D, e->mu_full, e->wmu, st));
        } else {
            PET_CUDA(cudaMemsetAsync(e->wmu, 0, e->ldH * 8, st));
        }
    }
    if (e->oz_on)
        PET_CHECK(ozaki_slice_rows(e->Wt, e->ldY, e->H, e->D, e->oz_ns, e->ozW, (int64_t)e->H * e->oz_kpd, e->ozWs, st));
    e->timer.end(st);
    return PET_OK;
}

// model-specific log-prior constants
static int fill_iter(const pet_engine *e, const pet_anneal *a, const pet_params *p, GLIter &it) {
    memset(&it, 0, sizeof(it));
    if (!a || !(a->T > 0.0)) { set_error("annealing temperature T must be > 0"); return PET_EINVAL; }
    if (!(p->sigma > 0.0)) { set_error("sigma must be > 0"); return PET_EINVAL; }
    it.beta = 1.0 / a->T;                          // bsc_et.py:152
    it.pre1 = -1.0 / 2.0 / p->sigma / p->sigma;    // bsc_et.py:153
    it.anneal_prior = a->anneal_prior ? 1 : 0;
    const GLStatic &g = e->gls;
    if (e->model == PET_MODEL_BSC) {
        double pi = p->pi_host[0];
        double pil_bar = log(pi / (1.0 - pi));     // bsc_et.py:154
        it.prior_null = 0.0; it.prior_block[0] = pil_bar; it.lp[0] = pil_bar; it.lp0 = 0.0;
    } else if (e->model == PET_MODEL_TSC) {
        double pi = p->pi_host[0];                 // tsc_et.py:316-322
        it.lp[0] = log(pi / 2.0); it.lp[1] = log(pi / 2.0); it.lp0 = log(1.0 - pi);
    } else {
        if (p->n_pi != e->K) { set_error("DSC: pi must have K=%d entries", e->K); return PET_EINVAL; }
        double l0 = log(p->pi_host[e->k0]);
        it.lp0 = l0;
        it.prior_null = e->H * l0;                 // dsc_et.py:550
        int b = 0;
        for (int k = 0; k < e->K; ++k) if (k != e->k0) {
            it.lp[b] = log(p->pi_host[k]);
            it.prior_block[b] = it.lp[b] + (e->H - 1) * l0;   // dsc_et.py:555
            it.sel_prior[b] = it.prior_block[b];               // dsc_et.py:393
            ++b;
        }
    }
    (void)g;
    return PET_OK;
}

// The upload moves in GRANULES of up_gran rows (independent of the compute chunks, which are several times longer):
// 1-D copies into a ring of staging slots on the copy stream, expansion (padding, ones column, ||y||^2) on the CONSUMER's
// stream just before the chunk that needs the rows -- a kernel on the copy stream would queue behind the persistent compute
// kernels for SMs and stall the copy engine.  The ring holds two chunks' worth of granules, so the copies of chunk c + 1
// run while chunk c computes.  Granules are enqueued lazily, one compute chunk ahead of the consumer.
static int upload_enqueue_upto(pet_engine *e, int64_t row_end) {
    const int64_t ngran = ceil_div(e->n, e->up_gran);
    const int64_t g_end = std::min<int64_t>(ngran, ceil_div(row_end, e->up_gran));
    const int ns = (int)e->up_slots.size();
    while (e->up_copied < g_end) {
        const int64_t j = e->up_copied, r0 = j * e->up_gran, rows = std::min(e->up_gran, e->n - r0);
        if (e->up_staged) {
            if (j >= ns) PET_CUDA(cudaStreamWaitEvent(e->copy_stream, e->up_frees[j % ns], 0));
            PET_CUDA(cudaMemcpyAsync(e->up_slots[j % ns], e->up_src + r0 * e->up_ld, size_t(rows) * e->D * 8, cudaMemcpyHostToDevice,
                                     e->copy_stream));
        } else {
            PET_CUDA(cudaMemcpy2DAsync(e->Y + r0 * e->ldY, e->ldY * 8, e->up_src + r0 * e->up_ld, e->up_ld * 8, size_t(e->D) * 8,
                                       rows, cudaMemcpyHostToDevice, e->copy_stream));
        }
        PET_CUDA(cudaEventRecord(e->chunk_ready[j], e->copy_stream));
        e->up_copied++;
    }
    return PET_OK;
}

// wait for and expand the granules covering rows < row_end on stream st
static int upload_consume_upto(pet_engine *e, int64_t row_end, cudaStream_t st) {
    const int64_t g_end = std::min<int64_t>(ceil_div(e->n, e->up_gran), ceil_div(row_end, e->up_gran));
    const int ns = (int)e->up_slots.size();
    while (e->up_expanded < g_end) {
        const int64_t k = e->up_expanded, r0 = k * e->up_gran, rows = std::min(e->up_gran, e->n - r0);
        PET_CHECK(upload_enqueue_upto(e, (k + 1) * e->up_gran));
        PET_CUDA(cudaStreamWaitEvent(st, e->chunk_ready[k], 0));
        if (e->up_staged) {
            PET_CHECK(launch_rownorm_pad(e->up_slots[k % ns], e->D, e->Y + r0 * e->ldY, e->ldY, rows, e->D, e->yy + r0, st));
            PET_CUDA(cudaEventRecord(e->up_frees[k % ns], st));
        }
        e->up_expanded++;
    }
    if (e->up_expanded * e->up_gran >= e->n) e->upload_pending = false;
    return PET_OK;
}

static int flush_upload(pet_engine *e, cudaStream_t st) {
    if (!e->upload_pending) return PET_OK;
    return upload_consume_upto(e, e->n, st);
}

static int ensure_chunk_inputs(pet_engine *e, int64_t c, int64_t r0, int64_t rows, cudaStream_t st) {
    bool fresh = false;      // yy and the padding columns of rows that arrive through the staged expansion are already written
    if (e->upload_pending) {
        fresh = e->up_staged && e->up_expanded * e->up_gran <= r0;
        PET_CHECK(upload_consume_upto(e, r0 + rows, st));
        if (e->upload_pending) PET_CHECK(upload_enqueue_upto(e, std::min<int64_t>(e->n, r0 + rows + e->chunk_rows)));
    }
    if (!e->yy_valid) {
        if (!fresh) PET_CHECK(launch_rownorm_pad(e->Y + r0 * e->ldY, e->ldY, e->Y + r0 * e->ldY, e->ldY, rows, e->D, e->yy + r0, st));
        if (e->oz_on) {
            PET_CHECK(ozaki_slice_rows(e->Y + r0 * e->ldY, e->ldY, rows, e->D, e->oz_ns, e->ozY + r0 * e->oz_kpd,
                                       e->oz_rows * e->oz_kpd, e->ozYs + r0, st));
            const int64_t plane = (int64_t)(e->D + 1) * e->chunk_rows;
            PET_CHECK(ozaki_slice_cols(e->Y + r0 * e->ldY, e->ldY, rows, e->D + 1, e->oz_ns2, e->oz_colmax, false,
                                       e->ozYT + c * e->oz_ns2 * plane, e->chunk_rows, plane, e->ozYTs + c * e->ldY, st));
        }
    }
    return PET_OK;
}

static void mark_compute_done(pet_engine *e, cudaStream_t st) {
    cudaEventRecord(e->compute_done, st);
    e->compute_done_valid = true;
}

enum { PASS_SELECT = 1, PASS_REUSE_SCORES = 2, PASS_DEFER_STATS = 4 };
static inline bool user_logpj_flags(int kflags) { return (kflags & (GLF_READ_LOGPJ | GLF_WRITE_LOGPJ)) != 0; }

// One sweep over the shard.  kflags: GLF_* for the posterior kernel.
static int sweep_gl(pet_engine *e, const pet_anneal *a, const pet_params *p, int kflags, int pass_flags,
                 const double *logpj_user, int64_t ld_logpj, bool logpj_is_output, const double *cut_dev,
                 double *stats_dev, cudaStream_t st) {
    if (e->n <= 0) { set_error("no data bound (pet_set_data)"); return PET_ESTATE; }
    if (!(kflags & GLF_SELECT) && e->cand_state == 0) { set_error("no candidates: run select_Hprimes first"); return PET_ESTATE; }
    GLArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.st = e->gls;
    PET_CHECK(fill_iter(e, a, p, ga.it));
    const bool reuse = (pass_flags & PASS_REUSE_SCORES) && e->yw_all;
    if (!reuse) PET_CHECK(prepare(e, p, st));
    ga.flags = kflags;
    if (e->model == PET_MODEL_DSC) ga.flags |= (kflags & GLF_USE_CUT) ? GLF_CUT_STRICT : 0;
    ga.yy = e->yy; ga.wn2 = e->wn2; ga.invn = e->invn; ga.wmu = e->wmu_nonzero ? e->wmu : nullptr; ga.G = e->G;
    ga.state_prior = e->d_state_prior;
    PET_CHECK(launch_state_prior(ga.st, ga.it, e->d_state_prior, st));
    ga.cand = e->cand; ga.lse = e->lse; ga.cut = cut_dev;
    ga.rs = e->rs; ga.ywc = e->ywc; ga.scl = e->scl;
    pet_stats_layout lay;
    pet_stats_layout_get(e, &lay);
    const bool do_stats = !(kflags & (GLF_LSE_ONLY | GLF_SELECT_ONLY));
    if (do_stats) {
        if (!stats_dev) { set_error("stats buffer is null"); return PET_EINVAL; }
        PET_CUDA(cudaMemsetAsync(stats_dev, 0, lay.total * 8, st));
        ga.Wq = stats_dev + lay.off_Wq;
        ga.scalars = stats_dev + lay.off_scalars;
        if (e->S2buf) PET_CUDA(cudaMemsetAsync(e->s2sum, 0, e->ldH * 8, st));
        if (e->oz_on) PET_CUDA(cudaMemsetAsync(e->oz_slabs, 0, (int64_t)e->oz_splits * e->H * e->oz_ldT * 8, st));
    }
    const bool user_logpj = (kflags & (GLF_READ_LOGPJ | GLF_WRITE_LOGPJ)) != 0;
    const bool logpj_on_dev = user_logpj && is_device_ptr(logpj_user);
    if (user_logpj && !logpj_on_dev) {
        int64_t need = e->chunk_rows * e->C;
        if (need > e->stage_logpj_doubles) {
            cudaStreamSynchronize(st);
            free_dev(e->stage_logpj); e->stage_logpj = nullptr; e->stage_logpj_doubles = 0;
            PET_CHECK(dev_alloc(&e->stage_logpj, need));
            e->stage_logpj_doubles = need;
        }
    }
    // the int8 statistics path slices <s> anyway: the normalisation is folded into that load, no scale kernel
    static const bool no_fold = getenv("PET_GL_NO_FOLD") != nullptr;
    bool fold_scale = do_stats && e->oz_on && !e->S2buf && !no_fold;
    // BSC with the register-resident row kernels: the <s> chunk is never written -- the slicer recomputes the singleton
    // posteriors from the score rows (one exp per entry) and adds the candidate marginals the state kernel left in scl
    static const bool no_defer = getenv("PET_GL_NO_DEFER") != nullptr || getenv("PET_GL_NO_FAST_ROW") != nullptr;
    const bool defer_s = fold_scale && !no_defer && !user_logpj_flags(kflags) && e->model == PET_MODEL_BSC && e->H <= 1024;
    if (defer_s) { ga.flags |= GLF_NO_SROW; fold_scale = false; }
    if (fold_scale) ga.flags |= GLF_FOLD_SCALE;
    const int64_t nchunks = (int64_t)e->chunk_start.size() - 1;
    // Truncated iteration with one posterior evaluation (bsc_et.py:250-257 needs the log-denominators of ALL datapoints
    // before it knows which ones enter the statistics): the log-denominator sweep already runs the statistics form of the
    // tensor-core state kernel and parks what it found per datapoint; the statistics sweep that follows with the cut only
    // adds up the parked records of the datapoints that stay (gl_finalize_cut) and runs the slicer and the GEMM.
    const bool no_single = getenv("PET_GL_NO_SINGLE_EVAL") != nullptr;     // (read per sweep: the tests switch it)
    const bool was_valid = e->defer_valid;
    e->defer_valid = false;
    const bool defer_ok = !no_single && !no_defer && e->model == PET_MODEL_BSC && e->oz_on && !e->S2buf && e->H <= 1024 &&
                          e->yw_all && !user_logpj_flags(kflags) && e->tc_ok;
    bool defer_eval = defer_ok && (kflags & GLF_LSE_ONLY) && (pass_flags & PASS_DEFER_STATS);
    const bool defer_use = defer_ok && defer_s && was_valid && reuse && (kflags & GLF_USE_CUT) &&
                           (int64_t)e->chunk_deferred.size() == nchunks;
    // (parking and adding up within the chunk of an un-truncated sweep as well was measured: 10.9 against 10.5 ms, not kept)
    if (defer_eval) {
        const int64_t npairs = (int64_t)e->Hp * (e->Hp - 1) / 2;
        if (!e->pairs || e->pairs_ld < e->n) {
            free_dev(e->pairs); e->pairs = nullptr; e->pairs_ld = 0;
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            if ((size_t)(npairs * e->n * 8) <= free_b / 2 && dev_alloc(&e->pairs, npairs * e->n) == PET_OK) e->pairs_ld = e->n;
            else defer_eval = false;                          // no room: the two-sweep form
        }
        e->chunk_deferred.assign(nchunks, 0);
    }
    ga.pairs = e->pairs; ga.pairs_ld = e->pairs_ld;
    if (!e->tile_counter) PET_CUDA(cudaMalloc(&e->tile_counter, 64));
    ga.tile_counter = e->tile_counter;
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t r0 = e->chunk_start[c], rows = e->chunk_start[c + 1] - r0;
        PET_CHECK(ensure_chunk_inputs(e, c, r0, rows, st));
        double *yw = e->yw_all ? e->YW + r0 * e->ldH : e->YW;
        if (!reuse) {
            e->timer.begin(ST_SCORE, st);
            if (e->oz_on) {
                const OzOperand oy{e->ozY + r0 * e->oz_kpd, e->oz_kpd, e->oz_rows * e->oz_kpd, e->ozYs + r0};
                const OzOperand ow{e->ozW, e->oz_kpd, (int64_t)e->H * e->oz_kpd, e->ozWs};
                PET_CHECK(ozaki_gemm(rows, e->H, e->oz_kpd, e->oz_ns, oy, ow, yw, e->ldH, 1, 0, false, e->sm_count, st));
            } else {
                PET_CHECK(dgemm_kk(rows, e->H, e->D, e->Y + r0 * e->ldY, e->ldY, e->Wt, e->ldY, yw, e->ldH, 1.0, 0, st));
            }
            e->timer.end(st);
        }
        ga.n_rows = rows; ga.row0 = r0; ga.YW = yw; ga.S = e->Sbuf; ga.S2 = e->S2buf;
        if (user_logpj) {
            if (logpj_on_dev) { ga.logpj = const_cast<double *>(logpj_user); ga.ld_logpj = ld_logpj; }
            else {
                // stage this chunk; the kernel indexes logpj by GLOBAL row, so bias the pointer
                ga.logpj = e->stage_logpj - r0 * e->C; ga.ld_logpj = e->C;
                if (!logpj_is_output)
                    PET_CUDA(cudaMemcpy2DAsync(e->stage_logpj, e->C * 8, logpj_user + r0 * ld_logpj, ld_logpj * 8,
                                               size_t(e->C) * 8, rows, cudaMemcpyHostToDevice, st));
            }
        }
        if (defer_use && e->chunk_deferred[c]) {
            e->timer.begin(ST_POST, st);
            PET_CHECK(launch_gl_finalize_cut(ga, st));
            e->timer.end(st);
        } else {
            e->timer.begin(ST_ROW, st);
            PET_CHECK(launch_gl_row(ga, e->sm_count, st));
            e->timer.end(st);
            e->timer.begin(ST_POST, st);
            if (defer_eval && use_state_tc(e, kflags, rows)) {
                GLArgs g2 = ga;
                g2.flags = (ga.flags & ~GLF_LSE_ONLY) | GLF_NO_SROW | GLF_DEFER_STATS;
                PET_CHECK(launch_gl_state_tc(g2, e->tc_host.dev, e->sm_count, st));
                e->chunk_deferred[c] = 1;
            } else if (use_state_tc(e, kflags, rows)) PET_CHECK(launch_gl_state_tc(ga, e->tc_host.dev, e->sm_count, st));
            else PET_CHECK(launch_gl_state(ga, e->gamma, e->binary, e->sm_count, st));
            e->timer.end(st);
        }
        if (!fold_scale && !defer_s) {
            e->timer.begin(ST_SCALE, st);
            PET_CHECK(launch_gl_scale(ga, st));
            e->timer.end(st);
        }
        if (user_logpj && !logpj_on_dev && logpj_is_output)
            PET_CUDA(cudaMemcpy2DAsync(const_cast<double *>(logpj_user) + r0 * ld_logpj, ld_logpj * 8, e->stage_logpj,
                                       e->C * 8, size_t(e->C) * 8, rows, cudaMemcpyDeviceToHost, st));
        if (do_stats) {
            e->timer.begin(ST_STATS, st);
            // Wp^T (D+1, H) += Y_chunk^T . <S>_chunk ; row D (all-ones column of Y) = sum_n <s>
            if (e->oz_on) {
                const int64_t plane = (int64_t)(e->D + 1) * e->chunk_rows;
                e->timer.end(st);
                e->timer.begin(ST_SLICE, st);
                if (defer_s)
                    PET_CHECK(launch_gl_post_slice(ga, e->oz_ns2, ozaki_kp(rows), e->ozS, e->chunk_rows, (int64_t)e->H * e->chunk_rows,
                                                   e->ozSs, st));
                else
                    PET_CHECK(ozaki_slice_cols(e->Sbuf, e->ldH, rows, e->H, e->oz_ns2, e->oz_colmax, false, e->ozS,
                                               e->chunk_rows, (int64_t)e->H * e->chunk_rows, e->ozSs, st,
                                               fold_scale ? e->scl + r0 * (1 + PET_MAXHP) : nullptr, 1 + PET_MAXHP));
                e->timer.end(st);
                e->timer.begin(ST_STATS, st);
                const OzOperand oy{e->ozYT + c * e->oz_ns2 * plane, e->chunk_rows, plane, e->ozYTs + c * e->ldY};
                const OzOperand os{e->ozS, e->chunk_rows, (int64_t)e->H * e->chunk_rows, e->ozSs};
                PET_CHECK(ozaki_gemm(e->H, e->D + 1, ozaki_kp(rows), e->oz_ns2, os, oy, e->oz_slabs, e->oz_ldT, e->oz_splits,
                                     (int64_t)e->H * e->oz_ldT, true, e->sm_count, st, defer_s ? 64 * 127 : 4096));
            } else {
                PET_CHECK(dgemm_mn(e->D + 1, e->H, rows, e->Y + r0 * e->ldY, e->ldY, e->Sbuf, e->ldH,
                                   stats_dev + lay.off_Wp, e->ldH, 1, e->gemm_work, e->gemm_work_doubles, e->sm_count, st));
            }
            if (e->S2buf) PET_CHECK(launch_colsum(e->s2sum, e->S2buf, e->ldH, rows, e->H, st));
            e->timer.end(st);
        }
    }
    e->yy_valid = true;
    e->defer_valid = defer_eval;
    if (kflags & GLF_SELECT) e->cand_state = 1;
    if (do_stats && e->oz_on)   // Wp^T = sum of the split-K slabs
        PET_CHECK(ozaki_add_slabs_t(stats_dev + lay.off_Wp, e->ldH, e->oz_slabs, e->H, e->D + 1, e->oz_ldT,
                                    (int64_t)e->H * e->oz_ldT, e->oz_splits, st));
    if (do_stats && e->S2buf)   // DSC: singleton second moments onto the diagonal (dsc_et.py:701)
        PET_CHECK(launch_add_diag(stats_dev + lay.off_Wq, e->ldH, e->s2sum, e->H, st));
    mark_compute_done(e, st);
    return PET_OK;
}


// ---- MCA / MMCA -------------------------------------------------------------------------------
static double mca_rho(const pet_engine *e, double T) {
    if (e->model == PET_MODEL_MCA) return 1.0 / (1.0 - 1.0 / std::max(T, 1.05));          // mca_et.py:143-145
    double rho = 1.0 / (1.0 - 1.0 / std::max(T, 1.20));                                     // mmca_et.py:163-165
    return std::max(std::min(rho, 35.0), 1.0);
}

static int sweep_mca(pet_engine *e, const pet_anneal *a, const pet_params *p, int kflags, int pass_flags,
                     const double *logpj_user, int64_t ld_logpj, bool logpj_is_output, const double *cut_dev,
                     double *stats_dev, cudaStream_t st) {
    if (e->n <= 0) { set_error("no data bound (pet_set_data)"); return PET_ESTATE; }
    if (!(kflags & GLF_SELECT) && e->cand_state == 0) { set_error("no candidates: run select_Hprimes first"); return PET_ESTATE; }
    if (!a || !(a->T > 0.0)) { set_error("annealing temperature T must be > 0"); return PET_EINVAL; }
    if (!p || !(p->sigma > 0.0) || !p->pi_host) { set_error("bad model parameters"); return PET_EINVAL; }
    const int mmca = (e->model == PET_MODEL_MMCA) ? 1 : 0;
    const double rho = mca_rho(e, a->T);
    const bool reuse = (pass_flags & PASS_REUSE_SCORES) && e->yw_all;
    e->timer.begin(ST_PREPARE, st);
    if (!reuse) PET_CHECK(load_W(e, p, st));
    PET_CHECK(launch_mca_tables(e->Wt, e->ldY, e->H, e->D, rho, mmca, e->Wl, e->Wr, e->ldD, e->wn2, st));
    e->timer.end(st);

    GLArgs sel;                      // preselection reuses the top-H' code of the Gaussian-linear kernel
    memset(&sel, 0, sizeof(sel));
    sel.st = e->gls;
    sel.it.beta = 1.0; sel.it.pre1 = -1.0;
    sel.flags = GLF_SELECT | GLF_SELECT_ONLY;
    sel.yy = e->yy; sel.wn2 = e->wn2; sel.invn = e->invn; sel.wmu = e->wmu_nonzero ? e->wmu : nullptr; sel.G = e->G; sel.cand = e->cand; sel.lse = e->lse;
    sel.state_prior = e->d_state_prior;
    sel.rs = e->rs; sel.ywc = e->ywc; sel.scl = e->scl;

    MCAArgs m;
    memset(&m, 0, sizeof(m));
    m.mmca = mmca; m.D = e->D; m.H = e->H; m.Hp = e->Hp; m.S = (int)e->ss.S; m.C = (int)e->C; m.gamma = e->gamma;
    m.ldH = (int)e->ldH; m.ldY = (int)e->ldY; m.ldD = (int)e->ldD; m.ldc = e->D | 1;    // odd stride: conflict-free rows
    m.states = e->d_states;
    m.rho = rho; m.beta = 1.0 / a->T; m.pre1 = -1.0 / 2.0 / p->sigma / p->sigma;
    m.pil_bar = log(p->pi_host[0] / (1.0 - p->pi_host[0]));
    m.flags = kflags & (GLF_WRITE_LOGPJ | GLF_READ_LOGPJ | GLF_LSE_ONLY | GLF_USE_CUT);
    m.yy = e->yy; m.wn2 = e->wn2; m.Wl = e->Wl; m.Wr = e->Wr; m.cand = e->cand; m.lse = e->lse; m.cut = cut_dev;
    pet_stats_layout lay;
    pet_stats_layout_get(e, &lay);
    const bool select_only = (kflags & GLF_SELECT_ONLY) != 0;
    const bool do_stats = !(kflags & (GLF_LSE_ONLY | GLF_SELECT_ONLY));
    if (do_stats) {
        if (!stats_dev) { set_error("stats buffer is null"); return PET_EINVAL; }
        PET_CUDA(cudaMemsetAsync(stats_dev, 0, lay.total * 8, st));
        m.Wpm = stats_dev + lay.off_Wq;
        m.Wqm = m.Wpm + (int64_t)e->H * e->ldD;
        m.scalars = stats_dev + lay.off_scalars;
    }
    const bool user_logpj = (kflags & (GLF_READ_LOGPJ | GLF_WRITE_LOGPJ)) != 0;
    const bool logpj_on_dev = user_logpj && is_device_ptr(logpj_user);
    if (user_logpj && !logpj_on_dev) {
        int64_t need = e->chunk_rows * e->C;
        if (need > e->stage_logpj_doubles) {
            cudaStreamSynchronize(st);
            free_dev(e->stage_logpj); e->stage_logpj = nullptr; e->stage_logpj_doubles = 0;
            PET_CHECK(dev_alloc(&e->stage_logpj, need));
            e->stage_logpj_doubles = need;
        }
    }
    const int64_t nchunks = (int64_t)e->chunk_start.size() - 1;
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t r0 = e->chunk_start[c], rows = e->chunk_start[c + 1] - r0;
        PET_CHECK(ensure_chunk_inputs(e, c, r0, rows, st));
        double *yw = e->yw_all ? e->YW + r0 * e->ldH : e->YW;
        if (!reuse) {
            e->timer.begin(ST_SCORE, st);
            PET_CHECK(dgemm_kk(rows, e->H, e->D, e->Y + r0 * e->ldY, e->ldY, e->Wt, e->ldY, yw, e->ldH, 1.0, 0, st));
            e->timer.end(st);
        }
        if (kflags & GLF_SELECT) {
            e->timer.begin(ST_POST, st);
            sel.n_rows = rows; sel.row0 = r0;
            if (mmca) sel.YW = yw;                       // smallest ||W_h - y||^2  (mmca_et.py:118)
            else {                                       // smallest sum_d max(W_hd - y_d, 0)  (mca_et.py:105-106)
                PET_CHECK(launch_mca_sim(e->Y + r0 * e->ldY, e->ldY, rows, e->Wsrc, e->Wsrc_ld, e->D, e->H, e->simbuf, e->ldH, st));
                sel.YW = e->simbuf;
            }
            PET_CHECK(launch_gl_kernel(sel, e->gamma, true, e->sm_count, st));
            e->timer.end(st);
        }
        if (select_only) continue;
        m.n_rows = rows; m.row0 = r0; m.Y = e->Y + r0 * e->ldY; m.YW = yw; m.Spost = e->Sbuf;
        if (user_logpj) {
            if (logpj_on_dev) { m.logpj = const_cast<double *>(logpj_user); m.ld_logpj = ld_logpj; }
            else {
                m.logpj = e->stage_logpj - r0 * e->C; m.ld_logpj = e->C;
                if (!logpj_is_output)
                    PET_CUDA(cudaMemcpy2DAsync(e->stage_logpj, e->C * 8, logpj_user + r0 * ld_logpj, ld_logpj * 8,
                                               size_t(e->C) * 8, rows, cudaMemcpyHostToDevice, st));
            }
        }
        e->timer.begin(ST_POST, st);
        PET_CHECK(launch_mca_kernel(m, e->sm_count, st));
        e->timer.end(st);
        if (user_logpj && !logpj_on_dev && logpj_is_output)
            PET_CUDA(cudaMemcpy2DAsync(const_cast<double *>(logpj_user) + r0 * ld_logpj, ld_logpj * 8, e->stage_logpj,
                                       e->C * 8, size_t(e->C) * 8, rows, cudaMemcpyDeviceToHost, st));
        if (do_stats) {
            e->timer.begin(ST_STATS, st);
            PET_CHECK(dgemm_mn(e->D + 1, e->H, rows, e->Y + r0 * e->ldY, e->ldY, e->Sbuf, e->ldH,
                               stats_dev + lay.off_Wp, e->ldH, 1, e->gemm_work, e->gemm_work_doubles, e->sm_count, st));
            e->timer.end(st);
        }
    }
    e->yy_valid = true;
    if (kflags & GLF_SELECT) e->cand_state = 1;
    mark_compute_done(e, st);
    return PET_OK;
}


static int sweep(pet_engine *e, const pet_anneal *a, const pet_params *p, int kflags, int pass_flags,
                 const double *logpj_user, int64_t ld_logpj, bool logpj_is_output, const double *cut_dev,
                 double *stats_dev, cudaStream_t st) {
    if (e->model == PET_MODEL_MCA || e->model == PET_MODEL_MMCA)
        return sweep_mca(e, a, p, kflags, pass_flags, logpj_user, ld_logpj, logpj_is_output, cut_dev, stats_dev, st);
    return sweep_gl(e, a, p, kflags, pass_flags, logpj_user, ld_logpj, logpj_is_output, cut_dev, stats_dev, st);
}

extern "C" int pet_stats_layout_get(const pet_engine *e, pet_stats_layout *out) {
    if (!e || !out) { set_error("pet_stats_layout_get: null argument"); return PET_EINVAL; }
    memset(out, 0, sizeof(*out));
    out->off_Wp = 0; out->rows_Wp = e->D + 1; out->cols_Wp = e->H; out->ld_Wp = e->ldH;
    out->off_Wq = out->rows_Wp * out->ld_Wp; out->rows_Wq = e->H; out->cols_Wq = e->H; out->ld_Wq = e->ldH;
    if (e->model == PET_MODEL_MCA || e->model == PET_MODEL_MMCA) {   // two (H,D) blocks: multi-cause Wp then Wq
        out->rows_Wq = 2 * (int64_t)e->H; out->cols_Wq = e->D; out->ld_Wq = e->ldD;
    }
    out->off_scalars = out->off_Wq + out->rows_Wq * out->ld_Wq;
    out->n_scalars = 16;
    out->total = out->off_scalars + out->n_scalars;
    return PET_OK;
}

extern "C" int pet_select_hprimes(pet_engine *e, const pet_params *p, int64_t *cand_out, void *stream) {
    if (!e) { set_error("null engine"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    pet_anneal a{1.0, 0.0, 0};
    PET_CHECK(sweep(e, &a, p, GLF_SELECT | GLF_SELECT_ONLY, 0, nullptr, 0, false, nullptr, nullptr, st));
    if (cand_out) {
        int64_t count = e->n * e->Hp;
        if (is_device_ptr(cand_out)) PET_CHECK(launch_cand_to_i64(cand_out, e->cand, count, st));
        else {
            if (count > e->stage_i64_count) {
                cudaStreamSynchronize(st);
                free_dev(e->stage_i64); e->stage_i64 = nullptr; e->stage_i64_count = 0;
                PET_CHECK(dev_alloc(&e->stage_i64, count));
                e->stage_i64_count = count;
            }
            PET_CHECK(launch_cand_to_i64(e->stage_i64, e->cand, count, st));
            PET_CUDA(cudaMemcpyAsync(cand_out, e->stage_i64, count * 8, cudaMemcpyDeviceToHost, st));
            PET_CUDA(cudaStreamSynchronize(st));
        }
    }
    return PET_OK;
}

extern "C" int pet_set_candidates(pet_engine *e, const int64_t *cand, void *stream) {
    if (!e || !cand) { set_error("pet_set_candidates: null argument"); return PET_EINVAL; }
    if (e->n <= 0) { set_error("no data bound"); return PET_ESTATE; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    int64_t count = e->n * e->Hp;
    const int64_t *src = cand;
    if (!is_device_ptr(cand)) {
        if (count > e->stage_i64_count) {
            cudaStreamSynchronize(st);
            free_dev(e->stage_i64); e->stage_i64 = nullptr; e->stage_i64_count = 0;
            PET_CHECK(dev_alloc(&e->stage_i64, count));
            e->stage_i64_count = count;
        }
        PET_CUDA(cudaMemcpyAsync(e->stage_i64, cand, count * 8, cudaMemcpyHostToDevice, st));
        src = e->stage_i64;
    }
    PET_CHECK(launch_cand_from_i64(e->cand, src, count, e->H, st));
    e->cand_state = 1;
    return PET_OK;
}

extern "C" int pet_e_step(pet_engine *e, const pet_anneal *a, const pet_params *p, double *logpj_out,
                          int64_t ld_logpj, void *stream) {
    if (!e || !logpj_out || ld_logpj < e->C) { set_error("pet_e_step: bad arguments"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    PET_CHECK(sweep(e, a, p, GLF_WRITE_LOGPJ | GLF_LSE_ONLY, 0, logpj_out, ld_logpj, true, nullptr, nullptr, st));
    if (!is_device_ptr(logpj_out)) PET_CUDA(cudaStreamSynchronize(st));
    return PET_OK;
}

extern "C" int pet_posterior_topk(pet_engine *e, const double *logpj_dev, int64_t ld_logpj, int32_t topK, int32_t logprob,
                                  int32_t *idx_out_dev, double *p_out_dev, double *m_out_dev, void *stream) {
    if (!e || !logpj_dev || !idx_out_dev || !p_out_dev || ld_logpj < e->C) { set_error("pet_posterior_topk: bad arguments"); return PET_EINVAL; }
    if (!is_device_ptr(logpj_dev)) { set_error("pet_posterior_topk: logpj must be a device pointer"); return PET_EINVAL; }
    if (e->n <= 0 || e->cand_state == 0) { set_error("pet_posterior_topk: bind data and candidates first"); return PET_ESTATE; }
    PET_CUDA(cudaSetDevice(e->device));
    // marginals exist for the layouts [null | singleton blocks | states]: base class (BSC, MCA, MMCA) and DSC
    const int binary_layout = (e->gls.has_null && e->C == 1 + (int64_t)e->gls.n_blocks * e->H + e->ss.S) ? 1 : 0;
    if (m_out_dev && !binary_layout) { set_error("pet_posterior_topk: marginals need the [null | singletons | states] layout"); return PET_EINVAL; }
    return launch_infer(e->gls, (int)e->C, binary_layout, e->cand, logpj_dev, ld_logpj, e->n, topK, logprob, idx_out_dev,
                        p_out_dev, m_out_dev, e->sm_count, (cudaStream_t)stream);
}

extern "C" int pet_log_denominators(pet_engine *e, const pet_anneal *a, const pet_params *p, const double *logpj,
                                    int64_t ld_logpj, int32_t flags, double *logdenom_out_dev, void *stream) {
    if (!e) { set_error("null engine"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    int kf = GLF_LSE_ONLY | (logpj ? GLF_READ_LOGPJ : 0) | ((flags & PASS_SELECT) ? GLF_SELECT : 0);
    PET_CHECK(sweep(e, a, p, kf, flags, logpj, ld_logpj, false, nullptr, nullptr, st));
    if (logdenom_out_dev) PET_CUDA(cudaMemcpyAsync(logdenom_out_dev, e->lse, e->n * 8, cudaMemcpyDeviceToDevice, st));
    return PET_OK;
}

extern "C" int pet_m_step_stats(pet_engine *e, const pet_anneal *a, const pet_params *p, const double *logpj,
                                int64_t ld_logpj, int32_t flags, int32_t use_cut, const double *cut_dev,
                                double *stats_dev, void *stream) {
    if (!e || !stats_dev) { set_error("pet_m_step_stats: null argument"); return PET_EINVAL; }
    if (use_cut && !cut_dev) { set_error("pet_m_step_stats: use_cut without a cut value"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    int kf = (logpj ? GLF_READ_LOGPJ : 0) | ((flags & PASS_SELECT) ? GLF_SELECT : 0) | (use_cut ? GLF_USE_CUT : 0);
    return sweep(e, a, p, kf, flags, logpj, ld_logpj, false, cut_dev, stats_dev, st);
}

extern "C" int pet_kth_largest(pet_engine *e, const double *vals_dev, int64_t n, int64_t k, double *out_dev,
                               void *stream) {
    if (!e || !vals_dev || !out_dev) { set_error("pet_kth_largest: null argument"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    e->timer.begin(ST_KSEL, st);
    PET_CHECK(kth_largest(vals_dev, n, k, out_dev, e->ksel_state, e->sm_count, st));
    e->timer.end(st);
    return PET_OK;
}

extern "C" const double *pet_log_denominators_ptr(const pet_engine *e) { return e ? e->lse : nullptr; }

extern "C" int pet_m_step_solve(pet_engine *e, const pet_params *p, const double *stats_dev, double *W_new,
                                int32_t *info_host, void *stream) {
    if (!e || !stats_dev || !W_new) { set_error("pet_m_step_solve: null argument"); return PET_EINVAL; }
    (void)p;
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    pet_stats_layout lay;
    pet_stats_layout_get(e, &lay);
    if (e->model == PET_MODEL_MCA || e->model == PET_MODEL_MMCA) {
        // element-wise update; needs the OLD W (inertia term / W^2 factors): Wt still holds it
        e->timer.begin(ST_SOLVE, st);
        const double *Wpm = stats_dev + lay.off_Wq, *Wqm = Wpm + (int64_t)e->H * e->ldD;
        PET_CHECK(launch_mca_update(stats_dev + lay.off_Wp, e->ldH, Wpm, Wqm, e->ldD, e->Wt, e->ldY, e->D, e->H,
                                    e->model == PET_MODEL_MMCA, 1e-4, e->solveB, e->ldH, st));
        e->timer.end(st);
        cudaMemcpyKind kind2 = is_device_ptr(W_new) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        PET_CUDA(cudaMemcpy2DAsync(W_new, size_t(e->H) * 8, e->solveB, e->ldH * 8, size_t(e->H) * 8, e->D, kind2, st));
        PET_CUDA(cudaStreamSynchronize(st));
        if (info_host) info_host[0] = 0;
        return PET_OK;
    }
    e->timer.begin(ST_SOLVE, st);
    PET_CUDA(cudaMemcpyAsync(e->solveA, stats_dev + lay.off_Wq, (int64_t)e->H * e->ldH * 8, cudaMemcpyDeviceToDevice, st));
    PET_CUDA(cudaMemcpyAsync(e->solveB, stats_dev + lay.off_Wp, (int64_t)e->D * e->ldH * 8, cudaMemcpyDeviceToDevice, st));
    if (e->gls.diag_from_colsum)   // binary: <s_h s_h> = <s_h>  (bsc_et.py:350,356-358)
        PET_CHECK(launch_add_diag(e->solveA, e->ldH, stats_dev + lay.off_Wp + (int64_t)e->D * e->ldH, e->H, st));
    PET_CHECK(spd_solve_right(e->H, e->D, e->solveA, e->ldH, e->solveB, e->ldH, e->solve_work, st));
    e->timer.end(st);
    cudaMemcpyKind kind = is_device_ptr(W_new) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    PET_CUDA(cudaMemcpy2DAsync(W_new, size_t(e->H) * 8, e->solveB, e->ldH * 8, size_t(e->H) * 8, e->D, kind, st));
    double scal[2] = {0, 0};
    PET_CUDA(cudaMemcpyAsync(scal, e->solve_work + (int64_t)e->H * e->ldH + round_up(e->H, 2), 16, cudaMemcpyDeviceToHost, st));
    PET_CUDA(cudaStreamSynchronize(st));
    if (info_host) info_host[0] = (int32_t)scal[1];
    return PET_OK;
}


// ---- GSC (spike-and-slab) ----------------------------------------------------------------------
extern "C" int pet_rowdot(int64_t n, int32_t D, const double *A_dev, int64_t lda, const double *B_dev, int64_t ldb, double *out_dev,
                          int64_t out_stride, void *stream);

// inverse of a (D,D) row-major matrix on the host, Gauss-Jordan with partial pivoting (np.linalg.inv of gsc_et.py:416)
static bool host_inverse(std::vector<double> &a, int D) {
    std::vector<double> inv((size_t)D * D, 0.0);
    for (int i = 0; i < D; ++i) inv[(size_t)i * D + i] = 1.0;
    for (int c = 0; c < D; ++c) {
        int piv = c;
        for (int r = c + 1; r < D; ++r) if (fabs(a[(size_t)r * D + c]) > fabs(a[(size_t)piv * D + c])) piv = r;
        if (a[(size_t)piv * D + c] == 0.0) return false;
        if (piv != c)
            for (int k = 0; k < D; ++k) { std::swap(a[(size_t)piv * D + k], a[(size_t)c * D + k]); std::swap(inv[(size_t)piv * D + k], inv[(size_t)c * D + k]); }
        const double d = 1.0 / a[(size_t)c * D + c];
        for (int k = 0; k < D; ++k) { a[(size_t)c * D + k] *= d; inv[(size_t)c * D + k] *= d; }
        for (int r = 0; r < D; ++r) {
            if (r == c) continue;
            const double f = a[(size_t)r * D + c];
            if (f == 0.0) continue;
            for (int k = 0; k < D; ++k) { a[(size_t)r * D + k] -= f * a[(size_t)c * D + k]; inv[(size_t)r * D + k] -= f * inv[(size_t)c * D + k]; }
        }
    }
    a.swap(inv);
    return true;
}

extern "C" int pet_gsc_layout_get(const pet_engine *e, pet_gsc_layout *out) {
    if (!e || !out) { set_error("pet_gsc_layout_get: null argument"); return PET_EINVAL; }
    const int64_t H = e->H, ld = e->ldH;
    out->ld = ld;
    out->off_A = 0;                                   // (D+1, H): Y^T <sz>, row D = sum_n <sz>
    out->off_Mssz = (e->D + 1) * ld;                  // (H,H)  sum_n <s> (x) <sz>
    out->off_Mout = out->off_Mssz + H * ld;           // (H,H)  sum_n <sz> (x) <sz>
    out->off_ss = out->off_Mout + H * ld;             // (H,H)  sum_n <s s^T>, off-diagonal part
    out->off_szsz = out->off_ss + H * ld;             // (H,H)  sum_n <sz sz^T>, multi-cause part
    out->off_sum_s = out->off_szsz + H * ld;          // (H,)   sum_n <s>     (= diagonal of sum <s s^T>)
    out->off_sum_sz2 = out->off_sum_s + ld;           // (H,)   singleton part of diag sum <sz sz^T>
    out->off_ysq = out->off_sum_sz2 + ld;             // (D,)   sum_n y_nd^2
    out->off_scalars = out->off_ysq + e->ldY;         // [0] = datapoints
    out->off_yyT = out->off_scalars + 8;              // (D, ldY) sum_n y y^T ('full' noise covariance only)
    out->ld_yyT = e->ldY;
    out->total = out->off_yyT + (int64_t)e->D * e->ldY;
    return PET_OK;
}

static int prepare_gsc(pet_engine *e, const pet_gsc_params *p, cudaStream_t st) {
    if (!p || !p->W || !p->pi_host || !p->mu_host || !p->psi_sq_host || !p->sigma_sq_host) { set_error("bad GSC parameters"); return PET_EINVAL; }
    const bool full = (p->sigma_sq_type == 2);
    e->timer.begin(ST_PREPARE, st);
    pet_params pw;
    memset(&pw, 0, sizeof(pw));
    pw.W = p->W; pw.ldW = p->ldW; pw.pi_host = p->pi_host; pw.n_pi = e->H;
    PET_CHECK(load_W(e, &pw, st));
    // inverse noise variances: B = Sigma^-1 (gsc_et.py:415-423)
    std::vector<double> host((size_t)std::max<int64_t>(e->ldY, 8 * e->ldH), 0.0);
    double bscalar = 0.0;
    const double *bdiag = nullptr;
    if (p->sigma_sq_type == 1) {
        for (int d = 0; d < e->D; ++d) host[d] = 1.0 / p->sigma_sq_host[d];
        PET_CUDA(cudaMemcpyAsync(e->bdiag, host.data(), e->D * 8, cudaMemcpyHostToDevice, st));
        PET_CUDA(cudaStreamSynchronize(st));
        bdiag = e->bdiag;
    } else bscalar = 1.0 / p->sigma_sq_host[0];
    if (full) {
        // B = Sigma^-1 dense (symmetric Sigma assumed): Wt2 = (B W)^T by one GEMM
        std::vector<double> Bh(p->sigma_sq_host, p->sigma_sq_host + (size_t)e->D * e->D);
        if (!host_inverse(Bh, e->D)) { set_error("GSC: sigma_sq is singular"); return PET_EINVAL; }
        if (!e->Bfull) {
            PET_CHECK(dev_alloc(&e->Bfull, (int64_t)e->D * e->ldY));
            PET_CUDA(cudaMemset(e->Bfull, 0, (int64_t)e->D * e->ldY * 8));
        }
        if (!e->gsc_T) PET_CHECK(dev_alloc(&e->gsc_T, e->chunk_rows * e->ldY));
        PET_CUDA(cudaMemcpy2DAsync(e->Bfull, e->ldY * 8, Bh.data(), size_t(e->D) * 8, size_t(e->D) * 8, e->D, cudaMemcpyHostToDevice, st));
        PET_CUDA(cudaStreamSynchronize(st));
        PET_CHECK(dgemm_kk(e->H, e->D, e->D, e->Wt, e->ldY, e->Bfull, e->ldY, e->Wt2, e->ldY, 1.0, 0, st));
    } else {
        PET_CHECK(launch_scale_rows(e->Wt2, e->Wt, e->ldY, e->H, e->D, bdiag, bscalar, st));
    }
    PET_CHECK(dgemm_kk(e->H, e->H, e->D, e->Wt, e->ldY, e->Wt2, e->ldY, e->G, e->ldH, 1.0, 0, st));   // W^T Sigma^-1 W
    // psi_sq (H,H), pi, mu to the device
    PET_CUDA(cudaMemcpy2DAsync(e->psi_dev, e->ldH * 8, p->psi_sq_host, size_t(e->H) * 8, size_t(e->H) * 8, e->H, cudaMemcpyHostToDevice, st));
    double *tab = e->gsc_tab;
    PET_CUDA(cudaMemcpyAsync(tab + 4 * e->ldH, p->mu_host, e->H * 8, cudaMemcpyHostToDevice, st));
    PET_CUDA(cudaMemcpyAsync(tab + 5 * e->ldH, p->pi_host, e->H * 8, cudaMemcpyHostToDevice, st));
    PET_CUDA(cudaStreamSynchronize(st));       // the host vectors may be temporaries of the caller
    PET_CHECK(launch_gsc_tables(e->G, e->ldH, e->psi_dev, e->ldH, tab + 5 * e->ldH, tab + 4 * e->ldH, e->H, tab, tab + e->ldH,
                                tab + 2 * e->ldH, tab + 3 * e->ldH, st));
    e->timer.end(st);
    // y^T Sigma^-1 y for every datapoint (sigma changes every iteration)
    const int64_t nchunks = (int64_t)e->chunk_start.size() - 1;
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t r0 = e->chunk_start[c], rows = e->chunk_start[c + 1] - r0;
        PET_CHECK(ensure_chunk_inputs(e, c, r0, rows, st));
        if (full) {      // y^T B y = rowdot(Y B, Y)
            PET_CHECK(dgemm_kk(rows, e->D, e->D, e->Y + r0 * e->ldY, e->ldY, e->Bfull, e->ldY, e->gsc_T, e->ldY, 1.0, 0, st));
            PET_CHECK(pet_rowdot(rows, e->D, e->gsc_T, e->ldY, e->Y + r0 * e->ldY, e->ldY, e->yyw + r0, 1, st));
        } else {
            PET_CHECK(launch_weighted_rownorm(e->Y + r0 * e->ldY, e->ldY, rows, e->D, bdiag, bscalar, e->yyw + r0, st));
        }
    }
    e->yy_valid = true;
    return PET_OK;
}

// flags: GSCF_*.  Dense outputs are device pointers (n,H) / (n,H,H); dst maps datapoint -> output row.
static int sweep_gsc(pet_engine *e, const pet_anneal *a, const pet_gsc_params *p, int flags, const int64_t *dst_dev,
                     double *xs, double *xss, double *xsz, double *xszsz, double *stats_dev, cudaStream_t st,
                     double *logpj_dev = nullptr, int64_t ld_logpj = 0) {
    if (e->n <= 0) { set_error("no data bound (pet_set_data)"); return PET_ESTATE; }
    if (!(flags & GSCF_SELECT) && e->cand_state == 0) { set_error("no candidates: run select_Hprimes first"); return PET_ESTATE; }
    if (!a || !(a->T > 0.0)) { set_error("annealing temperature T must be > 0"); return PET_EINVAL; }
    PET_CHECK(prepare_gsc(e, p, st));
    GSCArgs g;
    memset(&g, 0, sizeof(g));
    g.st = e->gls;
    double *tab = e->gsc_tab;
    g.tb.g = tab; g.tb.ilam = tab + e->ldH; g.tb.lcdet = tab + 2 * e->ldH; g.tb.logit = tab + 3 * e->ldH; g.tb.mu = tab + 4 * e->ldH;
    g.flags = flags;
    g.beta = 1.0 / a->T;
    g.yyw = e->yyw; g.G = e->G; g.psi = e->psi_dev; g.cand = e->cand;
    g.dst = dst_dev; g.xpt_s = xs; g.xpt_ss = xss; g.xpt_sz = xsz; g.xpt_szsz = xszsz;
    g.logpj = logpj_dev; g.ld_logpj = ld_logpj;
    pet_gsc_layout lay;
    pet_gsc_layout_get(e, &lay);
    if (flags & GSCF_STATS) {
        if (!stats_dev) { set_error("stats buffer is null"); return PET_EINVAL; }
        PET_CUDA(cudaMemsetAsync(stats_dev, 0, lay.total * 8, st));
        g.sum_ss = stats_dev + lay.off_ss; g.sum_szsz = stats_dev + lay.off_szsz;
    }
    const int64_t nchunks = (int64_t)e->chunk_start.size() - 1;
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t r0 = e->chunk_start[c], rows = e->chunk_start[c + 1] - r0;
        double *yw = e->yw_all ? e->YW + r0 * e->ldH : e->YW;
        e->timer.begin(ST_SCORE, st);
        PET_CHECK(dgemm_kk(rows, e->H, e->D, e->Y + r0 * e->ldY, e->ldY, e->Wt2, e->ldY, yw, e->ldH, 1.0, 0, st));
        e->timer.end(st);
        g.n_rows = rows; g.row0 = r0; g.YW = yw; g.XS = e->Sbuf; g.XSZ = e->XSZ; g.SZ2 = e->SZ2;
        e->timer.begin(ST_POST, st);
        PET_CHECK(launch_gsc_kernel(g, e->gamma, e->sm_count, st));
        e->timer.end(st);
        if (flags & GSCF_STATS) {
            e->timer.begin(ST_STATS, st);
            const double *Yc = e->Y + r0 * e->ldY;
            PET_CHECK(dgemm_mn(e->D + 1, e->H, rows, Yc, e->ldY, e->XSZ, e->ldH, stats_dev + lay.off_A, e->ldH, 1,
                               e->gemm_work, e->gemm_work_doubles, e->sm_count, st));          // gsc_et.py:613-620
            PET_CHECK(dgemm_mn(e->H, e->H, rows, e->Sbuf, e->ldH, e->XSZ, e->ldH, stats_dev + lay.off_Mssz, e->ldH, 1,
                               e->gemm_work, e->gemm_work_doubles, e->sm_count, st));          // :665
            PET_CHECK(dgemm_mn(e->H, e->H, rows, e->XSZ, e->ldH, e->XSZ, e->ldH, stats_dev + lay.off_Mout, e->ldH, 1,
                               e->gemm_work, e->gemm_work_doubles, e->sm_count, st));          // :683,:697,:711
            PET_CHECK(launch_colsum(stats_dev + lay.off_sum_s, e->Sbuf, e->ldH, rows, e->H, st));
            PET_CHECK(launch_colsum(stats_dev + lay.off_sum_sz2, e->SZ2, e->ldH, rows, e->H, st));
            PET_CHECK(launch_colsumsq(stats_dev + lay.off_ysq, Yc, e->ldY, rows, e->D, st));
            if (p->sigma_sq_type == 2) {           // sum_n y y^T for the full covariance update (gsc_et.py:679-682)
                const int sp = dgemm_mn_splits(e->D, e->D, rows, e->sm_count);
                const int64_t need = int64_t(sp) * e->D * e->ldY;
                if (need > e->gemm_work_doubles) {
                    cudaStreamSynchronize(st);
                    free_dev(e->gemm_work); e->gemm_work = nullptr; e->gemm_work_doubles = 0;
                    PET_CHECK(dev_alloc(&e->gemm_work, need));
                    e->gemm_work_doubles = need;
                }
                PET_CHECK(dgemm_mn(e->D, e->D, rows, Yc, e->ldY, Yc, e->ldY, stats_dev + lay.off_yyT, e->ldY, 1,
                                   e->gemm_work, e->gemm_work_doubles, e->sm_count, st));
            }
            e->timer.end(st);
        }
    }
    if (flags & GSCF_STATS) {
        double nloc = (double)e->n;
        PET_CUDA(cudaMemcpyAsync(stats_dev + lay.off_scalars, &nloc, 8, cudaMemcpyHostToDevice, st));
        PET_CUDA(cudaStreamSynchronize(st));
    }
    if (flags & GSCF_SELECT) e->cand_state = 1;
    mark_compute_done(e, st);
    return PET_OK;
}

static int copy_cand_out(pet_engine *e, int64_t *cand_out, cudaStream_t st) {
    int64_t count = e->n * e->Hp;
    if (is_device_ptr(cand_out)) return launch_cand_to_i64(cand_out, e->cand, count, st);
    if (count > e->stage_i64_count) {
        cudaStreamSynchronize(st);
        free_dev(e->stage_i64); e->stage_i64 = nullptr; e->stage_i64_count = 0;
        PET_CHECK(dev_alloc(&e->stage_i64, count));
        e->stage_i64_count = count;
    }
    PET_CHECK(launch_cand_to_i64(e->stage_i64, e->cand, count, st));
    PET_CUDA(cudaMemcpyAsync(cand_out, e->stage_i64, count * 8, cudaMemcpyDeviceToHost, st));
    PET_CUDA(cudaStreamSynchronize(st));
    return PET_OK;
}

extern "C" int pet_gsc_select(pet_engine *e, const pet_gsc_params *p, int64_t *cand_out, void *stream) {
    if (!e || e->model != PET_MODEL_GSC) { set_error("pet_gsc_select: not a GSC engine"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    pet_anneal a{1.0, 0.0, 0};
    PET_CHECK(sweep_gsc(e, &a, p, GSCF_SELECT | GSCF_SELECT_ONLY, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
    if (cand_out) PET_CHECK(copy_cand_out(e, cand_out, st));
    return PET_OK;
}

extern "C" int pet_gsc_compute_lpj(pet_engine *e, const pet_gsc_params *p, double *logpj_dev, int64_t ld_logpj,
                                   int64_t *cand_out, void *stream) {
    if (!e || e->model != PET_MODEL_GSC || !logpj_dev || ld_logpj < e->C || !is_device_ptr(logpj_dev)) {
        set_error("pet_gsc_compute_lpj: bad arguments (logpj must be a device array with ld >= 1 + H + S)");
        return PET_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    pet_anneal a{1.0, 0.0, 0};
    PET_CHECK(sweep_gsc(e, &a, p, GSCF_SELECT | GSCF_LOGPJ, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st, logpj_dev, ld_logpj));
    if (cand_out) PET_CHECK(copy_cand_out(e, cand_out, st));
    return PET_OK;
}

extern "C" int pet_gsc_e_step(pet_engine *e, const pet_anneal *a, const pet_gsc_params *p, const int64_t *dst_rows,
                              double *xpt_s_dev, double *xpt_ss_dev, double *xpt_sz_dev, double *xpt_szsz_dev, void *stream) {
    if (!e || e->model != PET_MODEL_GSC || !xpt_s_dev || !xpt_ss_dev || !xpt_sz_dev || !xpt_szsz_dev) { set_error("pet_gsc_e_step: bad arguments"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    const int64_t *dst = dst_rows;
    if (dst_rows && !is_device_ptr(dst_rows)) {
        if (e->n > e->dst_cap) { cudaStreamSynchronize(st); free_dev(e->dst_dev); e->dst_dev = nullptr; PET_CHECK(dev_alloc(&e->dst_dev, e->n)); e->dst_cap = e->n; }
        PET_CUDA(cudaMemcpyAsync(e->dst_dev, dst_rows, e->n * 8, cudaMemcpyHostToDevice, st));
        PET_CUDA(cudaStreamSynchronize(st));
        dst = e->dst_dev;
    }
    return sweep_gsc(e, a, p, GSCF_DENSE, dst, xpt_s_dev, xpt_ss_dev, xpt_sz_dev, xpt_szsz_dev, nullptr, st);
}

extern "C" int pet_gsc_stats(pet_engine *e, const pet_anneal *a, const pet_gsc_params *p, int32_t flags, double *stats_dev,
                             void *stream) {
    if (!e || e->model != PET_MODEL_GSC || !stats_dev) { set_error("pet_gsc_stats: bad arguments"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CUDA(cudaSetDevice(e->device));
    return sweep_gsc(e, a, p, GSCF_STATS | ((flags & PASS_SELECT) ? GSCF_SELECT : 0), nullptr, nullptr, nullptr, nullptr, nullptr, stats_dev, st);
}

extern "C" int pet_data_sum(pet_engine *e, int32_t use_cut, const double *cut_dev, double *out_dev, void *stream) {
    if (!e || !out_dev || (use_cut && !cut_dev)) { set_error("pet_data_sum: bad arguments"); return PET_EINVAL; }
    if (e->n <= 0) return PET_OK;
    if (e->upload_pending) { set_error("pet_data_sum: run a pass over the shard first"); return PET_ESTATE; }
    PET_CUDA(cudaSetDevice(e->device));
    return launch_colsum_kept(out_dev, e->Y, e->ldY, e->n, e->D, e->lse, use_cut ? cut_dev : nullptr,
                              e->model == PET_MODEL_DSC ? 1 : 0, (cudaStream_t)stream);
}

extern "C" int pet_colsum(int64_t rows, int64_t cols, const double *M_dev, int64_t ld, double *out_dev, void *stream) {
    if (!M_dev || !out_dev || cols > ld) { set_error("pet_colsum: bad arguments"); return PET_EINVAL; }
    return launch_colsum(out_dev, M_dev, ld, rows, (int)cols, (cudaStream_t)stream);
}

// ---- exported building blocks ------------------------------------------------------------
extern "C" int pet_dgemm_kk(int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B,
                            int64_t ldb, double *C, int64_t ldc, double alpha, double beta, void *stream) {
    if (beta != 0.0 && beta != 1.0) { set_error("pet_dgemm_kk: beta must be 0 or 1"); return PET_EINVAL; }
    return dgemm_kk(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta == 1.0, (cudaStream_t)stream);
}

extern "C" int pet_dgemm_mn(int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B,
                            int64_t ldb, double *C, int64_t ldc, int32_t accumulate, double *work,
                            int64_t work_doubles, void *stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!C) return dgemm_mn_splits(M, N, K, sms);
    return dgemm_mn(M, N, K, A, lda, B, ldb, C, ldc, accumulate, work, work_doubles, sms, (cudaStream_t)stream);
}

extern "C" int pet_spd_solve_right(int64_t n, int64_t m, double *A, int64_t lda, double *B, int64_t ldb,
                                   double *work, int32_t *info_host, void *stream) {
    if (!A || !B || !work || (lda & 1) || (ldb & 1)) { set_error("pet_spd_solve_right: bad arguments"); return PET_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    PET_CHECK(spd_solve_right(n, m, A, lda, B, ldb, work, st));
    double scal[2] = {0, 0};
    PET_CUDA(cudaMemcpyAsync(scal, work + n * lda + round_up(n, 2), 16, cudaMemcpyDeviceToHost, st));
    PET_CUDA(cudaStreamSynchronize(st));
    if (info_host) info_host[0] = (int32_t)scal[1];
    return PET_OK;
}

extern "C" int64_t pet_spd_solve_work_doubles(int64_t n, int64_t lda) { return spd_solve_work_doubles(n, lda); }
