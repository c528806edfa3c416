// FP64 tensor-core GEMM tiles for the two contractions of the ET hot path.
//
//   KK:  C(M,N) = alpha * A(M,K) . B(N,K)^T (+ C)     both operands K-contiguous
//        -> score GEMM  YW = Y . W      (replaces the per-datapoint np.inner / np.dot
//           loops of bsc_et.py:110-112,176,180-184), Gram G = W^T W, and the rank-k
//           updates of the blocked Cholesky.
//   MN:  C(M,N) = A(K,M)^T . B(K,N) (+ C)             reduction over ROWS, split-K
//        -> statistics GEMM  Wp^T = Y^T . <S>   (replaces my_Wp += outer(...) of
//           bsc_et.py:349,355,363).
//
// tcgen05 has no f64 kind; the FP64 tensor path on sm_100a is the warp-level
// mma.sync.m8n8k4.f64 (DMMA).  Operands are staged global->shared with 16-byte cp.async
// in a 3-stage ring; shared tiles are padded (+4 doubles per row) so that the per-lane
// 8-byte fragment loads of a half-warp hit 16 distinct double-banks.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace pet {

struct GemmArgs {
    int64_t M, N, K;
    const double *A; int64_t lda;
    const double *B; int64_t ldb;
    double *C; int64_t ldc;
    double alpha;
    int accumulate;        // C += instead of C =
    int64_t k_per_split;   // MN only: rows of K handled per blockIdx.z
    int64_t split_stride;  // MN only: doubles between the partial C of consecutive splits
};

template <bool KK, int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES>
struct GemmCfg {
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;   // warp tile
    static constexpr int FM = WM / 8, FN = WN / 8;               // 8x8 fragments per warp
    static constexpr int SA = KK ? (BK + 4) : (BM + 4);           // padded smem row strides
    static constexpr int SB = KK ? (BK + 4) : (BN + 4);
    static constexpr int A_ROWS = KK ? BM : BK;
    static constexpr int B_ROWS = KK ? BN : BK;
    static constexpr int A_STAGE = A_ROWS * SA;
    static constexpr int B_STAGE = B_ROWS * SB;
    static constexpr size_t SMEM = size_t(STAGES) * (A_STAGE + B_STAGE) * sizeof(double);
    static_assert(WM % 8 == 0 && WN % 8 == 0 && BK % 4 == 0, "tile shape");
};

// Stage one (A or B) tile.  KK: ROWS x BK with k contiguous; MN: BK x COLS with m/n contiguous.
template <bool KK, int ROWS_OR_COLS, int BK, int STRIDE, int THREADS>
__device__ __forceinline__ void load_tile(double *s, const double *g, int64_t ld, int64_t mn0,
                                          int64_t MN, int64_t k0, int64_t Kend) {
    if (KK) {
        constexpr int CPR = BK / 2;                  // 16-byte chunks per row
        constexpr int TOTAL = ROWS_OR_COLS * CPR;
#pragma unroll
        for (int c = threadIdx.x; c < TOTAL; c += THREADS) {
            int r = c / CPR, kc = (c % CPR) * 2;
            int64_t row = mn0 + r, k = k0 + kc;
            int64_t rem = (Kend - k) * 8;
            int bytes = (row < MN && rem > 0) ? (rem >= 16 ? 16 : 8) : 0;
            const double *src = bytes ? (g + row * ld + k) : g;
            cp_async16(s + r * STRIDE + kc, src, bytes);
        }
    } else {
        constexpr int CPR = ROWS_OR_COLS / 2;
        constexpr int TOTAL = BK * CPR;
#pragma unroll
        for (int c = threadIdx.x; c < TOTAL; c += THREADS) {
            int r = c / CPR, mc = (c % CPR) * 2;
            int64_t k = k0 + r, col = mn0 + mc;
            int64_t rem = (MN - col) * 8;
            int bytes = (k < Kend && rem > 0) ? (rem >= 16 ? 16 : 8) : 0;
            const double *src = bytes ? (g + k * ld + col) : g;
            cp_async16(s + r * STRIDE + mc, src, bytes);
        }
    }
}

template <bool KK, int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES, int MINB>
__global__ void __launch_bounds__(WARPS_M *WARPS_N * 32, MINB)
dgemm_kernel(GemmArgs g) {
    using Cfg = GemmCfg<KK, BM, BN, BK, WARPS_M, WARPS_N, STAGES>;
    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + STAGES * Cfg::A_STAGE;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = (warp / WARPS_N) * Cfg::WM, wn = (warp % WARPS_N) * Cfg::WN;

    const int64_t tiles_n = (g.N + BN - 1) / BN;
    const int64_t m0 = (blockIdx.x / tiles_n) * BM, n0 = (blockIdx.x % tiles_n) * BN;

    int64_t kbeg = 0, kend = g.K;
    double *C = g.C;
    if (!KK) {
        kbeg = int64_t(blockIdx.z) * g.k_per_split;
        kend = min(g.K, kbeg + g.k_per_split);
        C += int64_t(blockIdx.z) * g.split_stride;
    }
    const int64_t ksteps = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

    double acc[Cfg::FM][Cfg::FN][2];
#pragma unroll
    for (int i = 0; i < Cfg::FM; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue: fill STAGES-1 stages
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ksteps) {
            load_tile<KK, BM, BK, Cfg::SA, Cfg::THREADS>(As + s * Cfg::A_STAGE, g.A, g.lda, m0, g.M,
                                                         kbeg + int64_t(s) * BK, kend);
            load_tile<KK, BN, BK, Cfg::SB, Cfg::THREADS>(Bs + s * Cfg::B_STAGE, g.B, g.ldb, n0, g.N,
                                                         kbeg + int64_t(s) * BK, kend);
        }
        cp_async_commit();
    }

    for (int64_t ks = 0; ks < ksteps; ++ks) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        // prefetch stage ks+STAGES-1 into the slot freed in the previous iteration
        {
            int64_t nk = ks + STAGES - 1;
            if (nk < ksteps) {
                int slot = int(nk % STAGES);
                load_tile<KK, BM, BK, Cfg::SA, Cfg::THREADS>(As + slot * Cfg::A_STAGE, g.A, g.lda, m0,
                                                             g.M, kbeg + nk * BK, kend);
                load_tile<KK, BN, BK, Cfg::SB, Cfg::THREADS>(Bs + slot * Cfg::B_STAGE, g.B, g.ldb, n0,
                                                             g.N, kbeg + nk * BK, kend);
            }
            cp_async_commit();
        }
        const double *a_s = As + int(ks % STAGES) * Cfg::A_STAGE;
        const double *b_s = Bs + int(ks % STAGES) * Cfg::B_STAGE;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double af[Cfg::FM], bf[Cfg::FN];
#pragma unroll
            for (int i = 0; i < Cfg::FM; ++i)
                af[i] = KK ? a_s[(wm + i * 8 + gid) * Cfg::SA + kk * 4 + tig]
                           : a_s[(kk * 4 + tig) * Cfg::SA + wm + i * 8 + gid];
#pragma unroll
            for (int j = 0; j < Cfg::FN; ++j)
                bf[j] = KK ? b_s[(wn + j * 8 + gid) * Cfg::SB + kk * 4 + tig]
                           : b_s[(kk * 4 + tig) * Cfg::SB + wn + j * 8 + gid];
#pragma unroll
            for (int i = 0; i < Cfg::FM; ++i)
#pragma unroll
                for (int j = 0; j < Cfg::FN; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: fragments straight to global (each quad writes 64 contiguous bytes per row)
    const bool vec_ok = ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < Cfg::FM; ++i) {
        int64_t row = m0 + wm + i * 8 + gid;
        if (row >= g.M) continue;
#pragma unroll
        for (int j = 0; j < Cfg::FN; ++j) {
            int64_t col = n0 + wn + j * 8 + 2 * tig;
            if (col >= g.N) continue;
            double *dst = C + row * g.ldc + col;
            double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
            if (col + 1 < g.N) {
                if (vec_ok) {
                    double2 o;
                    if (g.accumulate) {
                        o = *reinterpret_cast<double2 *>(dst);
                        o.x += v0; o.y += v1;
                    } else { o.x = v0; o.y = v1; }
                    *reinterpret_cast<double2 *>(dst) = o;
                } else {
                    if (g.accumulate) { dst[0] += v0; dst[1] += v1; } else { dst[0] = v0; dst[1] = v1; }
                }
            } else {
                if (g.accumulate) dst[0] += v0; else dst[0] = v0;
            }
        }
    }
}

// out = (out +) sum_z part[z]; fixed summation order => deterministic statistics.
// Only the M x N payload is touched (the ld padding of `out` keeps its zeros).
__global__ void splitk_reduce_kernel(double *out, const double *part, int64_t M, int64_t N, int64_t ldc,
                                     int splits, int64_t stride, int accumulate) {
    int64_t half = (N + 1) / 2;
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= M * half) return;
    int64_t row = i / half, col = (i % half) * 2;
    int64_t off = row * ldc + col;
    if (col + 1 < N && ((ldc & 1) == 0)) {
        double2 s = make_double2(0., 0.);
        for (int z = 0; z < splits; ++z) {
            double2 v = *reinterpret_cast<const double2 *>(part + z * stride + off);
            s.x += v.x; s.y += v.y;
        }
        double2 *o = reinterpret_cast<double2 *>(out + off);
        if (accumulate) { double2 c = *o; s.x += c.x; s.y += c.y; }
        *o = s;
    } else {
        for (int64_t c = col; c < N && c < col + 2; ++c) {
            double s = 0.;
            for (int z = 0; z < splits; ++z) s += part[z * stride + row * ldc + c];
            out[row * ldc + c] = accumulate ? out[row * ldc + c] + s : s;
        }
    }
}

template <bool KK, int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES, int MINB = 1>
static int launch_gemm(const GemmArgs &g, int splits, cudaStream_t st) {
    using Cfg = GemmCfg<KK, BM, BN, BK, WARPS_M, WARPS_N, STAGES>;
    auto kern = dgemm_kernel<KK, BM, BN, BK, WARPS_M, WARPS_N, STAGES, MINB>;
    static bool configured = false;   // per instantiation; attribute is per-device but we use one device per process
    if (!configured) {
        PET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM)));
        configured = true;
    }
    int64_t tiles = ceil_div(g.M, BM) * ceil_div(g.N, BN);
    if (tiles <= 0) return PET_OK;
    dim3 grid((unsigned)tiles, 1, (unsigned)splits);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(g);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

int dgemm_kk(int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B,
             int64_t ldb, double *C, int64_t ldc, double alpha, int accumulate, cudaStream_t st) {
    if (M <= 0 || N <= 0) return PET_OK;
    if (!aligned16(A) || !aligned16(B) || (lda & 1) || (ldb & 1)) {
        set_error("dgemm_kk: operands must be 16-byte aligned with even leading dimensions");
        return PET_EINVAL;
    }
    GemmArgs g{M, N, K, A, lda, B, ldb, C, ldc, alpha, accumulate, 0, 0};
    static const int variant = env_int("PET_GEMM_KK", 1);
    if (M * N >= int64_t(128) * 128 * 64 && N > 64) {
        switch (variant) {
            case 0: return launch_gemm<true, 128, 128, 16, 2, 4, 3>(g, 1, st);
            case 2: return launch_gemm<true, 128, 64, 32, 2, 2, 2, 2>(g, 1, st);
            case 3: return launch_gemm<true, 128, 128, 16, 2, 4, 4>(g, 1, st);
            default: return launch_gemm<true, 128, 64, 16, 2, 2, 3, 2>(g, 1, st);   // two CTAs per SM hide each other's prologue/epilogue
        }
    }
    return launch_gemm<true, 64, 64, 16, 2, 2, 3>(g, 1, st);
}

static int mn_variant() {
    static const int v = env_int("PET_GEMM_MN", 1);
    return v;
}
// tile height (rows of C = M) the MN kernel uses for a problem: 136 = 17 fragments fits D+1 = 677 -> 680 exactly
static int64_t mn_tile_m(int64_t M, int64_t N) {
    if (!(M > 64 && N > 64)) return 64;
    if (mn_variant() == 0) return 128;
    int64_t w128 = ceil_div(M, 128) * 128, w136 = ceil_div(M, 136) * 136;
    return (w136 < w128) ? 136 : 128;
}

// how many K-splits the MN kernel uses for a given problem (also sizes the workspace): fill whole waves
int dgemm_mn_splits(int64_t M, int64_t N, int64_t K, int sm_count) {
    bool big = (M > 64 && N > 64);
    int64_t bm = mn_tile_m(M, N), bn = (big && !(bm == 136 && mn_variant() == 2)) ? 128 : 64;
    int64_t tiles = ceil_div(M, bm) * ceil_div(N, bn);
    int64_t max_by_k = std::max<int64_t>(1, K / 512);               // >= 512 rows per split
    // time model in units of one row of K: the CTAs run in ceil(tiles s / SMs) waves of K / s rows each, and every split
    // costs one more slab for splitk_reduce to read.  (A single 64 x 64 output tile -- the H x H statistics of GSC --
    // spreads over all the SMs this way; the previous rule left it on one CTA.)
    int best = 1;
    double best_cost = 1e300;
    for (int64_t s = 1; s <= std::min<int64_t>(max_by_k, 2 * sm_count); ++s) {
        const double waves = ceil(double(tiles * s) / sm_count);
        const double cost = waves * double(ceil_div(std::max<int64_t>(K, 1), s)) + 24.0 * double(s);
        if (cost < best_cost) { best_cost = cost; best = int(s); }
    }
    return best;
}

int dgemm_mn(int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B,
             int64_t ldb, double *C, int64_t ldc, int accumulate, double *work, int64_t work_doubles,
             int sm_count, cudaStream_t st) {
    if (M <= 0 || N <= 0) return PET_OK;
    if (!aligned16(A) || !aligned16(B) || (lda & 1) || (ldb & 1)) {
        set_error("dgemm_mn: operands must be 16-byte aligned with even leading dimensions");
        return PET_EINVAL;
    }
    int splits = dgemm_mn_splits(M, N, K, sm_count);
    int64_t stride = M * ldc;
    if (splits > 1 && (work == nullptr || work_doubles < stride * splits))      // as many splits as the workspace holds
        splits = (work && stride > 0) ? int(std::max<int64_t>(1, work_doubles / stride)) : 1;
    int64_t kps = round_up(ceil_div(std::max<int64_t>(K, 1), splits), 16);
    const int64_t bm = mn_tile_m(M, N);
    auto launch = [&](const GemmArgs &g, int sp) -> int {
        if (bm == 64) return launch_gemm<false, 64, 64, 16, 2, 2, 3>(g, sp, st);
        if (bm == 136 && mn_variant() == 2) return launch_gemm<false, 136, 64, 16, 1, 4, 3, 2>(g, sp, st);
        if (bm == 136) return launch_gemm<false, 136, 128, 16, 1, 8, 3>(g, sp, st);
        return launch_gemm<false, 128, 128, 16, 2, 4, 3>(g, sp, st);
    };
    if (splits == 1) {
        GemmArgs g{M, N, K, A, lda, B, ldb, C, ldc, 1.0, accumulate, kps, 0};
        return launch(g, 1);
    }
    GemmArgs g{M, N, K, A, lda, B, ldb, work, ldc, 1.0, 0, kps, stride};
    PET_CHECK(launch(g, splits));
    int threads = 256;
    int64_t blocks = ceil_div(M * ((N + 1) / 2), threads);
    splitk_reduce_kernel<<<(unsigned)blocks, threads, 0, st>>>(C, work, M, N, ldc, splits, stride, accumulate);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

}  // namespace pet
