// FP64-accurate score GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// tcgen05 has no f64 kind, and the FP64 DMMA pipe caps the path at ~35 TFLOP/s.  This kernel computes
//     C(M,N) = A(M,K) . B(N,K)^T          (both K-contiguous, float64 in / float64 out)
// with INT8 tensor-core MMAs by error-free slicing (Ozaki scheme):
//   * every row of A and of B is scaled by a power of two and cut into NS signed 7-bit slices
//     (first slice 6 bits), a_mk = 2^eA[m] * sum_i 2^-(6+7i) A_i[m,k],  |A_i| <= 64;
//   * slice products A_i . B_j^T are accumulated EXACTLY in int32 (|sum| <= (t+1) K 2^12 < 2^31),
//     all pairs with i + j = t into the same TMEM accumulator, one accumulator per t = 0..NS-1;
//   * the epilogue converts the NS accumulators to float64, weights them by 2^-(12+7t) and the row /
//     column scales (all exact powers of two) and adds them up: the only rounding is that final sum and
//     the truncation of pairs with i + j >= NS  (NS = 7: |error| <= ~5e-14 * max|C|, measured in tests).
// One CTA computes a 128 x 64 tile; per 64-byte K block all NS slices of A and B are staged once by TMA
// (SWIZZLE_64B) and reused by the NS(NS+1)/2 pair MMAs, so shared memory / L2 traffic is per slice, not
// per pair.  Warp roles: 0 = TMA producer, 1 = MMA issuer (one elected lane), 2 = TMEM allocator,
// 4..7 = epilogue (tcgen05.ld -> FP64 accumulate in registers -> global).
#include <cuda.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "ozaki.cuh"

namespace pet {

namespace oz {
constexpr int BM = 128, BN = 64, KB = 64;          // tile rows, tile cols, bytes (= int8 elements) of K per stage
constexpr int UMMA_K = 32;                         // K of one kind::i8 MMA
constexpr int THREADS = 384;                       // 4 control warps + 8 epilogue warps (two per TMEM lane quadrant)

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// same, for waits that are off the critical path (the producer runs stages ahead): back off between polls so that the
// spinning lane does not take issue slots from the epilogue warp that shares its scheduler
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITR_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONER_%=;\n\t"
        "nanosleep.u32 64;\n\t"
        "bra WAITR_%=;\n\t"
        "DONER_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// K-major SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO = 1, SBO = 8 rows * 64 B = 512 B, version 1 (Blackwell), layout type 4 (SWIZZLE_64B)
template <int KBLK>
__device__ __forceinline__ uint64_t make_desc(const void *smem_ptr) {
    // rows of KBLK bytes: SWIZZLE_64B (layout type 4, SBO = 8 x 64 B) or SWIZZLE_32B (layout type 6, SBO = 8 x 32 B)
    uint64_t addr = smem_u32(smem_ptr);
    return ((addr & 0x3FFFFull) >> 4) | (1ull << 16) | (uint64_t((8 * KBLK) >> 4) << 32) | (1ull << 46) |
           ((KBLK == 64 ? 4ull : 6ull) << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::i8: D = S32 (2 << 4), A = B = signed int8 (1 << 7,
// 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24: make_idesc_n below
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

struct Args {
    int64_t M, N;
    int kblocks;                 // K padded / 64
    int splits, kb_per_split;    // split-K: unit (tile, split) covers k blocks [split * kb_per_split, ...)
    int accumulate;              // epilogue: C += result instead of C = result
    const double *sA, *sB;       // row scales 2^eA[m], 2^eB[n]
    double *C;
    int64_t ldc, split_stride;   // split s writes C + s * split_stride
};

__device__ __forceinline__ uint32_t make_idesc_n(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(BM >> 4) << 24);
}

// KBLK: bytes of K per pipeline stage (Args.kblocks / kb_per_split count blocks of KBLK).  64: two stages of 86 KB
// (the default); 32: five stages of 43 KB -- the same bytes in flight in finer grains (measured slower, see ozaki_gemm)
// SA / SB: ring depths of the A (128-row) and the B (64-row) slice tiles.  The rings are independent, because with 7
// slices three whole stages (3 x 84 KB) do not fit: two A stages (2 x 56 KB) + four B stages (4 x 28 KB) do.  Measured at
// the north-star statistics shape (7 slices): rings 2+2 14.3 ms, 3+2 14.4 ms, 2+4 12.8 ms; 6 slices: 2+2 12.4, 2+5 11.2,
// 3+3 9.2 ms -- the depth of the pipeline, not the operand bandwidth, is what the 2-stage kernel of round 1 was short of.
// (Also measured: cp.async.bulk.prefetch.tensor into L2 a few K blocks ahead of the loads, to make up for the depth that
// does not fit: 14.7 / 16.9 ms against 11.1 / 12.9 ms whatever the distance -- the prefetches cost the TMA unit as many
// row requests as the loads themselves.)
template <int NS, int SA, int SB, int KBLK>
__global__ void __launch_bounds__(THREADS, 1) gemm_kernel(const __grid_constant__ CUtensorMap mapA,
                                                          const __grid_constant__ CUtensorMap mapB, const Args a) {
    constexpr int KB = KBLK;
    constexpr int A_SLICE = BM * KB, B_SLICE = BN * KB;                 // bytes per slice tile
    constexpr int A_STAGE = NS * A_SLICE, B_STAGE = NS * B_SLICE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smemB = smem + SA * A_STAGE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smemB + SB * B_STAGE);
    uint64_t *fullA = bars, *emptyA = fullA + SA, *fullB = emptyA + SA, *emptyB = fullB + SB;
    uint64_t *tmem_full = emptyB + SB, *tmem_empty = tmem_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 1);
    double *sB_s = reinterpret_cast<double *>(tmem_empty + 3);             // [2][BN] column scales of the current tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int64_t tiles_n = (a.N + BN - 1) / BN, tiles_m = (a.M + BM - 1) / BM;
    const int64_t tiles = tiles_m * tiles_n, units = tiles * a.splits;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int sta = 0, stb = 0;
            uint32_t pha = 0, phb = 0;
            for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
                const int64_t tile = u % tiles;
                const int split = int(u / tiles);
                const int m0 = int(tile / tiles_n) * BM, n0 = int(tile % tiles_n) * BN;
                const int kb0 = split * a.kb_per_split, kb1 = min(kb0 + a.kb_per_split, a.kblocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    // B first: its ring is the shallow one, so its slot is the one the MMAs are waiting for
                    mbar_wait_relaxed(&emptyB[stb], phb ^ 1);
                    mbar_expect_tx(&fullB[stb], B_STAGE);
                    uint8_t *sb = smemB + stb * B_STAGE;
#pragma unroll
                    for (int i = 0; i < NS; ++i) tma_load_3d(sb + i * B_SLICE, &mapB, &fullB[stb], kb * KB, n0, i);
                    mbar_wait_relaxed(&emptyA[sta], pha ^ 1);
                    mbar_expect_tx(&fullA[sta], A_STAGE);
                    uint8_t *sa = smem + sta * A_STAGE;
#pragma unroll
                    for (int i = 0; i < NS; ++i) tma_load_3d(sa + i * A_SLICE, &mapA, &fullA[sta], kb * KB, m0, i);
                    if (++sta == SA) { sta = 0; pha ^= 1; }
                    if (++stb == SB) { stb = 0; phb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // A_i multiplies the stacked B slices 0 .. NS-1-i (consecutive 64-row tiles in shared memory) in one or two
        // wide MMAs whose N columns land on the consecutive accumulators t = i .. NS-1.
        if (lane == 0) {
            int sta = 0, stb = 0;
            uint32_t pha = 0, phb = 0, tphase = 0;
            for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
                const int split = int(u / tiles);
                const int kb0 = split * a.kb_per_split, kb1 = min(kb0 + a.kb_per_split, a.kblocks);
                mbar_wait_relaxed(tmem_empty, tphase ^ 1);   // epilogue has drained the accumulators
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&fullB[stb], phb);
                    mbar_wait(&fullA[sta], pha);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint8_t *sa = smem + sta * A_STAGE, *sb = smemB + stb * B_STAGE;
#pragma unroll
                    for (int kk = 0; kk < KB / UMMA_K; ++kk) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) {
                            const uint64_t da = make_desc<KBLK>(sa + i * A_SLICE + kk * UMMA_K);
                            const uint32_t acc = (i != 0 || kb != kb0 || kk != 0) ? 1u : 0u;
                            constexpr int MAXJ = 256 / BN;                       // B slices per MMA (N <= 256)
#pragma unroll
                            for (int j0 = 0; j0 < NS - i; j0 += MAXJ) {
                                const int cnt = (NS - i - j0 < MAXJ) ? NS - i - j0 : MAXJ;
                                const uint64_t db = make_desc<KBLK>(sb + j0 * B_SLICE + kk * UMMA_K);
                                mma_i8(tmem_base + (i + j0) * BN, da, db, make_idesc_n(cnt * BN), acc);
                            }
                        }
                    }
                    mma_commit(&emptyA[sta]);                // frees the two slots when these MMAs retire
                    mma_commit(&emptyB[stb]);
                    if (++sta == SA) { sta = 0; pha ^= 1; }
                    if (++stb == SB) { stb = 0; phb ^= 1; }
                }
                mma_commit(tmem_full);                       // accumulators complete
                tphase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> exact int64 combination -> FP64 -> global =====
        // sum_t v_t 2^-(12+7t) = 2^-33 (hi + lo 2^-(7 (NS-4))),  hi = sum_{t<4} v_t 2^(7(3-t)) (< 2^53, exact in FP64),
        // lo = sum_{t>=4} v_t 2^(7(NS-1-t)).  Eight warps: warp 4 + q and 8 + q share TMEM lane quadrant q and take one
        // 32-column half of the tile each (two warps per scheduler hide the conversion and store latencies).
        const int q = warp & 3, half = (warp - 4) >> 2;
        uint32_t tphase = 0;
        const double w_hi = __longlong_as_double((long long)(1023 - 33) << 52);
        const double w_lo = __longlong_as_double((long long)(1023 - 7 * (NS - 4)) << 52);
        for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
            const int64_t tile = u % tiles;
            const int split = int(u / tiles);
            const int64_t m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            // column scales of this tile through shared memory (one global load per column and tile instead of one per
            // element); double buffered by tile parity, one barrier of the 256 epilogue threads per tile
            double *sBt = sB_s + (tphase ? BN : 0);
            {
                const int et = threadIdx.x - 128;
                if (et < BN) sBt[et] = (n0 + et < a.N) ? a.sB[n0 + et] : 0.0;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait_relaxed(tmem_full, tphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t row = m0 + q * 32 + lane;
            const double sa = (row < a.M) ? a.sA[row] * w_hi : 0.0;
            double *crow = a.C + split * a.split_stride + row * a.ldc + n0 + half * 32;
#pragma unroll 1
            for (int sub = 0; sub < 2; ++sub) {
                const uint32_t tcol = tmem_base + (uint32_t(q * 32) << 16) + half * 32 + sub * 16;
                long long hi[16], lo[16];
                {
                    int32_t v0[16], v1[16], v2[16], v3[16];
                    tmem_ld16(tcol + 0 * BN, v0);
                    tmem_ld16(tcol + 1 * BN, v1);
                    tmem_ld16(tcol + 2 * BN, v2);
                    tmem_ld16(tcol + 3 * BN, v3);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        hi[c] = ((long long)v0[c] << 21) + ((long long)v1[c] << 14) + ((long long)v2[c] << 7) + (long long)v3[c];
                }
                {
                    int32_t v4[16], v5[16], v6[16];
                    tmem_ld16(tcol + 4 * BN, v4);
                    tmem_ld16(tcol + 5 * BN, v5);
                    if (NS == 7) tmem_ld16(tcol + 6 * BN, v6);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        lo[c] = (NS == 7) ? ((long long)v4[c] << 14) + ((long long)v5[c] << 7) + (long long)v6[c]
                                          : ((long long)v4[c] << 7) + (long long)v5[c];
                }
                if (sub == 1) {                              // all TMEM reads of this warp's half are done
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty);
                }
                if (row < a.M) {
                    const int64_t nb = n0 + half * 32 + sub * 16;
                    double *cp = crow + sub * 16;
                    const double *sbp = sBt + half * 32 + sub * 16;
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        double o0 = fma(double(lo[c]), w_lo, double(hi[c])) * sa;
                        double o1 = fma(double(lo[c + 1]), w_lo, double(hi[c + 1])) * sa;
                        if (nb + c + 1 < a.N) {
                            o0 *= sbp[c]; o1 *= sbp[c + 1];
                            double2 *dst = reinterpret_cast<double2 *>(cp + c);
                            if (a.accumulate) { double2 old = *dst; o0 += old.x; o1 += old.y; }
                            *dst = make_double2(o0, o1);
                        } else if (nb + c < a.N) {
                            o0 *= sbp[c];
                            if (a.accumulate) o0 += cp[c];
                            cp[c] = o0;
                        }
                    }
                }
            }
            tphase ^= 1;
        }
    }
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// One warp per row: power-of-two scale and NS int8 slices (first slice 6 bits + sign, the rest 7 bits + sign,
// round to nearest so every slice is in [-64, 64]).  out: [NS][rows][Kp] int8, zero padded in k.
__global__ void slice_rows_kernel(const double *X, int64_t ldx, int64_t rows, int K, int Kp, int ns, int8_t *out,
                                  int64_t slice_stride, double *scale) {
    const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const double *x = X + row * ldx;
    double mx = 0.0;
    for (int k = lane; k < K; k += 32) mx = fmax(mx, fabs(x[k]));
    mx = warp_max(mx);
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);                   // mx = m * 2^e, m in [0.5, 1)  ->  |x| / 2^e < 1
    if (lane == 0) scale[row] = ldexp(1.0, e);
    const double s0 = ldexp(64.0, -e);
    for (int k = lane; k < Kp; k += 32) {
        double r = (k < K) ? x[k] * s0 : 0.0;
        int8_t *o = out + row * int64_t(Kp) + k;
        for (int t = 0; t < ns; ++t) {
            double v = rint(r);
            o[t * slice_stride] = (int8_t)v;
            r = (r - v) * 128.0;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// slices: ns planes of (rows, Kp) int8, rows `row_stride` bytes apart, planes `slice_stride` bytes apart
static int make_map(CUtensorMap *map, const int8_t *slices, int64_t rows, int Kp, int64_t row_stride, int64_t slice_stride,
                    int ns, int box_rows, int kblk) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return PET_ECUDA; }
    if ((row_stride & 15) || (slice_stride & 15) || (reinterpret_cast<uintptr_t>(slices) & 15)) {
        set_error("ozaki_gemm: slice planes must be 16-byte aligned");
        return PET_EINVAL;
    }
    cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)ns};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride, (cuuint64_t)slice_stride};
    cuuint32_t box[3] = {(cuuint32_t)kblk, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t *>(slices), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, kblk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", int(r)); return PET_ECUDA; }
    return PET_OK;
}

// ---- column-wise slicing with a transposing store (operands whose reduction runs over ROWS) -------------------
// max |X[r][c]| over r, as the bit pattern of a non-negative double (monotonic under integer max)
__global__ void col_absmax_kernel(const double *X, int64_t ldx, int64_t rows, int cols, const double *rowscale, int64_t rs_stride,
                                  unsigned long long *out) {
    __shared__ double red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int64_t r_per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * r_per, r1 = min(rows, r0 + r_per);
    double m = 0.0;
    if (c < cols)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
            double v = X[r * ldx + c];
            if (rowscale) v *= rowscale[r * rs_stride];
            m = fmax(m, fabs(v));
        }
    red[threadIdx.y][threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        for (int j = 1; j < 8; ++j) m = fmax(m, red[j][threadIdx.x]);
        atomicMax(out + c, (unsigned long long)__double_as_longlong(m));
    }
}

// X (rows, cols) -> out[t][c][r] int8 for r < Kp (zero for r >= rows), column scale 2^e from the column maximum.
// CTA tile: 128 rows x 32 columns, staged through shared memory so that both sides are coalesced.
constexpr int SC_R = 128, SC_C = 32, SC_PITCH = SC_R + 4;
__global__ void __launch_bounds__(256) slice_cols_kernel(const double *X, int64_t ldx, int64_t rows, int cols, int Kp,
                                                         const unsigned long long *colmax, int ns, int8_t *out,
                                                         int64_t row_stride, int64_t slice_stride, double *scale,
                                                         const double *rowscale, int64_t rs_stride) {
    extern __shared__ int8_t tile[];                  // [ns][SC_C][SC_PITCH]
    int32_t *tile32 = reinterpret_cast<int32_t *>(tile);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * SC_C + tx;
    const int64_t r0 = int64_t(blockIdx.y) * SC_R;
    double s0 = 0.0;
    if (c < cols) {
        const double mx = __longlong_as_double((long long)colmax[c]);
        int e = 0;
        if (mx > 0.0) frexp(mx, &e);
        s0 = ldexp(64.0, -e);
        if (blockIdx.y == 0 && ty == 0) scale[c] = ldexp(1.0, e);
    }
    // thread (tx, ty): column c, the 16 consecutive rows r0 + 16 ty ...; four rows pack into one shared word per slice
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int64_t r = r0 + ty * 16 + i;
        v[i] = (c < cols && r < rows) ? X[r * ldx + c] * (rowscale ? rowscale[r * rs_stride] : 1.0) * s0 : 0.0;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        for (int t = 0; t < ns; ++t) {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double q = rint(v[4 * g + i]);
                w |= (uint32_t(int(q)) & 0xFFu) << (8 * i);
                v[4 * g + i] = (v[4 * g + i] - q) * 128.0;
            }
            tile32[((t * SC_C + tx) * SC_PITCH) / 4 + ty * 4 + g] = int32_t(w);
        }
    }
    __syncthreads();
    // one warp stores one (slice, column) row of 128 bytes per step
    const int n_rows_out = ns * SC_C;
    for (int ro = ty; ro < n_rows_out; ro += 8) {
        const int t = ro / SC_C, cc = ro % SC_C;
        const int col = blockIdx.x * SC_C + cc;
        const int64_t r = r0 + tx * 4;
        if (col < cols && r < Kp)
            *reinterpret_cast<int32_t *>(out + t * slice_stride + col * row_stride + r) = tile32[((t * SC_C + cc) * SC_PITCH) / 4 + tx];
    }
}

__global__ void add_slabs_kernel(double *dst, const double *slabs, int64_t count, int n_slabs) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double v = dst[i];
    for (int s = 0; s < n_slabs; ++s) v += slabs[s * count + i];
    dst[i] = v;
}
}  // namespace oz

int ozaki_kp(int64_t K) { return int(round_up(K, oz::KB)); }

// split-K factor for an (M, N, Kp) product: whole waves of (tile, split) units, each unit paying a fixed epilogue
int ozaki_splits(int64_t M, int64_t N, int Kp, int sm_count, int max_kb_per_split) {
    const int64_t tiles = ceil_div(M, oz::BM) * ceil_div(N, oz::BN);
    const int kblocks = Kp / oz::KB;
    int best = std::max(1, int(ceil_div(kblocks, max_kb_per_split)));
    double best_cost = 1e300;
    // int32 accumulators: at most 1170 K blocks per split with |slices| <= 64 on both sides, 589 when one side holds
    // unsigned 7-bit digits (<= 127)
    const int s_min = int(ceil_div(kblocks, max_kb_per_split));
    for (int s = s_min; s <= std::max(16, 4 * s_min) && s <= kblocks; ++s) {
        const int kbs = int(ceil_div(kblocks, s));
        if (kbs < 16 && s > s_min) break;
        const double waves = double(ceil_div(tiles * ceil_div(kblocks, kbs), sm_count));
        const double cost = waves * (kbs + 3.0);
        if (cost < best_cost * 0.999) { best_cost = cost; best = s; }
    }
    return best;
}

// X (rows, cols) row-major -> transposed slices out[t][c][r] with per-column scales (colmax: cols scratch words)
// have_colmax: colmax already holds (an upper bound of) the column maxima, e.g. from the scale kernel
// rowscale (optional): every row r is multiplied by rowscale[r * rs_stride] on load (the posterior normalisation)
int ozaki_slice_cols(const double *X, int64_t ldx, int64_t rows, int cols, int ns, unsigned long long *colmax, bool have_colmax,
                     int8_t *out, int64_t row_stride, int64_t slice_stride, double *scale, cudaStream_t st,
                     const double *rowscale, int64_t rs_stride) {
    if (rows <= 0 || cols <= 0) return PET_OK;
    const int Kp = ozaki_kp(rows);
    if (!have_colmax) {
        PET_CUDA(cudaMemsetAsync(colmax, 0, size_t(cols) * 8, st));
        dim3 g1((unsigned)ceil_div(cols, 32), (unsigned)std::min<int64_t>(ceil_div(rows, 256), 64));
        oz::col_absmax_kernel<<<g1, dim3(32, 8), 0, st>>>(X, ldx, rows, cols, rowscale, rs_stride, colmax);
        PET_LAUNCH_CHECK();
    }
    dim3 g2((unsigned)ceil_div(cols, oz::SC_C), (unsigned)ceil_div(Kp, oz::SC_R));
    const size_t smem = size_t(ns) * oz::SC_C * oz::SC_PITCH;
    oz::slice_cols_kernel<<<g2, 256, smem, st>>>(X, ldx, rows, cols, Kp, colmax, ns, out, row_stride, slice_stride, scale, rowscale,
                                                 rs_stride);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

int ozaki_add_slabs(double *dst, const double *slabs, int64_t count, int n_slabs, cudaStream_t st) {
    if (n_slabs <= 0 || count <= 0) return PET_OK;
    oz::add_slabs_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(dst, slabs, count, n_slabs);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// dst (cols, ld_dst) = transpose of the sum of n_slabs slabs (rows, ld_src), slab_stride doubles apart
__global__ void add_slabs_t_kernel(double *dst, int64_t ld_dst, const double *slabs, int rows, int cols, int64_t ld_src,
                                   int64_t slab_stride, int n_slabs) {
    __shared__ double t[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        double v = 0.0;
        if (r < rows && c < cols)
            for (int s = 0; s < n_slabs; ++s) v += slabs[s * slab_stride + int64_t(r) * ld_src + c];
        t[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[int64_t(c) * ld_dst + r] = t[threadIdx.x][i];
    }
}
int ozaki_add_slabs_t(double *dst, int64_t ld_dst, const double *slabs, int rows, int cols, int64_t ld_src, int64_t slab_stride,
                      int n_slabs, cudaStream_t st) {
    if (n_slabs <= 0 || rows <= 0 || cols <= 0) return PET_OK;
    dim3 g((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32)), b(32, 8);
    add_slabs_t_kernel<<<g, b, 0, st>>>(dst, ld_dst, slabs, rows, cols, ld_src, slab_stride, n_slabs);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// X (rows, K) row-major -> slices out[t][r][k] (rows Kp bytes apart, planes slice_stride bytes apart), per-row scales
int ozaki_slice_rows(const double *X, int64_t ldx, int64_t rows, int K, int ns, int8_t *out, int64_t slice_stride, double *scale,
                     cudaStream_t st) {
    if (rows <= 0) return PET_OK;
    oz::slice_rows_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, st>>>(X, ldx, rows, K, ozaki_kp(K), ns, out, slice_stride, scale);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

// C(M,N) (+)= A . B^T from pre-sliced operands (slice rows Kp bytes apart, Kp a multiple of 64).  ns in {6, 7}.
// splits > 1: split s covers an equal share of the K blocks and writes C + s * split_stride.
int ozaki_gemm(int64_t M, int64_t N, int Kp, int ns, const OzOperand &A, const OzOperand &B, double *C, int64_t ldc, int splits,
               int64_t split_stride, bool accumulate, int sm_count, cudaStream_t st, int max_pair_product) {
    if (M <= 0 || N <= 0) return PET_OK;
    if ((ldc & 1) || (split_stride & 1) || (reinterpret_cast<uintptr_t>(C) & 15)) {
        set_error("ozaki_gemm: C must be 16-byte aligned with even ldc");
        return PET_EINVAL;
    }
    if (Kp <= 0 || Kp % oz::KB) { set_error("ozaki_gemm: padded K must be a positive multiple of %d", oz::KB); return PET_EINVAL; }
    const int kblocks = Kp / oz::KB;
    if (splits < 1) splits = 1;
    if (splits > kblocks) splits = kblocks;
    const int kbs = int(ceil_div(kblocks, splits));
    splits = int(ceil_div(kblocks, kbs));                  // no empty split
    if (int64_t(kbs) * oz::KB * ns * max_pair_product >= (int64_t(1) << 31)) {
        set_error("ozaki_gemm: %d K elements per split overflow the int32 accumulators", kbs * oz::KB);
        return PET_EINVAL;
    }
    // stage granularity: 64-byte K blocks, two stages of 86 KB (default), or 32-byte blocks, five stages of 43 KB
    // (PET_OZ_KB=32).  Measured at the north-star shape (round 2): 14.1 / 13.9 ms (score / statistics GEMM) against
    // 16.7 / 16.9 ms -- finer grains cost more TMA requests and barrier round trips than the deeper pipeline hides.
    static const int kblk = []() { const char *e = getenv("PET_OZ_KB"); return (e && atoi(e) == 32) ? 32 : 64; }();
    CUtensorMap mapA, mapB;
    PET_CHECK(oz::make_map(&mapA, A.slices, M, Kp, A.row_stride, A.slice_stride, ns, oz::BM, kblk));
    PET_CHECK(oz::make_map(&mapB, B.slices, N, Kp, B.row_stride, B.slice_stride, ns, oz::BN, kblk));
    const int per = oz::KB / kblk;                               // pipeline blocks per 64-byte K block
    oz::Args a{M, N, kblocks * per, splits, kbs * per, accumulate ? 1 : 0, A.scale, B.scale, C, ldc, split_stride};
    const int64_t units = ceil_div(M, oz::BM) * ceil_div(N, oz::BN) * splits;
    const unsigned grid = (unsigned)std::min<int64_t>(units, sm_count);
    auto launch = [&](auto kern, int sa, int sb) -> int {
        const size_t smem = size_t(ns) * (size_t(sa) * oz::BM + size_t(sb) * oz::BN) * kblk + 1024 + (2 * (sa + sb) + 4) * 8 +
                            2 * oz::BN * 8 + 64;
        PET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        kern<<<grid, oz::THREADS, smem, st>>>(mapA, mapB, a);
        return PET_OK;
    };
    // 7 slices: two A stages + four B stages (229 KB); PET_OZ_RING=22 keeps the two whole stages of round 1
    static const int ring = []() { const char *e = getenv("PET_OZ_RING"); return e ? atoi(e) : 0; }();
    const bool ring22 = ring == 22;
    if (ns == 7 && kblk == 32) PET_CHECK(launch(oz::gemm_kernel<7, 5, 5, 32>, 5, 5));
    else if (ns == 7 && ring22) PET_CHECK(launch(oz::gemm_kernel<7, 2, 2, 64>, 2, 2));
    else if (ns == 7 && ring == 24) PET_CHECK(launch(oz::gemm_kernel<7, 2, 4, 64>, 2, 4));
    else if (ns == 6 && ring22) PET_CHECK(launch(oz::gemm_kernel<6, 2, 2, 64>, 2, 2));
    else if (ns == 6 && ring == 25) PET_CHECK(launch(oz::gemm_kernel<6, 2, 5, 64>, 2, 5));
    else if (ns == 7 && ring == 32) PET_CHECK(launch(oz::gemm_kernel<7, 3, 2, 64>, 3, 2));
    else if (ns == 7) PET_CHECK(launch(oz::gemm_kernel<7, 2, 4, 64>, 2, 4));
    else if (ns == 6 && kblk == 32) PET_CHECK(launch(oz::gemm_kernel<6, 6, 6, 32>, 6, 6));
    else if (ns == 6) PET_CHECK(launch(oz::gemm_kernel<6, 3, 3, 64>, 3, 3));
    else {
        set_error("ozaki_gemm: ns must be 6 or 7");
        return PET_EINVAL;
    }
    PET_LAUNCH_CHECK();
    return PET_OK;
}

}  // namespace pet

static double g_oz_last_ms = 0.0;
/* milliseconds per product of the last pet_ozaki_gemm_* call (CUDA events around the `repeat` loop, slicing excluded) */
extern "C" double pet_ozaki_last_ms(void) { return g_oz_last_ms; }

// Test / benchmark entry: slices both operands (workspace from cudaMalloc) and multiplies.
extern "C" int pet_ozaki_gemm_kk(int64_t M, int64_t N, int64_t K, const double *A_dev, int64_t lda, const double *B_dev,
                                 int64_t ldb, double *C_dev, int64_t ldc, int32_t nslices, int32_t repeat, void *stream) {
    using namespace pet;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int Kp = ozaki_kp((int)K);
    int8_t *As = nullptr, *Bs = nullptr;
    double *sA = nullptr, *sB = nullptr;
    PET_CUDA(cudaMalloc(&As, size_t(nslices) * M * Kp));
    PET_CUDA(cudaMalloc(&Bs, size_t(nslices) * N * Kp));
    PET_CUDA(cudaMalloc(&sA, M * 8));
    PET_CUDA(cudaMalloc(&sB, N * 8));
    int rc = ozaki_slice_rows(A_dev, lda, M, (int)K, nslices, As, M * int64_t(Kp), sA, st);
    if (rc == PET_OK) rc = ozaki_slice_rows(B_dev, ldb, N, (int)K, nslices, Bs, N * int64_t(Kp), sB, st);
    const OzOperand opA{As, Kp, M * int64_t(Kp), sA}, opB{Bs, Kp, N * int64_t(Kp), sB};
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
    for (int r = 0; rc == PET_OK && r < (repeat > 0 ? repeat : 1); ++r)
        rc = ozaki_gemm(M, N, Kp, nslices, opA, opB, C_dev, ldc, 1, 0, false, sms, st);
    cudaEventRecord(ev1, st);
    cudaError_t e = cudaStreamSynchronize(st);
    float ms = 0.f;
    if (e == cudaSuccess) { cudaEventElapsedTime(&ms, ev0, ev1); g_oz_last_ms = ms / (repeat > 0 ? repeat : 1); }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    if (rc == PET_OK && e != cudaSuccess) { set_error("ozaki gemm failed: %s", cudaGetErrorString(e)); rc = PET_ECUDA; }
    cudaFree(As); cudaFree(Bs); cudaFree(sA); cudaFree(sB);
    return rc;
}

// Test / benchmark entry for operands whose reduction runs over rows: C(M,N) = A^T . B, A (K,M) lda, B (K,N) ldb.
extern "C" int pet_ozaki_gemm_mn(int64_t M, int64_t N, int64_t K, const double *A_dev, int64_t lda, const double *B_dev,
                                 int64_t ldb, double *C_dev, int64_t ldc, int32_t nslices, int32_t repeat, void *stream) {
    using namespace pet;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int Kp = ozaki_kp(K);
    const int splits = ozaki_splits(M, N, Kp, sms);
    int8_t *As = nullptr, *Bs = nullptr;
    double *sA = nullptr, *sB = nullptr, *slabs = nullptr;
    unsigned long long *cm = nullptr;
    PET_CUDA(cudaMalloc(&As, size_t(nslices) * M * Kp));
    PET_CUDA(cudaMalloc(&Bs, size_t(nslices) * N * Kp));
    PET_CUDA(cudaMalloc(&sA, M * 8));
    PET_CUDA(cudaMalloc(&sB, N * 8));
    PET_CUDA(cudaMalloc(&cm, std::max(M, N) * 8));
    PET_CUDA(cudaMalloc(&slabs, size_t(splits) * M * ldc * 8));
    int rc = ozaki_slice_cols(A_dev, lda, K, (int)M, nslices, cm, false, As, Kp, M * int64_t(Kp), sA, st);
    if (rc == PET_OK) rc = ozaki_slice_cols(B_dev, ldb, K, (int)N, nslices, cm, false, Bs, Kp, N * int64_t(Kp), sB, st);
    const OzOperand opA{As, Kp, M * int64_t(Kp), sA}, opB{Bs, Kp, N * int64_t(Kp), sB};
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
    for (int r = 0; rc == PET_OK && r < (repeat > 0 ? repeat : 1); ++r) {
        rc = ozaki_gemm(M, N, Kp, nslices, opA, opB, slabs, ldc, splits, M * ldc, false, sms, st);
        if (rc == PET_OK) {
            cudaMemcpyAsync(C_dev, slabs, size_t(M) * ldc * 8, cudaMemcpyDeviceToDevice, st);
            rc = ozaki_add_slabs(C_dev, slabs + M * ldc, M * ldc, splits - 1, st);
        }
    }
    cudaEventRecord(ev1, st);
    cudaError_t e = cudaStreamSynchronize(st);
    float ms = 0.f;
    if (e == cudaSuccess) { cudaEventElapsedTime(&ms, ev0, ev1); g_oz_last_ms = ms / (repeat > 0 ? repeat : 1); }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    if (rc == PET_OK && e != cudaSuccess) { set_error("ozaki gemm failed: %s", cudaGetErrorString(e)); rc = PET_ECUDA; }
    cudaFree(As); cudaFree(Bs); cudaFree(sA); cudaFree(sB); cudaFree(cm); cudaFree(slabs);
    return rc;
}
