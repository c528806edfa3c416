// Multi-cause state evaluation of the binary Gaussian-linear model (BSC) on the int8 tensor cores.
//
// For a datapoint with candidates c_1..c_H' every truncated state s (2..gamma active causes, bsc_et.py:180-185) has
//     F(s) - beta pre1 ||y||^2 = sum_{j in s} (lp + c lin_j) + sum_{j<k in s} 2 c G[c_j, c_k]  =  (M_f v)_s ,
//     c = beta pre1,  lin_j = G[c_j,c_j] - 2 (yW)[c_j],
// a product of the CONSTANT 0/1 membership matrix M_f (states x (H' + H'(H'-1)/2) features) with a per-datapoint
// feature vector v, and the statistics the M-step needs (bsc_et.py:349-366, 395-415) are the reverse product
//     [ <s_j> , <s_j s_k> ]  =  p^T M_f ,   p_s = exp(F(s) - max).
// Both run as tcgen05.mma kind::i8 with float64 accuracy:
//   * v is scaled per datapoint by a power of two and cut into 8 digits of 7 bits in two's complement: seven unsigned
//     digits and a SIGNED top digit (the MMAs of the top accumulator read the A operand as s8; an unsigned digit
//     <= 127 is the same byte either way), so no offset has to be added or removed; digits
//     2u and 2u+1 share accumulator u because the membership operand is stored twice, once with weight 1 and once with
//     weight 128 (u8), concatenated along K: 4 int32 accumulators instead of 8, each an EXACT integer
//   * one CTA (16 warps) owns a tile of 128 datapoints = the 128 TMEM lanes; four warps share a lane quadrant and split
//     the 64 columns of a chunk, so an epilogue thread owns ONE datapoint and 16 states per chunk: running maximum,
//     partition sum and posterior need no cross-thread traffic until the tile is finished
//   * pass 1 (top two accumulators only) bounds the maximum of F(s) from above to 2^-20 of the feature scale; pass 2
//     evaluates exp(F(s) - bound) in float64 (batches of 8 states x 32 datapoints that are all below e^-45 of their
//     maximum are skipped by one warp-uniform test), cuts the posterior into 6 digits of 7 bits (42 bits) and writes
//     them as the A operand of the reverse product, whose 3 accumulators (78 features x 128 datapoints) stay in TMEM
//     for the whole tile
//   * GLF_DEFER_STATS (first sweep of a truncated iteration): the pair sums and scalar contributions are parked per
//     datapoint instead of being added to Wq / scalars; gl_finalize_cut adds up the datapoints that survive the cut
//   * operands are written by the threads in the un-swizzled K-major core-matrix layout (row r of K chunk kc at
//     kc * rows * 16 + r * 16: consecutive lanes store consecutive 16-byte units, no bank conflicts); the constant
//     membership tables are kept in global memory as shared-memory images and fetched per 64-state chunk with
//     cp.async.bulk
// Outputs are those of gl_state_kernel (gl_kernel.cu): lse, the normalisation record scl, candidate marginals folded
// into the <s> row, second moments scattered into Wq, the scalar statistics.
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "gl_kernel.cuh"
#include "gl_state_tc.cuh"

namespace pet {
namespace tc {

constexpr int TM = 128;                 // datapoints per tile (TMEM lanes)
constexpr int NC = TC_NC;               // states per chunk (MMA N of the forward product)
constexpr int KF = 96;                  // features, padded (K of one digit plane)
constexpr int KFC = KF / 16;            // 16-byte K chunks per digit plane
constexpr int NDF = 8;                  // digits of the features (56 bits)
constexpr int NDP = 6;                  // digits of the posterior (42 bits)
constexpr int NOUT = TC_NOUT;           // outputs of the reverse product, padded (MMA N)
constexpr int NPART = 4;                // warps per TMEM lane quadrant: each owns 16 of the 64 columns of a chunk
constexpr int THREADS = 128 * NPART;
constexpr int CSTR = 13;                // stride of the candidate rows in shared memory (conflict-free)
constexpr double EXP_CUTOFF = -100.0;   // as gl_kernel.cu
constexpr double SKIP_CUTOFF = -45.0;   // a batch whose posteriors all lie below e^-45 of the largest one is skipped
constexpr int XB = 7 * NDF - 1;         // the features are scaled to integers of magnitude below 2^XB (the top digit lies in [-64, 63])

constexpr int A_FWD_BYTES = NDF * KFC * TM * 16;            // 98304
constexpr int B_FWD_BYTES = TC_BFWD_BYTES;                  // 2 * KF * NC = 12288 per chunk
constexpr int A_REV_BYTES = NDP * (NC / 16) * TM * 16;      // 49152
constexpr int B_REV_BYTES = TC_BREV_BYTES;                  // 2 * NC * NOUT = 10240 per chunk
constexpr int OFF_A_FWD = 0;
constexpr int OFF_B_FWD = OFF_A_FWD + A_FWD_BYTES;
constexpr int OFF_A_REV = OFF_B_FWD + 2 * B_FWD_BYTES;
constexpr int OFF_B_REV = OFF_A_REV + A_REV_BYTES;
constexpr int OFF_CAND = OFF_B_REV + 2 * B_REV_BYTES;       // int [TM][CSTR]
constexpr int OFF_DBL = OFF_CAND + TM * CSTR * 4;           // double arrays, see the kernel
constexpr int N_DBL = (NPART + 2 + 2 * NPART + 3) * TM + 32; // rowmax[NPART], scale, bias, part[2][NPART], fin[3], exp table
constexpr int OFF_INT = OFF_DBL + N_DBL * 8;                // int imax[NPART][TM]
constexpr int OFF_FEAT = OFF_INT + NPART * TM * 4;          // uint8 fj[KF], fk[KF]
constexpr int OFF_BAR = OFF_FEAT + 2 * KF;                  // 3 mbarriers + tmem slot
constexpr int SMEM_BYTES = OFF_BAR + 64;
static_assert(OFF_BAR % 8 == 0 && OFF_DBL % 8 == 0, "alignment");
static_assert(SMEM_BYTES + 128 <= 227 * 1024, "shared memory");

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Every wait has a watchdog: a protocol error surfaces as a trapped kernel with a message, never as a hung GPU.
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __noinline__ void mbar_timeout(int which, uint32_t parity) {
    printf("gl_state_tc_kernel: wait on barrier %d (parity %u) timed out in block %d thread %d\n", which, parity, blockIdx.x, threadIdx.x);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int which) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 22)) mbar_timeout(which, parity);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// un-swizzled K-major matrix descriptor: LBO = byte distance of core matrices adjacent in K, SBO = of 8-row groups
// (checked on the hardware by tools/ubench/umma_probe.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t((saddr & 0x3FFFFu) >> 4)) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) | (1ull << 46);
}
// kind::i8, D = s32, B = unsigned 8 bit, A = unsigned or (a_signed) signed 8 bit, both K-major, M = 128
__device__ __forceinline__ constexpr uint32_t make_idesc(int n, bool a_signed = false) {
    return (2u << 4) | (a_signed ? (1u << 7) : 0u) | (uint32_t(n >> 3) << 17) | (uint32_t(TM >> 4) << 24);
}
__device__ __forceinline__ void mma_u8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 7-bit digit t of the integer whose low / high words are lo / hi, for 16 integers -> 16 bytes.  TOP: the highest
// digit keeps 8 bits (a posterior of exactly 1 is 2^42: digit 5 = 128, representable because the operand is u8)
template <int T, bool TOP>
__device__ __forceinline__ uint4 pack_digit16(const uint32_t (&lo)[16], const uint32_t (&hi)[16]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int s = 4 * j + i;
            uint32_t d;
            if constexpr (7 * T + 7 <= 32) d = lo[s] >> (7 * T);
            else if constexpr (7 * T >= 32) d = hi[s] >> (7 * T - 32);
            else d = __funnelshift_r(lo[s], hi[s], 7 * T);
            d &= TOP ? 0xFFu : 0x7Fu;
            acc |= d << (8 * i);
        }
        w[j] = acc;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

struct Smem {
    uint8_t *base;
    __device__ __forceinline__ uint8_t *a_fwd() const { return base + OFF_A_FWD; }
    __device__ __forceinline__ uint8_t *b_fwd(int b) const { return base + OFF_B_FWD + b * B_FWD_BYTES; }
    __device__ __forceinline__ uint8_t *a_rev() const { return base + OFF_A_REV; }
    __device__ __forceinline__ uint8_t *b_rev(int b) const { return base + OFF_B_REV + b * B_REV_BYTES; }
    __device__ __forceinline__ int *cand() const { return reinterpret_cast<int *>(base + OFF_CAND); }
    __device__ __forceinline__ double *dbl() const { return reinterpret_cast<double *>(base + OFF_DBL); }
    __device__ __forceinline__ int *imax() const { return reinterpret_cast<int *>(base + OFF_INT); }
    __device__ __forceinline__ uint8_t *feat() const { return base + OFF_FEAT; }
    __device__ __forceinline__ uint64_t *bars() const { return reinterpret_cast<uint64_t *>(base + OFF_BAR); }
};

// forward product of one chunk: accumulators U0 .. 3 (pass 1: the top two only), table buffer b
__device__ __forceinline__ void issue_fwd(const Smem &sm, uint32_t tmem, int b, int u0) {
    const uint32_t a0 = smem_u32(sm.a_fwd()), b0 = smem_u32(sm.b_fwd(b));
    for (int u = u0; u < 4; ++u) {
        const int nk = (2 * u + 1 < NDF) ? 2 * KF / 32 : KF / 32;          // a lone top digit uses the weight-1 half only
        for (int kk = 0; kk < nk; ++kk) {
            const uint64_t da = make_desc(a0 + (2 * u * KFC + 2 * kk) * (TM * 16), TM * 16, 128);
            const uint64_t db = make_desc(b0 + (2 * kk) * (NC * 16), NC * 16, 128);
            mma_u8(tmem + u * NC, da, db, make_idesc(NC, 2 * u + 2 == NDF), kk > 0 ? 1u : 0u);    // top accumulator: signed top digit
        }
    }
}
// reverse product of one chunk into the three tile-long accumulators
__device__ __forceinline__ void issue_rev(const Smem &sm, uint32_t tmem, int b, bool first) {
    const uint32_t a0 = smem_u32(sm.a_rev()), b0 = smem_u32(sm.b_rev(b));
    for (int v = 0; v < NDP / 2; ++v)
        for (int kk = 0; kk < 2 * NC / 32; ++kk) {
            const uint64_t da = make_desc(a0 + (2 * v * (NC / 16) + 2 * kk) * (TM * 16), TM * 16, 128);
            const uint64_t db = make_desc(b0 + (2 * kk) * (NOUT * 16), NOUT * 16, 128);
            mma_u8(tmem + 4 * NC + v * NOUT, da, db, make_idesc(NOUT), (first && kk == 0) ? 0u : 1u);
        }
}

// STATS = false: log-denominators only (first sweep of a truncated iteration)
template <bool STATS>
__global__ void __launch_bounds__(THREADS, 1) gl_state_tc_kernel(const __grid_constant__ GLArgs a, const __grid_constant__ GLTc t) {
    extern __shared__ uint8_t smem_raw[];
    Smem sm{reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127))};
    const GLStatic &st = a.st;
    const GLIter &it = a.it;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, part = warp >> 2;             // TMEM lane quadrant; which 16 of the 64 columns of a chunk
    const int r = q * 32 + lane;                          // datapoint of this thread within the tile
    const int Hp = st.Hp, nf = t.n_feat;
    uint64_t *bar_tab = sm.bars();                        // [2] table buffers landed
    uint64_t *bar_mma = sm.bars() + 2;                    // MMAs issued so far have completed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm.bars() + 3);

    double *rowmax_s = sm.dbl();                          // [NPART][TM]
    double *scale_s = rowmax_s + NPART * TM;              // [TM]
    double *bias_s = scale_s + TM;                        // [TM]
    double *part_s = bias_s + TM;                         // [2 (Z, SF)][NPART][TM]
    double *fin_s = part_s + 2 * NPART * TM;              // [3 (1/Z or 0, singleton scale, keep)][TM]
    double *exptab = fin_s + 3 * TM;                      // [32] 2^(j/32)
    int *imax_s = sm.imax();                              // [NPART][TM]
    int *cand_s = sm.cand() + r * CSTR;

    if (tid == 0) {
        mbar_init(&bar_tab[0], 1);
        mbar_init(&bar_tab[1], 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 2 * KF; i += THREADS) sm.feat()[i] = t.feat[i];
    exp_tab32_init(exptab);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + (uint32_t(q * 32) << 16);

    // c = beta pre1 and the per-member log-prior as combine() of gl_kernel.cu applies them
    const double cq = it.beta * it.pre1;
    const double lpm = it.anneal_prior ? it.beta * it.lp[0] : it.lp[0];
    const bool fold = (a.flags & GLF_FOLD_SCALE) != 0;
    const bool use_cut = STATS && (a.flags & GLF_USE_CUT);
    const bool defer = STATS && (a.flags & GLF_DEFER_STATS);      // park the statistics per datapoint (see GLF_DEFER_STATS)
    const double cut = use_cut ? *a.cut : 0.0;
    // K chunks of the feature operand built by this thread: parts 0, 1 two each, parts 2, 3 one each
    const int kc0 = (part < 2) ? 2 * part : part + 2, nkc = (part < 2) ? 2 : 1;

    double acc_n = 0.0, acc_lse = 0.0, acc_sig = 0.0, acc_cnt = 0.0;
    uint32_t item = 0;            // table fetches issued so far (thread 0), = items consumed by everybody
    uint32_t mma_phase = 0;       // commits waited for so far (all threads)

    // tiles are handed out dynamically (the first one is the CTA's own index): the time of a tile depends on how many of
    // its posteriors are live, and with a handful of tiles per CTA a static assignment leaves the slowest CTA 10-15 %
    // behind the average
    const int64_t n_tiles = (a.n_rows + TM - 1) / TM;
    int *next_tile_s = reinterpret_cast<int *>(sm.bars() + 7);
    for (int64_t tile = blockIdx.x; tile < n_tiles;) {
        const int64_t rr = tile * TM + r;                 // row within the chunk of datapoints
        const bool valid = rr < a.n_rows;
        const int64_t n = a.row0 + rr;                    // global datapoint index

        // ---- table fetch for the first item of this tile (overlaps the feature build) ----
        if (tid == 0) {
            const int b = item & 1;
            mbar_expect_tx(&bar_tab[b], B_FWD_BYTES);
            bulk_g2s(sm.b_fwd(b), t.bfwd, B_FWD_BYTES, &bar_tab[b]);
        }

        // ---- features: gather, scale, 8 digits -> A operand of the forward product ----
        if (part == 0) {
#pragma unroll 4
            for (int j = 0; j < Hp; ++j) cand_s[j] = valid ? a.cand[n * Hp + j] : 0;
        }
        __syncthreads();
        double v[32];
        double vmax = 0.0;
        {
            // straight-line gathers (a branch per feature would serialise 32 L2 round trips): out-of-range features and
            // rows read a safe address and are zeroed by a select
            const uint8_t *fj = sm.feat(), *fk = sm.feat() + KF;
            const int64_t nn = valid ? n : a.row0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int f = kc0 * 16 + i;
                const bool on = valid && f < nf && i < nkc * 16;
                const int fs = on ? f : 0;
                const int cj = cand_s[fj[fs]], ck = cand_s[fk[fs]];
                const double g = a.G[int64_t(cj) * st.ldH + ck];
                const double yw = a.ywc[nn * Hp + (fs < Hp ? fs : 0)];
                double x = (fs < Hp) ? fma(cq, fma(-2.0, yw, g), lpm) : 2.0 * cq * g;
                x = on ? x : 0.0;
                v[i] = x;
                vmax = fmax(vmax, fabs(x));
            }
        }
        rowmax_s[part * TM + r] = vmax;
        __syncthreads();
        {
            double m = rowmax_s[r];
#pragma unroll
            for (int pp = 1; pp < NPART; ++pp) m = fmax(m, rowmax_s[pp * TM + r]);
            int e = 0;
            if (m > 0.0 && m < INFINITY) frexp(m, &e);                   // m = f 2^e, f in [0.5, 1): |v| 2^-e < 1
            const double up = ldexp(1.0, XB - e);                        // x = rint(v 2^(XB-e)) in (-2^XB, 2^XB)
            if (part == 0) scale_s[r] = ldexp(1.0, e - XB);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                if (kk < nkc) {
                    uint32_t lo[16], hi[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int f = (kc0 + kk) * 16 + i;
                        long long xi = __double2ll_rn(v[kk * 16 + i] * up);      // (padding features are 0)
                        xi = max(min(xi, (1ll << XB) - 1), 1 - (1ll << XB));     // (NaN / inf inputs: keep the digits in range)
                        lo[i] = uint32_t(xi);
                        hi[i] = uint32_t(xi >> 32);
                    }
                    uint8_t *dst = sm.a_fwd() + (kc0 + kk) * (TM * 16) + r * 16;
                    *reinterpret_cast<uint4 *>(dst + 0 * KFC * TM * 16) = pack_digit16<0, false>(lo, hi);
                    *reinterpret_cast<uint4 *>(dst + 1 * KFC * TM * 16) = pack_digit16<1, false>(lo, hi);
                    *reinterpret_cast<uint4 *>(dst + 2 * KFC * TM * 16) = pack_digit16<2, false>(lo, hi);
                    *reinterpret_cast<uint4 *>(dst + 3 * KFC * TM * 16) = pack_digit16<3, false>(lo, hi);
                    *reinterpret_cast<uint4 *>(dst + 4 * KFC * TM * 16) = pack_digit16<4, false>(lo, hi);
                    *reinterpret_cast<uint4 *>(dst + 5 * KFC * TM * 16) = pack_digit16<5, false>(lo, hi);
                    *reinterpret_cast<uint4 *>(dst + 6 * KFC * TM * 16) = pack_digit16<6, false>(lo, hi);
                    static_assert(NDF == 8, "eight digit planes, the last one signed");
                    *reinterpret_cast<uint4 *>(dst + 7 * KFC * TM * 16) = pack_digit16<7, true>(lo, hi);    // bits 49..56: s8
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        // ---- pass 1: upper bound of max_s F(s) from the two leading accumulators ----
        int imax = INT_MIN;
        for (int c = 0; c < t.n_chunks; ++c) {
            if (tid == 0) {
                const int b = item & 1;
                mbar_wait(&bar_tab[b], (item >> 1) & 1, b);
                tc_fence_after();
                issue_fwd(sm, tmem, b, 2);
                mma_commit(bar_mma);
            }
            mbar_wait(bar_mma, mma_phase & 1, 2);
            ++mma_phase;
            tc_fence_after();
            if (tid == 0) {                                   // next item: chunk c + 1 of pass 1, or chunk 0 of pass 2
                const int b = (item + 1) & 1;
                const int cn = (c + 1 < t.n_chunks) ? c + 1 : 0;
                const bool rev_too = STATS && (c + 1 == t.n_chunks);
                mbar_expect_tx(&bar_tab[b], B_FWD_BYTES + (rev_too ? B_REV_BYTES : 0));
                bulk_g2s(sm.b_fwd(b), t.bfwd + size_t(cn) * B_FWD_BYTES, B_FWD_BYTES, &bar_tab[b]);
                if (rev_too) bulk_g2s(sm.b_rev(b), t.brev + size_t(cn) * B_REV_BYTES, B_REV_BYTES, &bar_tab[b]);
            }
            ++item;
            __syncwarp();                                     // the TMEM loads below are warp-collective
            const int cnt = t.chunk_cnt[c];
            {
                uint32_t a2[16], a3[16];
                tmem_ld16(tlane + 2 * NC + part * 16, a2);
                tmem_ld16(tlane + 3 * NC + part * 16, a3);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int v = int(a3[i] * 16384u + a2[i]);
                    imax = max(imax, (part * 16 + i < cnt) ? v : INT_MIN);
                }
            }
            tc_fence_before();
            __syncthreads();
        }
        imax_s[part * TM + r] = imax;
        __syncthreads();
        if (part == 0) {
            int im = imax_s[r];
#pragma unroll
            for (int pp = 1; pp < NPART; ++pp) im = max(im, imax_s[pp * TM + r]);
            const double yy = valid ? a.yy[n] : 0.0;
            const double m1 = valid ? a.rs[n * (4 + PET_MAXV)] : 0.0;
            // F(s) - c yy = scale (hi 2^28 + lo),  0 <= lo < n_feat 2^28
            const double m2 = fma(scale_s[r] * 268435456.0, double(im) + double(t.max_nfeat), cq * yy);
            const double mx = fmax(m1, m2);
            bias_s[r] = cq * yy - mx;
        }
        __syncthreads();
        const double scale = scale_s[r], bias = bias_s[r];

        // ---- pass 2: posterior, partition sum, digits of the posterior -> reverse product ----
        double Z2 = 0.0, SF = 0.0;
        for (int c = 0; c < t.n_chunks; ++c) {
            if (tid == 0) {
                const int b = item & 1;
                mbar_wait(&bar_tab[b], (item >> 1) & 1, b);
                tc_fence_after();
                if (STATS && c > 0) issue_rev(sm, tmem, b ^ 1, c == 1);
                issue_fwd(sm, tmem, b, 0);
                mma_commit(bar_mma);
            }
            mbar_wait(bar_mma, mma_phase & 1, 2);
            ++mma_phase;
            tc_fence_after();
            if (tid == 0 && c + 1 < t.n_chunks) {
                const int b = (item + 1) & 1;
                mbar_expect_tx(&bar_tab[b], B_FWD_BYTES + (STATS ? B_REV_BYTES : 0));
                bulk_g2s(sm.b_fwd(b), t.bfwd + size_t(c + 1) * B_FWD_BYTES, B_FWD_BYTES, &bar_tab[b]);
                if (STATS) bulk_g2s(sm.b_rev(b), t.brev + size_t(c + 1) * B_REV_BYTES, B_REV_BYTES, &bar_tab[b]);
            }
            ++item;
            __syncwarp();
            const int cnt = t.chunk_cnt[c];
            uint32_t ylo[16], yhi[16];
            bool live = false;                                // warp-uniform: some posterior of the 16 states is not negligible
#pragma unroll
            for (int sb = 0; sb < 2; ++sb) {
                const int col0 = part * 16 + sb * 8;
                uint32_t a0[8], a1[8], a2[8], a3[8];
                tmem_ld8(tlane + 0 * NC + col0, a0);
                tmem_ld8(tlane + 1 * NC + col0, a1);
                tmem_ld8(tlane + 2 * NC + col0, a2);
                tmem_ld8(tlane + 3 * NC + col0, a3);
                tmem_wait_ld();
                double x8[8];
                double xm = -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int hi = int(a3[i] * 16384u + a2[i]);
                    const uint32_t lo = a1[i] * 16384u + a0[i];
                    const double f = fma(double(hi), 268435456.0, double(lo));
                    x8[i] = (col0 + i < cnt) ? fma(f, scale, bias) : -INFINITY;      // padding columns of a partial chunk
                    xm = fmax(xm, x8[i]);
                }
                // one warp-uniform decision per batch of 8 states x 32 datapoints: either nobody is within e^-45 of its
                // largest posterior (2.9e-20: below the last bit of the partition sum even when all 1573 states add up),
                // or the eight exps run as straight-line code so that their dependency chains interleave
                if (__any_sync(0xffffffffu, xm > SKIP_CUTOFF)) {
                    live = true;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const double p = exp_tab32(fmax(x8[i], EXP_CUTOFF), exptab);     // (x = -inf: e^-100 = 4e-44, harmless)
                        Z2 += p;
                        SF = fma(p, fmax(x8[i], EXP_CUTOFF), SF);
                        if (STATS) {
                            const double T = fma(p, 4398046511104.0, 4503599627370496.0);    // p 2^42 + 2^52: mantissa = rint(p 2^42)
                            ylo[sb * 8 + i] = uint32_t(__double2loint(T));
                            yhi[sb * 8 + i] = uint32_t(__double2hiint(T)) & 0xFFFFFu;
                        }
                    }
                } else if (STATS) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { ylo[sb * 8 + i] = 0u; yhi[sb * 8 + i] = 0u; }
                }
            }
            if (STATS) {
                uint8_t *dst = sm.a_rev() + part * (TM * 16) + r * 16;
                constexpr int PL = (NC / 16) * TM * 16;
                if (live) {
                    *reinterpret_cast<uint4 *>(dst + 0 * PL) = pack_digit16<0, false>(ylo, yhi);
                    *reinterpret_cast<uint4 *>(dst + 1 * PL) = pack_digit16<1, false>(ylo, yhi);
                    *reinterpret_cast<uint4 *>(dst + 2 * PL) = pack_digit16<2, false>(ylo, yhi);
                    *reinterpret_cast<uint4 *>(dst + 3 * PL) = pack_digit16<3, false>(ylo, yhi);
                    *reinterpret_cast<uint4 *>(dst + 4 * PL) = pack_digit16<4, false>(ylo, yhi);
                    *reinterpret_cast<uint4 *>(dst + 5 * PL) = pack_digit16<5, true>(ylo, yhi);
                } else {
#pragma unroll
                    for (int pl = 0; pl < NDP; ++pl) *reinterpret_cast<uint4 *>(dst + pl * PL) = make_uint4(0u, 0u, 0u, 0u);
                }
                fence_async_smem();
            }
            tc_fence_before();
            __syncthreads();
        }
        if (STATS) {
            if (tid == 0) {
                tc_fence_after();
                issue_rev(sm, tmem, (item & 1) ^ 1, t.n_chunks == 1);
                mma_commit(bar_mma);
            }
            mbar_wait(bar_mma, mma_phase & 1, 2);
            ++mma_phase;
            tc_fence_after();
        }
        part_s[(0 * NPART + part) * TM + r] = Z2;
        part_s[(1 * NPART + part) * TM + r] = SF;
        __syncthreads();

        // ---- per-datapoint results.  Part 0 owns the scalars of its datapoint; the four parts then share the 80 columns
        // of the reverse accumulators (the TMEM loads are warp-collective: every lane walks them, side effects are
        // predicated) ----
        double m1 = 0.0, sig1 = 0.0, cnt1 = 0.0, mx = 0.0, e1 = 0.0, lse = 0.0;
        if (part == 0) {
            bool keep = valid;
            if (valid && use_cut) {
                const double l = a.lse[n];
                keep = (a.flags & GLF_CUT_STRICT) ? (l > cut) : (l >= cut);
            }
            double *scl = a.scl + n * (1 + PET_MAXHP);
            if (valid && !keep)                               // truncated away: contributes nothing (bsc_et.py:254-257)
                for (int j = 0; j <= Hp; ++j) scl[j] = 0.0;
            double inv = 0.0, sce = 1.0;
            if (keep) {
                const double *rs = a.rs + n * (4 + PET_MAXV);
                m1 = rs[0]; sig1 = rs[2]; cnt1 = rs[4];
                mx = cq * a.yy[n] - bias;
                Z2 = 0.0; SF = 0.0;
#pragma unroll
                for (int pp = 0; pp < NPART; ++pp) { Z2 += part_s[pp * TM + r]; SF += part_s[(NPART + pp) * TM + r]; }
                e1 = (m1 == -INFINITY) ? 0.0 : exp(m1 - mx);
                const double Z = fma(rs[1], e1, Z2);
                lse = mx + log(Z);
                a.lse[n] = lse;
                if (STATS) {
                    inv = 1.0 / Z;
                    sce = e1 * inv;
                    if (fold && sce == 0.0) {      // the singletons vanish next to the multi-cause states: zero row, unit scale
                        double *Srow = a.S + rr * st.ldH;
                        for (int h = 0; h < st.ldH; ++h) Srow[h] = 0.0;
                        sce = 1.0;
                    }
                    scl[0] = sce;
                }
            }
            fin_s[r] = inv; fin_s[TM + r] = sce; fin_s[2 * TM + r] = keep ? 1.0 : 0.0;
        }
        if (STATS) {
            __syncthreads();
            const double inv = fin_s[r], sce = fin_s[TM + r];
            const bool keep = fin_s[2 * TM + r] != 0.0;
            double *scl = a.scl + n * (1 + PET_MAXHP);
            double *Srow = a.S + rr * st.ldH;
            const uint8_t *fj = sm.feat(), *fk = sm.feat() + KF;
            double sum_marg = 0.0;
            // reverse accumulators: value = (a2 2^28 + a1 2^14 + a0) 2^-42, columns = features; part p reads columns
            // 16 p .. 16 p + 15, part 0 also 64 .. 79
#pragma unroll 1
            for (int c0 = 16 * part; c0 < NOUT; c0 += 64) {
                uint32_t r0[16], r1[16], r2[16];
                tmem_ld16(tlane + 4 * NC + 0 * NOUT + c0, r0);
                tmem_ld16(tlane + 4 * NC + 1 * NOUT + c0, r1);
                tmem_ld16(tlane + 4 * NC + 2 * NOUT + c0, r2);
                tmem_wait_ld();
                if (keep) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int f = c0 + i;
                        if (f >= nf) continue;
                        const long long R = ((long long)r2[i] << 28) + ((long long)r1[i] << 14) + (long long)r0[i];
                        const double val = double(R) * 2.2737367544323206e-13;            // 2^-42
                        if (f < Hp) {                                                     // (Hp <= 12: all in part 0's first batch)
                            sum_marg += val;
                            const double mj = val * inv;
                            if (fold) {
                                if (mj != 0.0) Srow[cand_s[f]] += mj / sce;
                                scl[1 + f] = 0.0;
                            } else {
                                scl[1 + f] = mj;
                            }
                        } else {
                            const double w = val * inv;
                            if (defer) {
                                a.pairs[int64_t(f - Hp) * a.pairs_ld + n] = w;
                            } else if (w != 0.0) {
                                const int cj = cand_s[fj[f]], ck = cand_s[fk[f]];
                                atomicAdd(&a.Wq[int64_t(cj) * st.ldH + ck], w);
                                atomicAdd(&a.Wq[int64_t(ck) * st.ldH + cj], w);
                            }
                        }
                    }
                }
                if (part != 0) break;                         // parts 1..3: one batch; part 0: columns 0..15 and 64..79
            }
            if (part == 0 && keep) {
                // sum_s p_s q_s from sum_s p_s (F_s - mx):  F_s = c q_s + lpm |s|,  sum_s p_s |s| = sum_j marginal_j
                const double sig2 = (fma(mx, Z2, SF) - lpm * sum_marg) / cq;
                if (defer) {
                    double *rs = a.rs + n * (4 + PET_MAXV);
                    rs[5] = fma(sig1, e1, sig2) * inv;
                    rs[6] = fma(cnt1, e1, sum_marg) * inv;
                } else {
                    acc_n += 1.0;
                    acc_lse += lse;
                    acc_sig += fma(sig1, e1, sig2) * inv;
                    acc_cnt += fma(cnt1, e1, sum_marg) * inv;
                }
            }
        }
        if (tid == 0) *next_tile_s = int(gridDim.x) + atomicAdd(a.tile_counter, 1);
        tc_fence_before();
        __syncthreads();            // TMEM, the candidate rows and the per-datapoint arrays are reused by the next tile
        tc_fence_after();
        tile = *next_tile_s;
    }

    if (STATS && !defer) {
        acc_n = warp_sum(acc_n); acc_lse = warp_sum(acc_lse); acc_sig = warp_sum(acc_sig); acc_cnt = warp_sum(acc_cnt);
        if (lane == 0 && part == 0) {
            atomicAdd(&a.scalars[0], acc_n);
            atomicAdd(&a.scalars[1], acc_lse);
            atomicAdd(&a.scalars[2], acc_sig);
            atomicAdd(&a.scalars[3], acc_cnt);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace tc

// ---- host side: membership tables as shared-memory images ------------------------------------------------------
bool gl_tc_supported(const GLStatic &st, int gamma, bool binary) {
    const int nf = st.Hp + st.Hp * (st.Hp - 1) / 2;
    return binary && st.S >= 1 && gamma >= 2 && gamma <= 5 && nf + 1 <= tc::NOUT && nf <= tc::KF && st.n_blocks == 1 && st.has_null;
}

// matrix: S x Hp latent values (0/1), states ordered by size (size_start as in GLStatic)
int gl_tc_build_tables(const GLStatic &st, int gamma, const std::vector<double> &matrix, GLTcHost &out) {
    const int Hp = st.Hp, nf = Hp + Hp * (Hp - 1) / 2;
    out = GLTcHost();
    GLTc &t = out.dev;
    memset(&t, 0, sizeof(t));
    t.n_feat = nf;
    for (int j = 0; j < Hp; ++j) { t.feat[j] = uint8_t(j); t.feat[tc::KF + j] = uint8_t(j); }
    {
        int f = Hp;
        for (int j = 0; j < Hp; ++j)
            for (int k = j + 1; k < Hp; ++k, ++f) { t.feat[f] = uint8_t(j); t.feat[tc::KF + f] = uint8_t(k); }
    }
    // chunks of 64 consecutive multi-cause states, across the size groups (1573 states: 25 chunks)
    struct Chunk { int first, cnt; };
    std::vector<Chunk> chunks;
    for (int s = st.size_start[2]; s < st.S; s += tc::NC) chunks.push_back({s, std::min(tc::NC, st.S - s)});
    if (chunks.empty() || (int)chunks.size() > TC_MAX_CHUNKS) { set_error("tensor-core state kernel: %zu chunks unsupported", chunks.size()); return PET_EINVAL; }
    t.n_chunks = (int)chunks.size();
    out.bfwd.assign(size_t(t.n_chunks) * TC_BFWD_BYTES, 0);
    out.brev.assign(size_t(t.n_chunks) * TC_BREV_BYTES, 0);
    for (int c = 0; c < t.n_chunks; ++c) {
        const Chunk &ch = chunks[c];
        t.chunk_cnt[c] = uint8_t(ch.cnt);
        uint8_t *bf = out.bfwd.data() + size_t(c) * TC_BFWD_BYTES, *br = out.brev.data() + size_t(c) * TC_BREV_BYTES;
        for (int i = 0; i < ch.cnt; ++i) {
            const double *row = matrix.data() + size_t(ch.first + i) * Hp;
            std::vector<int> member(nf, 0);
            int cnt_members = 0;
            for (int j = 0; j < Hp; ++j) if (row[j] != 0.0) { member[j] = 1; ++cnt_members; }
            int f = Hp;
            for (int j = 0; j < Hp; ++j)
                for (int k = j + 1; k < Hp; ++k, ++f) member[f] = (row[j] != 0.0 && row[k] != 0.0) ? 1 : 0;
            if (cnt_members < 2 || cnt_members > gamma) { set_error("tensor-core state kernel: unexpected state size %d", cnt_members); return PET_EINVAL; }
            const int nfeat = cnt_members + cnt_members * (cnt_members - 1) / 2;
            t.max_nfeat = std::max(t.max_nfeat, nfeat);
            for (int ft = 0; ft < nf; ++ft) {
                if (!member[ft]) continue;
                // forward operand: row = state i, K = [feature (weight 1) | feature (weight 128)], K chunk major
                for (int half = 0; half < 2; ++half) {
                    const int k = half * tc::KF + ft;
                    bf[size_t(k >> 4) * tc::NC * 16 + i * 16 + (k & 15)] = half ? 128 : 1;
                }
                // reverse operand: row = feature, K = [state (weight 1) | state (weight 128)]
                for (int half = 0; half < 2; ++half) {
                    const int k = half * tc::NC + i;
                    br[size_t(k >> 4) * tc::NOUT * 16 + ft * 16 + (k & 15)] = half ? 128 : 1;
                }
            }
        }
    }
    return PET_OK;
}

int launch_gl_state_tc(const GLArgs &a, const GLTc &t, int sm_count, cudaStream_t stream) {
    if (a.n_rows <= 0 || (a.flags & GLF_SELECT_ONLY)) return PET_OK;
    const bool stats = !(a.flags & GLF_LSE_ONLY);
    static bool configured = false;
    if (!configured) {
        PET_CUDA(cudaFuncSetAttribute(tc::gl_state_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES + 128));
        PET_CUDA(cudaFuncSetAttribute(tc::gl_state_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES + 128));
        configured = true;
    }
    const int64_t tiles = ceil_div(a.n_rows, tc::TM);
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, sm_count);
    if (!a.tile_counter) { set_error("gl_state_tc: no tile counter"); return PET_EINVAL; }
    PET_CUDA(cudaMemsetAsync(a.tile_counter, 0, sizeof(int), stream));
    if (stats) tc::gl_state_tc_kernel<true><<<grid, tc::THREADS, tc::SMEM_BYTES + 128, stream>>>(a, t);
    else tc::gl_state_tc_kernel<false><<<grid, tc::THREADS, tc::SMEM_BYTES + 128, stream>>>(a, t);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

}  // namespace pet
