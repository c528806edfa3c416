// First-contact probe for the tcgen05 building blocks the state-space kernel relies on (tools only, not product).
//   1. kind::i8 MMA with THREAD-WRITTEN, un-swizzled K-major operands (8 x 16 B core matrices): which of the
//      descriptor's two strides is the K direction, which the M/N direction; start-address advance per K step
//   2. operand formats: s8 x s8, s8 x u8 (weight 128), u8 x s8
//   3. MMA shapes M=128 with N = 32 / 80 / 96 / 256, tcgen05.ld 32x32b.x16 / .x32 register mapping
//   4. cp.async.bulk (1-D, no tensor map) global -> shared with an mbarrier
//   5. the int8 tensor-pipe peak of the chip: all SMs issuing back-to-back M=128 N=256 K=32 MMAs on resident tiles
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// un-swizzled K-major descriptor: start >> 4, "leading" and "stride" byte offsets >> 4, version 1, layout type 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t((saddr & 0x3FFFFu) >> 4)) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_signed, int b_signed) {
    return (2u << 4) | (uint32_t(a_signed) << 7) | (uint32_t(b_signed) << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

struct ProbeArgs {
    int N, K;                 // MMA N (multiple of 16), K (multiple of 32)
    int chunk_major;          // layout of both operands: 0: row-group major (cores of one 8-row group contiguous along K)
                              //                          1: K-chunk major (all rows of one 16-byte K chunk contiguous)
    int swap;                 // 0: LBO = K-direction stride, SBO = M/N-direction stride; 1: the other way round
    int a_signed, b_signed;
    int use_bulk;             // stage B through cp.async.bulk from a pre-arranged global image instead of thread stores
    const int8_t *A, *B;      // A (128, K), B (N, K) row-major int8 (bit patterns)
    const int8_t *Bimg;       // B already in the shared-memory layout (use_bulk)
    int32_t *C;               // (128, N)
};

__device__ __forceinline__ uint32_t op_offset(int r, int k, int rows, int K, int chunk_major) {
    const int rg = r >> 3, ri = r & 7, kc = k >> 4, ki = k & 15;
    if (chunk_major) return uint32_t(kc * rows * 16 + r * 16 + ki);            // S_K = rows*16, S_M = 128
    return uint32_t(rg * (K / 16) * 128 + kc * 128 + ri * 16 + ki);            // S_K = 128, S_M = (K/16)*128
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const ProbeArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem, *sB = smem + 128 * a.K;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + ((a.N * a.K + 1023) & ~1023));
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // operands
    for (int i = threadIdx.x; i < 128 * a.K; i += blockDim.x) { int r = i / a.K, k = i % a.K; sA[op_offset(r, k, 128, a.K, a.chunk_major)] = uint8_t(a.A[i]); }
    if (!a.use_bulk)
        for (int i = threadIdx.x; i < a.N * a.K; i += blockDim.x) { int r = i / a.K, k = i % a.K; sB[op_offset(r, k, a.N, a.K, a.chunk_major)] = uint8_t(a.B[i]); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        if (a.use_bulk) {
            mbar_expect_tx(&bars[1], uint32_t(a.N * a.K));
            bulk_g2s(sB, a.Bimg, uint32_t(a.N * a.K), &bars[1]);
            mbar_wait(&bars[1], 0);
        }
        const uint32_t skA = a.chunk_major ? 128 * 16 : 128, smA = a.chunk_major ? 128 : (a.K / 16) * 128;
        const uint32_t skB = a.chunk_major ? a.N * 16 : 128, smB = smA;
        const uint32_t idesc = make_idesc(128, a.N, a.a_signed, a.b_signed);
        for (int kk = 0; kk < a.K / 32; ++kk) {
            const uint64_t da = a.swap ? make_desc(smem_u32(sA) + kk * 2 * skA, smA, skA) : make_desc(smem_u32(sA) + kk * 2 * skA, skA, smA);
            const uint64_t db = a.swap ? make_desc(smem_u32(sB) + kk * 2 * skB, smB, skB) : make_desc(smem_u32(sB) + kk * 2 * skB, skB, smB);
            mma_i8(tmem_base, da, db, idesc, kk > 0 ? 1u : 0u);
        }
        mma_commit(&bars[0]);
    }
    mbar_wait(&bars[0], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < a.N; c0 += 16) {
        int32_t v[16];
        tmem_ld16(tmem_base + (uint32_t(warp * 32) << 16) + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) a.C[(warp * 32 + lane) * a.N + c0 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

static int run_probe(int N, int K, int chunk_major, int swap, int a_signed, int b_signed, int use_bulk, int bweight) {
    std::vector<int8_t> A(128 * K), B(N * K), Bimg(N * K);
    srand(1234 + N * 7 + K);
    for (auto &v : A) v = a_signed ? int8_t(rand() % 127 - 63) : int8_t(uint8_t(rand() % 200));
    for (auto &v : B) {
        int r = rand() % 4;
        if (bweight) v = int8_t(uint8_t(r == 0 ? bweight : (r == 1 ? 1 : 0)));      // 0 / 1 / weight, as the membership tables
        else v = b_signed ? int8_t(rand() % 127 - 63) : int8_t(uint8_t(rand() % 200));
    }
    for (int r = 0; r < N; ++r)
        for (int k = 0; k < K; ++k) {
            const int rg = r >> 3, ri = r & 7, kc = k >> 4, ki = k & 15;
            const size_t off = chunk_major ? size_t(kc) * N * 16 + r * 16 + ki : size_t(rg) * (K / 16) * 128 + kc * 128 + ri * 16 + ki;
            Bimg[off] = B[r * K + k];
        }
    std::vector<int32_t> ref(128 * N), got(128 * N);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            int32_t s = 0;
            for (int k = 0; k < K; ++k) {
                const int av = a_signed ? int(A[m * K + k]) : int(uint8_t(A[m * K + k]));
                const int bv = b_signed ? int(B[n * K + k]) : int(uint8_t(B[n * K + k]));
                s += av * bv;
            }
            ref[m * N + n] = s;
        }
    int8_t *dA, *dB, *dBi; int32_t *dC;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dBi, B.size())); CK(cudaMalloc(&dC, ref.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBi, Bimg.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xEE, ref.size() * 4));
    ProbeArgs pa{N, K, chunk_major, swap, a_signed, b_signed, use_bulk, dA, dB, dBi, dC};
    const size_t smem = 128 * K + ((N * K + 1023) & ~1023) + 1024 + 256;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    probe_kernel<<<1, 128, smem>>>(pa);
    cudaError_t e = cudaDeviceSynchronize();
    int bad = -1;
    if (e == cudaSuccess) {
        CK(cudaMemcpy(got.data(), dC, got.size() * 4, cudaMemcpyDeviceToHost));
        bad = 0;
        for (size_t i = 0; i < ref.size(); ++i) bad += (ref[i] != got[i]);
    }
    printf("probe N=%3d K=%3d layout=%s desc=%s A=%s B=%s%s bulk=%d : %s (%d of %d wrong)%s\n", N, K,
           chunk_major ? "chunk-major" : "rowgroup-major", swap ? "LBO=MN,SBO=K" : "LBO=K,SBO=MN", a_signed ? "s8" : "u8",
           b_signed ? "s8" : "u8", bweight ? " (0/1/w)" : "", use_bulk, bad == 0 ? "OK" : "MISMATCH", bad, 128 * N,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("fatal: device error, stopping\n"); exit(3); }
    cudaFree(dA); cudaFree(dB); cudaFree(dBi); cudaFree(dC);
    return bad;
}

// ---- int8 tensor-pipe peak ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, int n, unsigned long long *cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 64 * 1024);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(smem)[i] = (i * 2654435761u) & 0x3F3F3F3Fu;           // digits 0..63
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        // SWIZZLE_128B K-major tiles: A 128 x 128 B (16 KB), B n x 128 B (up to 32 KB); 4 K steps of 32 per tile
        const uint32_t sa = smem_u32(smem), sb = sa + 16 * 1024;
        const uint32_t idesc = make_idesc(128, n, 1, 1);
        const uint64_t hi = (uint64_t(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61) | (1ull << 16);
        const unsigned long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t da = hi | uint64_t(((sa + kk * 32) & 0x3FFFFu) >> 4), db = hi | uint64_t(((sb + kk * 32) & 0x3FFFFu) >> 4);
                mma_i8(tmem_base + ((it & 1) ? 256 : 0), da, db, idesc, 1u);
            }
        }
        mma_commit(&bars[0]);
        mbar_wait(&bars[0], 0);
        const unsigned long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

static void run_peak(int n, int iters, int grid, int reps) {
    unsigned long long *dcy;
    CK(cudaMalloc(&dcy, grid * 8));
    const size_t smem = 64 * 1024 + 1024 + 256;
    CK(cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    peak_kernel<<<grid, 128, smem>>>(iters / 8 + 1, n, dcy);
    CK(cudaDeviceSynchronize());
    float best = 1e30f, total = 0.f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        peak_kernel<<<grid, 128, smem>>>(iters, n, dcy);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
        total += ms;
    }
    std::vector<unsigned long long> cy(grid);
    CK(cudaMemcpy(cy.data(), dcy, grid * 8, cudaMemcpyDeviceToHost));
    unsigned long long mx = 0;
    for (auto c : cy) mx = c > mx ? c : mx;
    const double ops = 2.0 * 128 * n * 128 * double(iters) * grid;          // 4 K steps of 32 per iteration
    printf("int8 peak: grid=%d N=%d iters=%d  best %.3f ms = %.1f TOP/s (burst), mean of %d back-to-back %.3f ms = %.1f TOP/s; "
           "%.2f cycles per M128 N%d K32 MMA (slowest CTA)\n", grid, n, iters, best, ops / best * 1e-9, reps, total / reps,
           ops / (total / reps) * 1e-9, double(mx) / (4.0 * iters), n);
    cudaFree(dcy);
}

struct Case { int N, K, cm, sw, as, bs, bulk, w; };

int main(int argc, char **argv) {
    std::vector<Case> cases;
    // descriptor convention (a wrong convention should give wrong numbers, not a fault; every case still runs in its own
    // process: `umma_probe <index>`, `umma_probe count`, `umma_probe peak`)
    for (int sw = 0; sw < 2; ++sw)
        for (int cm = 0; cm < 2; ++cm) { cases.push_back({32, 64, cm, sw, 1, 1, 0, 0}); cases.push_back({80, 96, cm, sw, 1, 1, 0, 0}); }
    for (int sw = 0; sw < 2; ++sw) {
        cases.push_back({32, 192, 1, sw, 1, 1, 0, 64});
        cases.push_back({96, 192, 1, sw, 1, 1, 1, 64});
        cases.push_back({256, 64, 1, sw, 1, 1, 1, 64});
        cases.push_back({80, 64, 1, sw, 1, 1, 1, 64});
        cases.push_back({80, 64, 1, sw, 1, 0, 1, 128});     // u8 weights 128 against s8 digits
        cases.push_back({80, 64, 1, sw, 0, 1, 1, 64});      // u8 digits against s8 weights
        cases.push_back({80, 64, 1, sw, 0, 0, 1, 128});     // both unsigned
    }
    if (argc > 1 && !strcmp(argv[1], "count")) { printf("%zu\n", cases.size()); return 0; }
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (argc > 1 && strcmp(argv[1], "peak")) {
        const int i = atoi(argv[1]);
        if (i < 0 || i >= (int)cases.size()) return 1;
        const Case &c = cases[i];
        return run_probe(c.N, c.K, c.cm, c.sw, c.as, c.bs, c.bulk, c.w) == 0 ? 0 : 1;
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, sms);
    run_peak(256, 4096, sms, 5);
    run_peak(256, 1 << 20, sms, 5);        // ~0.3 s per launch, 1.5 s back to back: the figure under the power cap
    run_peak(128, 8192, sms, 3);
    run_peak(64, 16384, sms, 3);
    run_peak(32, 16384, sms, 3);
    run_peak(80, 16384, sms, 3);
    return 0;
}
