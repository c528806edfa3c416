// Inference on the posterior of one E-step (camodels/__init__.py:255-375): per datapoint the normalised
// log-posterior over all evaluated states, the K most probable states and -- for the binary layout
// [null | h = 0..H-1 | multi-cause states] -- the marginal log-probability of every cause.
// One CTA per row of logpj (grid-stride), the row in shared memory.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "gl_kernel.cuh"

namespace pet {

constexpr int INF_THREADS = 128;

__device__ __forceinline__ double block_max(double v, double *red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double m = red[0];
    for (int w = 1; w < INF_THREADS / 32; ++w) m = fmax(m, red[w]);
    return m;
}
__device__ __forceinline__ double block_sum(double v, double *red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < INF_THREADS / 32; ++w) s += red[w];
    return s;
}

struct InferArgs {
    const double *logpj; int64_t ld, n;
    int C, H, Hp, S, binary_layout;      // binary_layout: columns [null | H singletons (x value blocks) | S states], marginals wanted
    int multi_off, valued;               // first multi-state column; valued: records are pos | vidx << 4
    double vals[PET_MAXV];
    const unsigned long long *states;    // S records (binary: member positions, unused = Hp)
    const int *cand;                     // (n, Hp)
    int topK, logprob;
    int *idx_out; double *p_out, *m_out; // (n, topK), (n, topK), (n, H) or null
};

__global__ void __launch_bounds__(INF_THREADS) infer_kernel(const InferArgs a) {
    extern __shared__ __align__(16) double smem[];
    double *row = smem;                                   // C doubles: logpj - rowmax
    double *red = row + ((a.C + 1) & ~1);                 // 8 doubles
    int *redi = reinterpret_cast<int *>(red + 8);         // 8 ints
    unsigned short *mask = reinterpret_cast<unsigned short *>(redi + 8);   // S membership masks
    const int tid = threadIdx.x;
    const bool marg = a.binary_layout && a.m_out;
    if (marg)
        for (int s = tid; s < a.S; s += INF_THREADS) {
            const unsigned long long rec = a.states[s];
            unsigned m = 0;
            for (int b = 0; b < 8; ++b) {
                const unsigned p = unsigned(rec >> (8 * b)) & 0xFFu;
                if (a.valued) {      // dsc_et.py:1014: states whose entry at this candidate equals 1
                    if (p != 0xFFu && a.vals[p >> 4] == 1.0) m |= 1u << (p & 15u);
                } else if (p < unsigned(a.Hp)) {
                    m |= 1u << p;
                }
            }
            mask[s] = (unsigned short)m;
        }
    for (int64_t r = blockIdx.x; r < a.n; r += gridDim.x) {
        __syncthreads();
        const double *src = a.logpj + r * a.ld;
        double mx = -INFINITY;
        for (int c = tid; c < a.C; c += INF_THREADS) { const double v = src[c]; row[c] = v; mx = fmax(mx, v); }
        mx = block_max(mx, red);                                           // my_corr            (:307)
        double sum = 0.0;
        for (int c = tid; c < a.C; c += INF_THREADS) { const double t = row[c] - mx; row[c] = t; sum += exp(t); }
        sum = block_sum(sum, red);                                         // my_denomc          (:310)
        const double nlz = -log(sum);                                      // my_logpjc += -log  (:311)
        __syncthreads();
        // ---- marginals (:336-342) ---------------------------------------------------------------------
        if (marg) {
            double *mrow = a.m_out + r * a.H;
            for (int h = tid; h < a.H; h += INF_THREADS) {
                const double v = row[1 + h] + nlz;
                mrow[h] = v;
            }
            __syncthreads();
            const int *cand = a.cand + r * a.Hp;
            for (int j = 0; j < a.Hp; ++j) {
                const int h = cand[j];
                const double single = row[1 + h] + nlz;
                double m1 = (tid == 0) ? single : -INFINITY;                // scipy logsumexp: max-shifted
                for (int s = tid; s < a.S; s += INF_THREADS)
                    if ((mask[s] >> j) & 1) m1 = fmax(m1, row[a.multi_off + s] + nlz);
                m1 = block_max(m1, red);
                double acc = (tid == 0) ? exp(single - m1) : 0.0;
                for (int s = tid; s < a.S; s += INF_THREADS)
                    if ((mask[s] >> j) & 1) acc += exp((row[a.multi_off + s] + nlz) - m1);
                acc = block_sum(acc, red);
                if (tid == 0) {
                    const double v = (m1 == -INFINITY) ? -INFINITY : log(acc) + m1;
                    mrow[h] = v;
                }
            }
        }
        // ---- K most probable states (:312-333): value descending, larger column first among equals -----
        for (int k = 0; k < a.topK; ++k) {
            double bv = -INFINITY;
            int bi = -1;
            for (int c = tid; c < a.C; c += INF_THREADS) {
                const double v = row[c];
                if (v == v && (v > bv || (v == bv && c > bi) || bi < 0)) { bv = v; bi = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
            }
            __syncthreads();
            if ((tid & 31) == 0) { red[tid >> 5] = bv; redi[tid >> 5] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < INF_THREADS / 32; ++w) {
                    const double ov = red[w];
                    const int oi = redi[w];
                    if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
                }
                a.idx_out[r * a.topK + k] = bi;
                // logprob: normalised log-posterior; else exp(logpj - rowmax), NOT normalised (reference :309,:324)
                a.p_out[r * a.topK + k] = (bi < 0) ? (a.logprob ? -INFINITY : 0.0) : (a.logprob ? bv + nlz : exp(bv));
                if (bi >= 0) row[bi] = nan("");                            // taken
            }
            __syncthreads();
        }
    }
}

int launch_infer(const GLStatic &st, int C, int binary_layout, const int *cand, const double *logpj, int64_t ld, int64_t n,
                 int topK, int logprob, int *idx_out, double *p_out, double *m_out, int sm_count, cudaStream_t stream) {
    if (n <= 0) return PET_OK;
    if (topK < 1 || topK > C) { set_error("inference: topK must be in 1..%d", C); return PET_EINVAL; }
    InferArgs a{};
    a.logpj = logpj; a.ld = ld; a.n = n; a.C = C; a.H = st.H; a.Hp = st.Hp; a.S = st.S; a.binary_layout = binary_layout;
    a.multi_off = st.has_null + st.n_blocks * st.H; a.valued = st.binary ? 0 : 1;
    for (int v = 0; v < PET_MAXV; ++v) a.vals[v] = st.vals[v];
    a.states = st.states; a.cand = cand; a.topK = topK; a.logprob = logprob; a.idx_out = idx_out; a.p_out = p_out; a.m_out = m_out;
    const size_t smem = size_t((C + 1) & ~1) * 8 + 8 * 8 + 8 * 4 + size_t(st.S) * 2 + 16;
    if (smem > 227 * 1024) { set_error("inference: %d columns do not fit in shared memory", C); return PET_EINVAL; }
    static size_t configured = 0;
    if (smem > configured) {
        PET_CUDA(cudaFuncSetAttribute(infer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    const int per_sm = int(std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 1024))));
    const unsigned grid = (unsigned)std::min<int64_t>(n, int64_t(sm_count) * per_sm);
    infer_kernel<<<grid, INF_THREADS, smem, stream>>>(a);
    PET_LAUNCH_CHECK();
    return PET_OK;
}

}  // namespace pet
