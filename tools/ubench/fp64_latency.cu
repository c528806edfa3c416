// Dependent-chain latency and per-SM throughput of the FP64 instructions the solve and posterior kernels lean on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double *out, long long *cyc, double a, double b, int iters) {
    double x = a + threadIdx.x * 1e-9, y = b;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = fma(x, y, b);
            else if (OP == 1) x = x * y;
            else if (OP == 2) x = x + y;
            else if (OP == 3) x = rsqrt(x) + b;
            else if (OP == 4) x = sqrt(x) + b;
            else if (OP == 5) x = (double)__fmaf_rn((float)x, (float)y, (float)b);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void thr(double *out, long long *cyc, double a, double b, int iters) {
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = a + threadIdx.x * 1e-9 + k;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], b, a);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 256;
    const char *names[] = {"DFMA", "DMUL", "DADD", "rsqrt(double)+DADD", "sqrt(double)+DADD", "cvt+FFMA+cvt"};
#define RUN(OP) chain<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, iters); chain<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, iters); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-22s dependent latency %.1f cycles\n", names[OP], double(h) / (iters * 16));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        thr<8><<<1, warps * 32>>>(out, cyc, 0.5, 0.9999999, iters);
        thr<8><<<1, warps * 32>>>(out, cyc, 0.5, 0.9999999, iters);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA throughput, %2d warps x ILP 8 on one SM: %.1f lane-FMA/clk\n", warps, double(iters) * 4 * 8 * warps * 32 / double(h));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
