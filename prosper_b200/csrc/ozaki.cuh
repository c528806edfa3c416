// FP64-accurate GEMMs on the int8 tcgen05 tensor cores (ozaki.cu): operand slicing + the sliced product.
#pragma once
#include "common.cuh"

namespace pet {

// ns int8 planes of a K-contiguous operand: plane t holds rows of Kp bytes, `row_stride` bytes apart
struct OzOperand {
    const int8_t *slices;
    int64_t row_stride, slice_stride;
    const double *scale;          // per-row power-of-two scale
};

int ozaki_kp(int64_t K);
// max_kb_per_split: 1024 K blocks of 64 when both operands hold signed slices in [-64, 64]; 512 when one of them holds
// unsigned 7-bit digits (gl_post_slice: products up to 64 * 127)
int ozaki_splits(int64_t M, int64_t N, int Kp, int sm_count, int max_kb_per_split = 1024);
int ozaki_slice_rows(const double *X, int64_t ldx, int64_t rows, int K, int ns, int8_t *out, int64_t slice_stride, double *scale,
                     cudaStream_t st);
int ozaki_slice_cols(const double *X, int64_t ldx, int64_t rows, int cols, int ns, unsigned long long *colmax, bool have_colmax, int8_t *out,
                     int64_t row_stride, int64_t slice_stride, double *scale, cudaStream_t st,
                     const double *rowscale = nullptr, int64_t rs_stride = 0);
int ozaki_gemm(int64_t M, int64_t N, int Kp, int ns, const OzOperand &A, const OzOperand &B, double *C, int64_t ldc, int splits,
               int64_t split_stride, bool accumulate, int sm_count, cudaStream_t st, int max_pair_product = 4096);
int ozaki_add_slabs(double *dst, const double *slabs, int64_t count, int n_slabs, cudaStream_t st);
// dst (cols, ld_dst) = transpose of the sum of n_slabs slabs of shape (rows, ld_src), slab_stride doubles apart
int ozaki_add_slabs_t(double *dst, int64_t ld_dst, const double *slabs, int rows, int cols, int64_t ld_src, int64_t slab_stride,
                      int n_slabs, cudaStream_t st);

}  // namespace pet
