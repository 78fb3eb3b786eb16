// topk.cuh — internal interface of the segmented radix select + sort (topk.cu).
#pragma once
#include "common.cuh"

namespace od {

// Number of index bits used in the composite key for rows of `cols` elements.
inline int index_bits(int64_t cols) {
  int ib = 1;
  while (((int64_t)1 << ib) < cols) ++ib;
  return ib;
}
inline int64_t next_pow2(int64_t v) {
  int64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Workspace bytes needed by topk_launch for (rows, cols, k).
size_t topk_workspace_bytes(int64_t rows, int64_t cols, int64_t k);

// k largest per row of a strided fp32 matrix, sorted by (value desc, index asc).
// idx_out [rows,k] i32 (required); val_out [rows,k] f32 or nullptr (gathered original values).
// All work is enqueued on `st`; ws must hold topk_workspace_bytes(...).
int topk_launch(const float* scores, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride, int64_t k,
                int32_t* idx_out, float* val_out, void* ws, size_t ws_bytes, cudaStream_t st);

// In-place descending sort of `rows` segments of n_pow2 uint64 keys (n_pow2 a power of two; pad with 0).
int sort_u64_desc_launch(unsigned long long* keys, int64_t rows, int64_t n_pow2, cudaStream_t st);

// Block-wide bitonic sort (descending) of n_pow2 keys in shared memory; all threads of the CTA call it.
__device__ __forceinline__ void block_bitonic_sort_desc(unsigned long long* s, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const bool desc = ((i & k) == 0);
        const unsigned long long a = s[i], b = s[l];
        if ((a < b) == desc) {
          s[i] = b;
          s[l] = a;
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace od
