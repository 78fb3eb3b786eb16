// gather_bw.cu — developer aid: what HBM sustains for the ROIAlign ACCESS PATTERN, independent of any ROIAlign kernel.
//   work item = read one random `chunk`-byte block (16-byte aligned, inside a `src_mb` MB buffer) and write `wr` x chunk
//   bytes to a linear output stream (ROIAlign 14x14 on the bench step: ~10 KiB rows read, ~2-3x as many bytes written).
// Plain LDG.128 / STG.128 (streaming stores), one warp per item, 8 independent 16-byte loads in flight per lane, grid
// sized for full occupancy: an upper bound for any gather kernel with this traffic shape.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/gather_bw tools/ubench/gather_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void gather_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ offs, float4* __restrict__ dst,
                              int n_items, int chunk16, int wr) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int it = warp; it < n_items; it += nwarps) {
    const float4* s = src + offs[it];
    float4* d = dst + (size_t)it * chunk16 * wr;
    for (int i = lane; i < chunk16; i += 32 * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = i + 32 * u;
        v[u] = j < chunk16 ? __ldg(s + j) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = i + 32 * u;
        if (j < chunk16)
          for (int w = 0; w < wr; ++w) __stcs(d + (size_t)w * chunk16 + j, v[u]);
      }
    }
  }
}

int main(int argc, char** argv) {
  const size_t src_mb = 178;
  const int configs[][2] = {{1024, 1}, {1024, 3}, {2048, 1}, {10240, 1}, {10240, 2}, {10240, 3}, {65536, 2}, {1 << 20, 2}};
  float4 *src, *dst;
  const size_t src_bytes = src_mb << 20, dst_bytes = (size_t)1200 << 20;
  cudaMalloc(&src, src_bytes);
  cudaMalloc(&dst, dst_bytes);
  cudaMemset(src, 1, src_bytes);
  cudaMemset(dst, 0, dst_bytes);
  uint32_t* d_offs;
  cudaMalloc(&d_offs, 4 << 20);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  printf("chunk_bytes wr_ratio items  read_MB write_MB  ms  GB/s(read+write)\n");
  for (auto& c : configs) {
    const int chunk = c[0], wr = c[1];
    const size_t read_total = (size_t)190 << 20;
    int n = (int)(read_total / chunk);
    if ((size_t)n * chunk * wr > dst_bytes) n = (int)(dst_bytes / ((size_t)chunk * wr));
    uint32_t* h = (uint32_t*)malloc((size_t)n * 4);
    srand(7);
    const uint32_t slots = (uint32_t)((src_bytes - chunk) / 1024);
    for (int i = 0; i < n; ++i) h[i] = (uint32_t)(((uint64_t)rand() * 32768u + rand()) % slots) * 64u;   // 1 KiB aligned, in float4 units
    cudaMemcpy(d_offs, h, (size_t)n * 4, cudaMemcpyHostToDevice);
    free(h);
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaMemset(dst, rep, 300 << 20);   // flush L2 with something else
      cudaEventRecord(a);
      gather_kernel<<<148 * 8, 256>>>(src, d_offs, dst, n, chunk / 16, wr);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (rep >= 1 && ms < best) best = ms;
    }
    const double rd = (double)n * chunk, wrb = rd * wr;
    printf("%10d %8d %6d %8.1f %8.1f %7.4f %8.0f\n", chunk, wr, n, rd / 1e6, wrb / 1e6, best, (rd + wrb) / best / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
