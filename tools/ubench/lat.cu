// Dependent-chain latencies of the warp primitives the NMS scan's critical path is made of (developer aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int OP>
__global__ void chain(long long* out, int iters) {
  __shared__ unsigned long long sm[64];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 64) sm[threadIdx.x] = threadIdx.x * 0x9E3779B97F4A7C15ull;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");   // phase 0 complete
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  unsigned long long x = sm[lane];
  uint32_t y = (uint32_t)x | 1u;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) {   // shared load, address depends on the previous value
      x = sm[(x ^ i) & 63];
    } else if (OP == 1) {   // ballot
      y = __ballot_sync(0xffffffffu, ((y >> lane) & 1u) != 0u) + i;
    } else if (OP == 2) {   // REDUX.OR
      y = __reduce_or_sync(0xffffffffu, y ^ (i << lane)) >> 1;
    } else if (OP == 3) {   // popc
      y = __popc(y) + (y << 3) + i;
    } else if (OP == 4) {   // shuffle
      y = __shfl_xor_sync(0xffffffffu, y, 1) + i;
    } else if (OP == 5) {   // mbarrier test_wait on a completed phase
      uint32_t d;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(d) : "r"(smem_u32(&bar) + (y & 0u)), "r"(0u) : "memory");
      y += d;
    } else if (OP == 6) {   // mbarrier try_wait on a completed phase
      uint32_t d;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(d) : "r"(smem_u32(&bar) + (y & 0u)), "r"(0u) : "memory");
      y += d;
    } else if (OP == 7) {   // named barrier arrive (count 64: never completes a wait, two arrivals flip it)
      asm volatile("bar.arrive 5, 64;" ::: "memory");
      y += i;
    } else if (OP == 8) {   // shared store then load of the same word (lane 0 writes, all read): the ring hand-off
      if (lane == 0) sm[0] = x + i;
      __syncwarp();
      x = sm[0];
    } else if (OP == 9) {   // 64-bit shared atomicOr (CAS loop) by lane 0
      if (lane == 0) atomicOr(&sm[1], x | i);
      __syncwarp();
      x = sm[1];
    }
  }
  const long long t1 = clock64();
  if (lane == 0) {
    out[0] = t1 - t0;
    out[1] = (long long)(x + y);
  }
}

// Same chains on warp 0 while warp 1 keeps `bytes`-sized cp.async.bulk copies (global -> shared) in flight.
template <int OP>
__global__ void chain_tma(long long* out, int iters, const unsigned long long* src, int bytes) {
  extern __shared__ __align__(128) unsigned long long dyn[];   // [bytes/8] copy target
  __shared__ unsigned long long sm[64];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ volatile int stop;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 64) sm[threadIdx.x] = threadIdx.x * 0x9E3779B97F4A7C15ull;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    uint32_t parity = 0;
    int n = 0;
    while (!stop && n < 100000) {
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dyn)), "l"(src),
                     "r"((uint32_t)bytes), "r"(smem_u32(&bar))
                     : "memory");
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
      }
      __syncwarp();
      parity ^= 1u;
      ++n;
    }
    if (lane == 0) out[2] = n;
    return;
  }
  if (warp != 0) return;
  unsigned long long x = sm[lane];
  uint32_t y = (uint32_t)x | 1u;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) x = sm[(x ^ i) & 63];
    else if (OP == 1) y = __ballot_sync(0xffffffffu, ((y >> lane) & 1u) != 0u) + i;
    else if (OP == 2) y = __reduce_or_sync(0xffffffffu, y ^ (i << lane)) >> 1;
  }
  const long long t1 = clock64();
  if (lane == 0) {
    out[0] = t1 - t0;
    out[1] = (long long)(x + y);
    stop = 1;
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 32);
  const char* names[] = {"LDS.64 dependent", "ballot", "REDUX.OR", "POPC+ALU", "SHFL", "mbarrier.test_wait (done)", "mbarrier.try_wait (done)",
                         "bar.arrive", "STS->syncwarp->LDS", "ATOMS.OR.64 -> LDS"};
  const int iters = 256;
  for (int op = 0; op < 10; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: chain<0><<<1, 64>>>(d, iters); break;
        case 1: chain<1><<<1, 64>>>(d, iters); break;
        case 2: chain<2><<<1, 64>>>(d, iters); break;
        case 3: chain<3><<<1, 64>>>(d, iters); break;
        case 4: chain<4><<<1, 64>>>(d, iters); break;
        case 5: chain<5><<<1, 64>>>(d, iters); break;
        case 6: chain<6><<<1, 64>>>(d, iters); break;
        case 7: chain<7><<<1, 64>>>(d, iters); break;
        case 8: chain<8><<<1, 64>>>(d, iters); break;
        case 9: chain<9><<<1, 64>>>(d, iters); break;
      }
      cudaDeviceSynchronize();
    }
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-28s %7.1f cycles per op\n", names[op], (double)h[0] / iters);
  }
  unsigned long long* src;
  cudaMalloc(&src, 1 << 20);
  cudaMemset(src, 1, 1 << 20);
  cudaFuncSetAttribute(chain_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(chain_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(chain_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int bytes : {512, 12288, 49152}) {
    for (int op = 0; op < 3; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) chain_tma<0><<<1, 64, 64 * 1024>>>(d, 4096, src, bytes);
        if (op == 1) chain_tma<1><<<1, 64, 64 * 1024>>>(d, 4096, src, bytes);
        if (op == 2) chain_tma<2><<<1, 64, 64 * 1024>>>(d, 4096, src, bytes);
        cudaDeviceSynchronize();
      }
      long long h[3];
      cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
      printf("with %5d-byte bulk copies in flight: %-18s %7.1f cycles per op (%lld copies, %.0f cycles per copy)\n", bytes, names[op],
             (double)h[0] / 4096, h[2], (double)h[0] / (double)(h[2] > 0 ? h[2] : 1));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
