// Read-only streaming bandwidth probe (not part of the product): what does a trivially simple kernel reach on this
// B200 for (a) LDG.128 grid-stride reads and (b) 16 KB TMA bulk copies into a shared-memory ring?  Used to put the
// attention kernels' achieved GB/s into perspective (MEASURED_PEAKS.json's hbm_gbs is a read+write copy).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/read_bw tools/read_bw.cu && ./tools/read_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void ldg_read(const uint4* __restrict__ p, size_t n, unsigned* out) {
  unsigned acc = 0;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
    acc ^= a.x ^ b.y ^ c.z ^ d.w;
  }
  for (; i < n; i += stride) acc ^= p[i].x;
  if (acc == 0x12345678u) *out = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NST, int STG>
__global__ void tma_read(const uint8_t* __restrict__ p, size_t bytes_per_cta, unsigned* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[NST];
  const uint8_t* src = p + (size_t)blockIdx.x * bytes_per_cta;
  const int n = (int)(bytes_per_cta / STG);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned acc = 0;
  for (int i = 0; i < n + NST - 1; ++i) {
    if (threadIdx.x == 0 && i < n) {   // issue chunk i (the stage was consumed in iteration i-NST, guarded by the syncthreads below)
      const int st = i % NST;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(STG) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + st * STG)),
                   "l"(src + (size_t)i * STG), "r"(STG), "r"(s32(&full[st]))
                   : "memory");
    }
    const int j = i - (NST - 1);
    if (j >= 0) {
      const int st = j % NST;
      const uint32_t ph = (j / NST) & 1;
      uint32_t done;
      do {
        asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.b32 %0, 1, 0, q; }" : "=r"(done) : "r"(s32(&full[st])), "r"(ph) : "memory");
      } while (!done);
      acc ^= reinterpret_cast<const unsigned*>(smem + st * STG)[threadIdx.x];
      __syncthreads();
    }
  }
  if (acc == 0x12345678u) *out = acc;
}

int main() {
  const size_t bytes = 1ull << 30;
  uint8_t* d; unsigned* o;
  cudaMalloc(&d, bytes); cudaMalloc(&o, 4); cudaMemset(d, 1, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas_per_sm : {2, 4, 8}) {
    const int grid = 148 * ctas_per_sm;
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
      cudaEventRecord(e0); ldg_read<<<grid, 512>>>((const uint4*)d, bytes / 16, o); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("LDG.128 read 1 GiB, %d CTAs/SM x 512 thr: %.1f GB/s\n", ctas_per_sm, bytes / best / 1e6);
  }
  {
    constexpr int NST = 6, STG = 16384;
    auto k = tma_read<NST, STG>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, NST * STG);
    for (int ctas_per_sm : {1, 2}) {
      const int grid = 148 * ctas_per_sm;
      const size_t per = (bytes / grid) / STG * STG;
      float best = 1e9;
      for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0); k<<<grid, 128, NST * STG>>>(d, per, o); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("TMA bulk 16 KB x %d stages, %d CTAs/SM: %.1f GB/s (%s)\n", NST, ctas_per_sm, (double)per * grid / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  // small working set (64 MB, L2 resident), repeated
  {
    const size_t small = 64ull << 20;
    float best = 1e9;
    for (int it = 0; it < 10; ++it) {
      cudaEventRecord(e0); ldg_read<<<148 * 8, 512>>>((const uint4*)d, small / 16, o); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("LDG.128 read 64 MiB (L2-resident after first pass): %.1f GB/s, %.1f us\n", small / best / 1e6, best * 1e3);
  }
  // the same small working sets through the TMA bulk-copy ring, repeated: do bulk copies allocate / hit in L2?
  {
    constexpr int NST = 6, STG = 16384;
    auto k = tma_read<NST, STG>;
    for (size_t mb : {32, 48, 64, 96}) {
      const int grid = 148 * 2;
      const size_t per = ((mb << 20) / grid) / STG * STG;
      float best = 1e9, first = 0;
      cudaMemset(d + (512ull << 20), 2, 256ull << 20);         // push the set out of L2 first
      for (int it = 0; it < 10; ++it) {
        cudaEventRecord(e0); k<<<grid, 128, NST * STG>>>(d, per, o); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it == 0) first = ms; if (ms < best) best = ms;
      }
      printf("TMA bulk read %zu MiB repeated: first pass %.1f us, best %.1f us = %.1f GB/s\n", mb, first * 1e3, best * 1e3, (double)per * grid / best / 1e6);
    }
  }
  return 0;
}
