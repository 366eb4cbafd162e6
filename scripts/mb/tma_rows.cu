// Microbenchmark: gather bandwidth of whole rows through 1-D bulk (TMA) copies, one copy per row, into a ring of
// shared-memory stages — the load path of k_update_steps3 (update_kernel3.cuh).  Open question it answers: does the
// TMA unit sustain ~5 TB/s with 400-byte operations (one per row), issued by the 32 lanes of one producer warp per
// CTA (the compiler serialises them: ~9 SASS instructions per lane)?  Compared with the LSU gather of gather.cu.
// Self-checking: the consumers sum the first float of every row, the host verifies the total.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_rows tma_rows.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

constexpr int SR = 30;        // rows per stage (one producer lane per row)
constexpr int NCONS = 8;      // consumer warps
constexpr int THREADS = 32 * (1 + NCONS);

// rows [lo, hi) of idx are this CTA's share; D stages; every stage is consumed by one warp (round robin, D >= NCONS)
__global__ void __launch_bounds__(THREADS, 1) k_tma_rows(const float* __restrict__ src, const int* __restrict__ idx, int nrows, int KS,
                                                          int D, double* out) {
  extern __shared__ __align__(16) float smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + D;
  float* ring = reinterpret_cast<float*>(empty + D);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < D; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t per = (nrows + gridDim.x - 1) / gridDim.x;
  const int lo = (int)(blockIdx.x * per), hi = (int)(((blockIdx.x + 1) * per < nrows) ? (blockIdx.x + 1) * per : nrows);
  const int nst = (hi > lo) ? (hi - lo + SR - 1) / SR : 0;
  const uint32_t row_bytes = (uint32_t)KS * 4u;
  if (warp == 0) {
    for (int g = 0; g < nst; ++g) {
      const int slot = g % D, use = g / D;
      if (use >= 1) mbar_wait(empty + slot, (use - 1) & 1);
      const int r0 = lo + g * SR;
      const int nr = (hi - r0 < SR) ? hi - r0 : SR;
      const int cell = (lane < nr) ? idx[r0 + lane] : 0;
      if (lane == 0) mbar_expect(full + slot, (uint32_t)nr * row_bytes);
      __syncwarp();
      if (lane < nr) bulk_load(ring + ((size_t)slot * SR + lane) * KS, src + (size_t)cell * KS, row_bytes, full + slot);
    }
  } else {
    const int w = warp - 1;
    double acc = 0.0;
    for (int g = w; g < nst; g += NCONS) {
      const int slot = g % D, use = g / D;
      mbar_wait(full + slot, use & 1);
      const int r0 = lo + g * SR;
      const int nr = (hi - r0 < SR) ? hi - r0 : SR;
      if (lane < nr) acc += ring[((size_t)slot * SR + lane) * KS];  // first float of the row
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(out, acc);
  }
}

int main() {
  const int N = 1000000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int KS : {100, 52, 200}) {
    std::vector<float> hsrc((size_t)N * KS);
    for (int i = 0; i < N; ++i) hsrc[(size_t)i * KS] = (float)(i % 1000);
    float* d;
    cudaMalloc(&d, hsrc.size() * 4);
    cudaMemcpy(d, hsrc.data(), hsrc.size() * 4, cudaMemcpyHostToDevice);
    // the update kernel's pattern: a sorted 5 % sample of the rows per block step; here 20 such samples back to back
    std::vector<int> order(N);
    std::iota(order.begin(), order.end(), 0);
    std::mt19937 rng(1);
    std::shuffle(order.begin(), order.end(), rng);
    for (int b = 0; b < 20; ++b) std::sort(order.begin() + b * (N / 20), order.begin() + (b + 1) * (N / 20));
    double want = 0;
    for (int i = 0; i < N; ++i) want += order[i] % 1000;
    int* didx;
    cudaMalloc(&didx, N * 4);
    cudaMemcpy(didx, order.data(), N * 4, cudaMemcpyHostToDevice);
    double* dout;
    cudaMalloc(&dout, 8);
    for (int D : {8, 16}) {
      const size_t smem = 2 * (size_t)D * 8 + (size_t)D * SR * KS * 4;
      if (smem > 227 * 1024) continue;
      cudaFuncSetAttribute(k_tma_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      float best = 1e9;
      double got = 0;
      for (int rep = 0; rep < 6; ++rep) {
        cudaMemset(dout, 0, 8);
        cudaEventRecord(e0);
        k_tma_rows<<<sms, THREADS, smem>>>(d, didx, N, KS, D, dout);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
        cudaMemcpy(&got, dout, 8, cudaMemcpyDeviceToHost);
      }
      printf("KS=%3d (%3d-byte rows) stages=%2d: %.3f ms  %.2f TB/s  %s (err=%s)\n", KS, KS * 4, D, best,
             (double)N * KS * 4 / best * 1e-9, got == want ? "sum ok" : "SUM MISMATCH", cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(d);
    cudaFree(didx);
    cudaFree(dout);
  }
  return 0;
}
