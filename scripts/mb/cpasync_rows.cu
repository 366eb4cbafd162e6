// Microbenchmark: how fast can P warps of a persistent CTA (one per SM) gather 400-byte rows with 16-byte cp.async
// into shared memory, D batches of 8 rows in flight per warp, completion tracked per batch by an mbarrier
// (cp.async.mbarrier.arrive.noinc)?  This is the load path of k_update_steps4 / k_assign_tc3.  Also: the same gather
// through registers (ld.global.v4 -> st.shared.v4, the path of the first persistent generation).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o cpasync_rows cpasync_rows.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, unsigned par) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}

constexpr int BR = 8;  // rows per batch

// mode 0: cp.async + mbarrier; mode 1: ld.global.v4 (8 rows in registers) -> st.shared
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_gather(const float* __restrict__ src, const int* __restrict__ idx, int nrows, int KS, int P,
                                                    int D, float* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KS4 = KS >> 2;
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)warp * D * BR * KS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) + (size_t)P * D * BR * KS) + (size_t)warp * D;
  if (lane == 0)
    for (int i = 0; i < D; ++i) mbar_init(bars + i, 32);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (warp >= P) return;
  const int gw = blockIdx.x * P + warp, nw = gridDim.x * P;
  const int nbatch = nrows / BR;
  float acc = 0.f;
  int it = 0;
  for (int b = gw; b < nbatch; b += nw, ++it) {
    const int slot = it % D, use = it / D;
    float* dst = ring + (size_t)slot * BR * KS;
    if (MODE == 0) {
      if (use >= 1) {
        while (!mbar_try(bars + slot, (use - 1) & 1)) {
        }
        acc += dst[lane];  // "consume"
      }
      const int cell = (lane < BR) ? __ldg(idx + b * BR + lane) : 0;
#pragma unroll
      for (int r = 0; r < BR; ++r) {
        const int cr = __shfl_sync(0xffffffffu, cell, r);
        if (lane < KS4) {
          const unsigned sp = smem_u32(dst + (size_t)r * KS + 4 * lane);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp), "l"(src + (size_t)cr * KS + 4 * lane) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bars + slot)) : "memory");
    } else {
      const int cell = (lane < BR) ? __ldg(idx + b * BR + lane) : 0;
      float4 v[BR];
#pragma unroll
      for (int r = 0; r < BR; ++r) {
        const int cr = __shfl_sync(0xffffffffu, cell, r);
        v[r] = make_float4(0, 0, 0, 0);
        if (lane < KS4) v[r] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)cr * KS) + lane);
      }
#pragma unroll
      for (int r = 0; r < BR; ++r)
        if (lane < KS4) *reinterpret_cast<float4*>(dst + (size_t)r * KS + 4 * lane) = v[r];
      acc += dst[lane];
    }
  }
  if (MODE == 0) asm volatile("cp.async.wait_all;" ::: "memory");
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  const int N = 1000000, KS = 100;
  float* d;
  cudaMalloc(&d, (size_t)N * KS * 4);
  cudaMemset(d, 0, (size_t)N * KS * 4);
  float* out;
  cudaMalloc(&out, 4);
  std::vector<int> idx(N);
  std::iota(idx.begin(), idx.end(), 0);
  std::mt19937 rng(1);
  std::shuffle(idx.begin(), idx.end(), rng);
  // the update kernel's pattern: 20 blocks, each an ascending random 1/20 subset
  std::vector<int> blk(N), order(N);
  for (int i = 0; i < N; ++i) blk[idx[i]] = std::min(i / (N / 20), 19);
  {
    std::vector<int> cnt(21, 0);
    for (int i = 0; i < N; ++i) cnt[blk[i] + 1]++;
    for (int j = 0; j < 20; ++j) cnt[j + 1] += cnt[j];
    for (int i = 0; i < N; ++i) order[cnt[blk[i]]++] = i;
  }
  int* di;
  cudaMalloc(&di, N * 4);
  cudaMemcpy(di, order.data(), N * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(k_gather<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
  cudaFuncSetAttribute(k_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int P : {1, 2, 4, 8, 16, 32})
      for (int D : {2, 4, 8, 16}) {
        const size_t smem = (size_t)P * D * BR * KS * 4 + (size_t)P * D * 8 + 256;
        if (smem > 227 * 1024 - 1024) continue;
        if (mode == 1 && D != 2) continue;
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
          cudaEventRecord(e0);
          if (mode == 0)
            k_gather<0><<<148, 32 * P, smem>>>(d, di, N, KS, P, D, out);
          else
            k_gather<1><<<148, 32 * P, smem>>>(d, di, N, KS, P, D, out);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          best = std::min(best, ms);
        }
        printf("%s  warps/SM %2d  batches in flight/warp %2d (%3zu KB/SM): %7.3f ms  %7.1f GB/s  (%s)\n",
               mode == 0 ? "cp.async+mbarrier" : "ld.global->st.shared", P, D, (size_t)P * D * BR * KS * 4 / 1024, best,
               (double)N * KS * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
