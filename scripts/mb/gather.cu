// Microbenchmark: HBM read bandwidth for gathers of whole 400-byte rows (the update_R access pattern)
// vs a sequential sweep.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>

__global__ void k_seq(const float4* __restrict__ src, size_t n4, float* out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = src[i];
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) out[0] = acc;
}
// one warp handles RPW rows per iteration (lanes split), ITERS independent iterations unrolled
template <int UNR>
__global__ void k_gather(const float* __restrict__ src, const int* __restrict__ idx, int nrows, int KS, float* out) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int KS4 = KS / 4;
  float acc = 0.f;
  for (int r0 = gw * UNR; r0 < nrows; r0 += nw * UNR) {
    float4 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      int r = r0 + u;
      v[u] = make_float4(0, 0, 0, 0);
      if (r < nrows && lane < KS4) v[u] = reinterpret_cast<const float4*>(src + (size_t)idx[r] * KS)[lane];
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  const int N = 1000000, KS = 100;
  float* d;
  cudaMalloc(&d, (size_t)N * KS * 4);
  cudaMemset(d, 0, (size_t)N * KS * 4);
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto timeit = [&](auto f, double bytes, const char* name) {
    f();
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int i = 0; i < 5; ++i) {
      cudaEventRecord(e0);
      f();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
    }
    printf("%-44s %8.3f ms  %8.1f GB/s\n", name, best, bytes / best / 1e6);
  };
  timeit([&] { k_seq<<<148 * 8, 512>>>((const float4*)d, (size_t)N * KS / 4, out); }, (double)N * KS * 4, "sequential float4 sweep (400 MB)");
  std::mt19937 rng(1);
  // pattern A: all rows in random order
  std::vector<int> idx(N);
  std::iota(idx.begin(), idx.end(), 0);
  std::shuffle(idx.begin(), idx.end(), rng);
  int* di;
  cudaMalloc(&di, N * 4);
  cudaMemcpy(di, idx.data(), N * 4, cudaMemcpyHostToDevice);
  timeit([&] { k_gather<1><<<148 * 16, 256>>>(d, di, N, KS, out); }, (double)N * KS * 4, "gather all rows, random order, unroll 1");
  timeit([&] { k_gather<4><<<148 * 16, 256>>>(d, di, N, KS, out); }, (double)N * KS * 4, "gather all rows, random order, unroll 4");
  timeit([&] { k_gather<8><<<148 * 8, 256>>>(d, di, N, KS, out); }, (double)N * KS * 4, "gather all rows, random order, unroll 8");
  // pattern B: the rounds' pattern: 20 blocks, each an ascending random 1/20 subset; read block after block
  std::vector<int> blk(N);
  for (int i = 0; i < N; ++i) blk[idx[i]] = std::min(i / (N / 20), 19);
  std::vector<int> order;
  order.reserve(N);
  for (int b = 0; b < 20; ++b)
    for (int i = 0; i < N; ++i)
      if (blk[i] == b) order.push_back(i);
  cudaMemcpy(di, order.data(), N * 4, cudaMemcpyHostToDevice);
  timeit([&] { k_gather<4><<<148 * 16, 256>>>(d, di, N, KS, out); }, (double)N * KS * 4, "gather, 20 ascending subsets, unroll 4");
  timeit([&] { k_gather<8><<<148 * 8, 256>>>(d, di, N, KS, out); }, (double)N * KS * 4, "gather, 20 ascending subsets, unroll 8");
  // one block only (50k rows = 20 MB), cold and L2-warm
  timeit([&] { k_gather<4><<<148 * 4, 256>>>(d, di, N / 20, KS, out); }, (double)N / 20 * KS * 4, "one block (20 MB), repeated -> L2 warm");
  // sorted identity order = sequential rows through the gather kernel
  std::iota(idx.begin(), idx.end(), 0);
  cudaMemcpy(di, idx.data(), N * 4, cudaMemcpyHostToDevice);
  timeit([&] { k_gather<4><<<148 * 16, 256>>>(d, di, N, KS, out); }, (double)N * KS * 4, "gather kernel, identity order");
  return 0;
}
