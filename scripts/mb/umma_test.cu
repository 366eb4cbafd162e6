// Stand-alone check of the tcgen05 building blocks in harmony_b200/csrc/umma.cuh: one CTA computes
// D[128 x N] = A[128 x K] * B[N x K]^T with kind::tf32 (1 pass and 3xTF32) and compares with the CPU.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../harmony_b200/csrc/umma.cuh"
using namespace umma;

constexpr int M = 128, N = 112, K = 56;

// canonical no-swizzle K-major tile: [K/4 chunks][rows][4 floats]
__device__ __forceinline__ int tile_off(int row, int k, int rows) { return ((k >> 2) * rows + row) * 4 + (k & 3); }

__global__ void __launch_bounds__(128) k_test(const float* A, const float* B, float* D, int mode3x) {
  extern __shared__ __align__(128) float smem[];
  float* Ahi = smem;
  float* Alo = Ahi + M * K;
  float* Bhi = Alo + M * K;
  float* Blo = Bhi + N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * K; i += 128) {
    int r = i / K, k = i % K;
    float hi, lo;
    split_tf32(A[i], hi, lo);
    Ahi[tile_off(r, k, M)] = mode3x ? hi : A[i];
    Alo[tile_off(r, k, M)] = lo;
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    float hi, lo;
    split_tf32(B[i], hi, lo);
    Bhi[tile_off(r, k, N)] = mode3x ? hi : B[i];
    Blo[tile_off(r, k, N)] = lo;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N, 0, 0);
    const uint32_t lboA = M * 16, lboB = N * 16, sbo = 128;
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t ah = make_desc(smem_u32(Ahi) + ks * 2 * lboA, lboA, sbo);
      uint64_t al = make_desc(smem_u32(Alo) + ks * 2 * lboA, lboA, sbo);
      uint64_t bh = make_desc(smem_u32(Bhi) + ks * 2 * lboB, lboB, sbo);
      uint64_t bl = make_desc(smem_u32(Blo) + ks * 2 * lboB, lboB, sbo);
      if (mode3x) {
        mma_tf32(tmem, al, bh, idesc, acc);  // small terms first
        mma_tf32(tmem, ah, bl, idesc, 1);
        mma_tf32(tmem, ah, bh, idesc, 1);
      } else {
        mma_tf32(tmem, ah, bh, idesc, acc);
      }
      acc = 1;
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[row * N + c + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f);
  for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  size_t smem = sizeof(float) * (2 * M * K + 2 * N * K);
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dD, 0, D.size() * 4);
    k_test<<<1, 128, smem>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
        maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("mode %s: max abs err %.3e (max |ref| %.3f)  D[0]=%f D[last]=%f\n", mode ? "3xTF32" : "1xTF32", maxerr, maxref,
           D[0], D[M * N - 1]);
  }
  return 0;
}
