// MN-major tf32 operands with the 128-byte swizzle (the no-swizzle MN-major form returns zeros, see
// umma_mn_test.cu):  D[128 x 64] = sum_kk A[m][kk] * B[n][kk], operands given as row-major [kk][mn] tiles
// (an R tile / a Z tile) and stored as 8(kk) x 32(mn) atoms of 1024 B:
//   addr(mn, kk) = atom(mn/32, kk/8) * 1024 + (kk%8) * 128 + (((mn%32)/4) ^ (kk%8)) * 16 + (mn%4) * 4
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../harmony_b200/csrc/umma.cuh"
using namespace umma;

constexpr int M = 128, N = 64, KK = 64;

__device__ __forceinline__ int sw_off(int mn, int kk, int KKtot) {
  const int atom = (mn >> 5) * (KKtot >> 3) + (kk >> 3);
  return atom * 256 + (kk & 7) * 32 + ((((mn & 31) >> 2) ^ (kk & 7)) << 2) + (mn & 3);
}

__global__ void __launch_bounds__(128) k_test(const float* A, const float* B, float* D, int variant) {
  extern __shared__ __align__(1024) float smem[];
  float* As = smem;
  float* Bs = As + M * KK;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * KK; i += 128) {
    int kk = i / M, m = i % M;
    As[sw_off(m, kk, KK)] = round_tf32(A[i]);
  }
  for (int i = tid; i < N * KK; i += 128) {
    int kk = i / N, n = i % N;
    Bs[sw_off(n, kk, KK)] = round_tf32(B[i]);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N, 1, 1);
    const uint32_t atom = 1024;
    const uint32_t mn_stride = (KK / 8) * atom;  // next 32 rows of M / N
    const uint32_t k_stride = atom;              // next 8 of the reduction
    uint32_t acc = 0;
    for (int ks = 0; ks < KK / 8; ++ks) {
      uint64_t ad, bd;
      if (variant == 0) {  // LBO = MN-atom stride, SBO = K-group stride (CUTLASS comment for MN-major B128)
        ad = make_desc_sw128(smem_u32(As) + ks * k_stride, mn_stride, k_stride);
        bd = make_desc_sw128(smem_u32(Bs) + ks * k_stride, mn_stride, k_stride);
      } else {
        ad = make_desc_sw128(smem_u32(As) + ks * k_stride, k_stride, mn_stride);
        bd = make_desc_sw128(smem_u32(Bs) + ks * k_stride, k_stride, mn_stride);
      }
      mma_tf32(tmem, ad, bd, idesc, acc);
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
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<float> A(M * KK), B(N * KK), D(M * N);
  srand(2);
  for (auto& x : A) x = rand() / (float)RAND_MAX;
  for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  size_t smem = sizeof(float) * (M * KK + N * KK) + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dD, 0, D.size() * 4);
    k_test<<<1, 128, smem>>>(dA, dB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int kk = 0; kk < KK; ++kk) ref += (double)A[kk * M + m] * B[kk * N + n];
        maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("SW128 MN-major variant %d: max abs err %.3e (max |ref| %.3f) D[0]=%f D[1]=%f D[64]=%f\n", variant, maxerr,
           maxref, D[0], D[1], D[64]);
  }
  return 0;
}
