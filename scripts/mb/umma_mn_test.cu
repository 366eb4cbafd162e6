// Stand-alone check of MN-major (transposed) tf32 operands vs K-major, in all four combinations:
//   D[128 x 64] = sum_kk A[m][kk] * B[n][kk]
// K-major canonical no-swizzle tile : [kk/4][mn][kk%4]          (LBO = MN*16 B, SBO = 128 B)
// MN-major canonical no-swizzle tile: [kk/8][mn/4][kk%8][mn%4]  (SBO = 128 B between mn-groups, LBO = (MN/4)*128 B)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../harmony_b200/csrc/umma.cuh"
using namespace umma;

constexpr int M = 128, N = 64, KK = 64;

__device__ __forceinline__ int mn_off(int mn, int kk, int MN) { return (((kk >> 3) * (MN >> 2) + (mn >> 2)) * 8 + (kk & 7)) * 4 + (mn & 3); }
__device__ __forceinline__ int k_off(int mn, int kk, int MN) { return ((kk >> 2) * MN + mn) * 4 + (kk & 3); }

__global__ void __launch_bounds__(128) k_test(const float* A, const float* B, float* D, int a_mn, int b_mn, int swap) {
  extern __shared__ __align__(128) float smem[];
  float* As = smem;
  float* Bs = As + M * KK;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * KK; i += 128) {
    int kk = i / M, m = i % M;
    As[a_mn ? mn_off(m, kk, M) : k_off(m, kk, M)] = round_tf32(A[i]);
  }
  for (int i = tid; i < N * KK; i += 128) {
    int kk = i / N, n = i % N;
    Bs[b_mn ? mn_off(n, kk, N) : k_off(n, kk, N)] = round_tf32(B[i]);
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
    const uint32_t idesc = make_idesc_tf32(M, N, a_mn, b_mn);
    uint32_t acc = 0;
    for (int ks = 0; ks < KK / 8; ++ks) {
      uint64_t ad, bd;
      if (a_mn) {
        uint32_t lbo = (M / 4) * 128, sbo = 128;
        ad = swap ? make_desc(smem_u32(As) + ks * lbo, sbo, lbo) : make_desc(smem_u32(As) + ks * lbo, lbo, sbo);
      } else {
        ad = make_desc(smem_u32(As) + ks * 2 * M * 16, M * 16, 128);
      }
      if (b_mn) {
        uint32_t lbo = (N / 4) * 128, sbo = 128;
        bd = swap ? make_desc(smem_u32(Bs) + ks * lbo, sbo, lbo) : make_desc(smem_u32(Bs) + ks * lbo, lbo, sbo);
      } else {
        bd = make_desc(smem_u32(Bs) + ks * 2 * N * 16, N * 16, 128);
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
  size_t smem = sizeof(float) * (M * KK + N * KK);
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int swap = 0; swap < 2; ++swap)
    for (int a_mn = 0; a_mn < 2; ++a_mn)
      for (int b_mn = 0; b_mn < 2; ++b_mn) {
        if (swap && !a_mn && !b_mn) continue;
        cudaMemset(dD, 0, D.size() * 4);
        k_test<<<1, 128, smem>>>(dA, dB, dD, a_mn, b_mn, swap);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("a_mn=%d b_mn=%d swap=%d: CUDA error %s\n", a_mn, b_mn, swap, cudaGetErrorString(e));
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
        printf("A %s-major, B %s-major, lbo/sbo %s: max abs err %.3e (max |ref| %.3f) D[0]=%f D[1]=%f D[64]=%f\n",
               a_mn ? "MN" : "K", b_mn ? "MN" : "K", swap ? "swapped" : "as-doc", maxerr, maxref, D[0], D[1], D[64]);
      }
  return 0;
}
