// Address-map probe for tcgen05.mma shared-memory descriptors (kind::tf32): which shared-memory word does the
// tensor core read for operand element (mn, kk) under a given (major, layout_type, LBO, SBO)?
// One K = 8 instruction; the OTHER operand is a K-major selector (row r has a single 1 at kk = r, r < 8), so
//   D[m][n] = probed(mn, kk)   with kk = the selector's row index
// and the probed operand's 64 KB region holds, word by word, its own index + 1 (pass 0: low 11 bits, pass 1: the bits
// above — both exact in tf32).  The printed table is the byte offset read for every (kk, mn).
//   usage: umma_layout_probe <which: 0 = B probed (N = 64), 1 = A probed (M = 128)> <mn_major 0|1>
//                            <layout_type 0 none | 2 128B | 4 64B | 6 32B> <LBO bytes> <SBO bytes>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../harmony_b200/csrc/umma.cuh"
using namespace umma;

constexpr int M = 128, N = 64, WORDS = 16384;

__global__ void __launch_bounds__(128) k_probe(float* D, int which, int mn_major, int layout, int lbo, int sbo, int pass) {
  extern __shared__ float smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  float* region = (float*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // 1024-byte aligned for the swizzles
  float* sel = region + WORDS;                                                 // K-major selector, [kk/4][row][4]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int SR = which ? N : M;  // rows of the selector operand
  for (int i = tid; i < WORDS; i += 128) region[i] = (float)(pass ? ((i + 1) >> 11) : ((i + 1) & 2047));  // 0 = "read nothing"
  for (int i = tid; i < SR * 8; i += 128) {
    const int kk = i / SR, r = i % SR;
    sel[((kk >> 2) * SR + r) * 4 + (kk & 3)] = (r == kk) ? 1.f : 0.f;
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
    const uint32_t idesc = make_idesc_tf32(M, N, which ? mn_major : 0, which ? 0 : mn_major);
    const uint64_t pd = make_desc(smem_u32(region), (uint32_t)lbo, (uint32_t)sbo) | ((uint64_t)layout << 61);
    const uint64_t sd = make_desc(smem_u32(sel), SR * 16, 128);
    mma_tf32(tmem, which ? pd : sd, which ? sd : pd, idesc, 0);
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

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const int which = atoi(argv[1]), mn_major = atoi(argv[2]), layout = atoi(argv[3]), lbo = atoi(argv[4]), sbo = atoi(argv[5]);
  std::vector<float> D0(M * N), D1(M * N);
  float* dD;
  cudaMalloc(&dD, M * N * 4);
  const size_t smem = sizeof(float) * (WORDS + 128 * 8) + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int pass = 0; pass < 2; ++pass) {
    cudaMemset(dD, 0xff, M * N * 4);
    k_probe<<<1, 128, smem>>>(dD, which, mn_major, layout, lbo, sbo, pass);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("probe %s %s-major layout %d lbo %d sbo %d: CUDA error %s\n", which ? "A" : "B", mn_major ? "MN" : "K", layout, lbo,
             sbo, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy((pass ? D1 : D0).data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  }
  const int MN = which ? M : N;
  printf("probe %s (%d rows) %s-major layout_type %d LBO %d SBO %d: byte offset read for (kk, mn)\n", which ? "A" : "B", MN,
         mn_major ? "MN" : "K", layout, lbo, sbo);
  const int cols[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 28, 31, 32, 33, 36, 63, 64, 65, 96, 127};
  printf("   mn:");
  for (int c : cols)
    if (c < MN) printf(" %6d", c);
  printf("\n");
  for (int kk = 0; kk < 8; ++kk) {
    printf("kk %d:", kk);
    for (int c : cols) {
      if (c >= MN) continue;
      // D[m][n]: which = 0 -> m = kk (selector row), n = mn;  which = 1 -> m = mn, n = kk
      const int idx = which ? c * N + kk : kk * N + c;
      const float lo = D0[idx], hi = D1[idx];
      if (!(lo >= 0.f && lo < 2048.f && hi >= 0.f && hi <= 8.f) || lo != floorf(lo) || hi != floorf(hi))
        printf(" %6s", "?");
      else if (lo == 0.f && hi == 0.f)
        printf(" %6s", "zero");
      else
        printf(" %6d", 4 * ((((int)hi << 11) + (int)lo) - 1));
    }
    printf("\n");
  }
  return 0;
}
