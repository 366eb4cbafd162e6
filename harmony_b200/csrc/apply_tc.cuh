// apply_tc.cuh — K5 on the tensor cores: the correction of harmony.cpp:347 + :615 for all clusters at once,
//   Zc_i = Zo_i - sum_k R_ik V_q[k][:]        (q = covariate tuple of cell i, V_q from k_ridge_solve)
// computed transposed so that the per-tile operand is the small one:
//   D'[c][cell] = sum_k A[c][k] * B[cell][k],   A = V_q^T (static while the tuple does not change, M = 128
//   rows of which d are used),  B = a 64-cell tile of R (N = 64),  3xTF32, fp32 accumulators in TMEM.
// Persistent CTA per SM over a contiguous range of 64-cell tiles:
//   warps 4-7  loader   : R rows -> tf32 hi/lo -> canonical no-swizzle K-major tile, 2 stages; reloads A on a
//                         tuple change
//   warp  2    issuer   : 3 x (K/8) tcgen05.mma per tile, commits to the stage / accumulator mbarriers
//   warps 0-1, 8-9 epilogue : one pair per accumulator stage; TMEM lane = embedding column c:
//                         Zc[cell][c] = Zo[cell][c] - D'[c][cell], coalesced across the warp
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace hb {

constexpr int AP_TN = 64;         // cells per tile (= UMMA N), equals the static tile size TM
constexpr int AP_THREADS = 320;   // warps 0,1 + 8,9 epilogue (TMEM lane quarters 0,1), 2 issuer, 4-7 loader
constexpr int AP_MAXCH = 16;      // float4 chunks of an R row held per loader thread (K <= 128)

struct ApplyTcArgs {
  const float* R;   // [n][KS]
  const float* Zo;  // [n][DS]
  const float* V;   // [J][K][d]
  float* Zc;        // [n][DS]
  const int* tile_cell0;
  const int* tile_len;
  const int* tile_tuple;
  int ntiles, d, K, KS, DS, KD;  // KD = K rounded up to a multiple of 8
  int tiles_per_cta;
  long long* dbg;  // optional [tile][3 roles][4] globaltimer stamps of CTA 0
};

__host__ __device__ inline size_t apply_tc_smem_bytes(int KD) {
  return sizeof(float) * (2 * (size_t)128 * KD + 4 * (size_t)AP_TN * KD) + 256;
}

__global__ void __launch_bounds__(AP_THREADS, 1) k_apply_tc(ApplyTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int KD = a.KD, K = a.K, d = a.d, KS = a.KS, DS = a.DS;
  float* Ahi = reinterpret_cast<float*>(smem_raw);  // [KD/4][128][4]   V_q^T
  float* Alo = Ahi + (size_t)128 * KD;
  float* Bst = Alo + (size_t)128 * KD;               // 2 stages x (hi | lo) x [KD/4][64][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bst + 4 * (size_t)AP_TN * KD);
  uint64_t* full = bars + 0;     // [2]
  uint64_t* empty = bars + 2;    // [2]
  uint64_t* t_full = bars + 4;   // [2]
  uint64_t* t_empty = bars + 6;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(full + i, 128);
      umma::mbar_init(empty + i, 1);
      umma::mbar_init(t_full + i, 1);
      umma::mbar_init(t_empty + i, 64);  // the two warps of the stage's epilogue pair
    }
    umma::fence_barrier_init();
  }
  if (warp == 2) umma::tmem_alloc(tmem_slot, 128);  // two 64-column accumulators
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  auto stamp = [&](int it, int role, int slot) {
    if (a.dbg && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 8 || warp == 2 || warp == 4) && it < 64) {
      long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      a.dbg[((size_t)it * 3 + role) * 4 + slot] = tns;
    }
  };
  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int t_end = (t_begin + a.tiles_per_cta < a.ntiles) ? t_begin + a.tiles_per_cta : a.ntiles;
  const int nch = KD >> 2;            // 16-byte chunks along K
  const int nch0 = (nch + 1) >> 1;    // chunks handled by the first half of the loader threads

  if (warp >= 4 && warp < 8) {
    // =============================== loader ===============================
    const int lt = tid - 128;          // 0..127
    const int cell = lt & 63, half = lt >> 6;
    const int c_lo = half ? nch0 : 0, c_hi = half ? nch : nch0;
    const int KS4 = KS >> 2;
    int cur_q = -1;
    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int cell0 = a.tile_cell0[tile], len = a.tile_len[tile], q = a.tile_tuple[tile];
      const int s = it & 1, use = it >> 1;
      stamp(it, 0, 0);
      float4 rv[AP_MAXCH];
      const bool live = cell < len;
      const float4* rp = reinterpret_cast<const float4*>(a.R + (size_t)(cell0 + cell) * KS);
#pragma unroll
      for (int j = 0; j < AP_MAXCH; ++j) {
        const int c4 = c_lo + j;
        rv[j] = (live && c4 < c_hi && c4 < KS4) ? ld_stream4(rp + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      stamp(it, 0, 1);
      if (use >= 1) umma::mbar_wait(empty + s, (use - 1) & 1);  // stage s free (MMAs of tile it-2 done)
      stamp(it, 0, 2);
      if (q != cur_q) {
        // new tuple: A = V_q^T.  Every earlier MMA must have completed (they read A).
        if (it >= 1) umma::mbar_wait(empty + (s ^ 1), ((it - 1) >> 1) & 1);
        const float* Vq = a.V + (size_t)q * K * d;
        for (int idx = lt; idx < 128 * KD; idx += 128) {
          const int k = idx / 128, c = idx - k * 128;  // lanes along c: conflict-light 4-byte stores
          const float v = (k < K && c < d) ? Vq[(size_t)k * d + c] : 0.f;
          float hi, lo;
          umma::split_tf32(v, hi, lo);
          const int off = ((k >> 2) * 128 + c) * 4 + (k & 3);
          Ahi[off] = hi;
          Alo[off] = lo;
        }
        cur_q = q;
        asm volatile("bar.sync 1, 128;" ::: "memory");  // all loader threads wrote their part of A
      }
      float* Bhi = Bst + (size_t)s * 2 * AP_TN * KD;
      float* Blo = Bhi + (size_t)AP_TN * KD;
#pragma unroll
      for (int j = 0; j < AP_MAXCH; ++j) {
        const int c4 = c_lo + j;
        if (c4 < c_hi) {
          float4 hi, lo;
          umma::split_tf32(rv[j].x, hi.x, lo.x);
          umma::split_tf32(rv[j].y, hi.y, lo.y);
          umma::split_tf32(rv[j].z, hi.z, lo.z);
          umma::split_tf32(rv[j].w, hi.w, lo.w);
          *reinterpret_cast<float4*>(Bhi + ((size_t)c4 * AP_TN + cell) * 4) = hi;
          *reinterpret_cast<float4*>(Blo + ((size_t)c4 * AP_TN + cell) * 4) = lo;
        }
      }
      umma::fence_proxy_async();
      umma::mbar_arrive(full + s);
      stamp(it, 0, 3);
    }
  } else if (warp == 2) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = umma::make_idesc_tf32(128, AP_TN, 0, 0);
    const uint32_t lboA = 128 * 16, lboB = AP_TN * 16, sbo = 128;
    const uint32_t aH = umma::smem_u32(Ahi), aL = umma::smem_u32(Alo);
    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int s = it & 1, use = it >> 1;
      stamp(it, 1, 0);
      umma::mbar_wait(full + s, use & 1);
      stamp(it, 1, 1);
      if (use >= 1) umma::mbar_wait(t_empty + s, (use - 1) & 1);
      umma::fence_after_sync();
      stamp(it, 1, 2);
      if (lane == 0) {
        const uint32_t bH = umma::smem_u32(Bst + (size_t)s * 2 * AP_TN * KD);
        const uint32_t bL = bH + (uint32_t)(AP_TN * KD * 4);
        const uint32_t dt = tmem + s * AP_TN;
        uint32_t accum = 0;
        for (int ks = 0; ks < KD / 8; ++ks) {
          const uint64_t ah = umma::make_desc(aH + ks * 2 * lboA, lboA, sbo);
          const uint64_t al = umma::make_desc(aL + ks * 2 * lboA, lboA, sbo);
          const uint64_t bh = umma::make_desc(bH + ks * 2 * lboB, lboB, sbo);
          const uint64_t bl = umma::make_desc(bL + ks * 2 * lboB, lboB, sbo);
          umma::mma_tf32(dt, al, bh, idesc, accum);
          umma::mma_tf32(dt, ah, bl, idesc, 1);
          umma::mma_tf32(dt, ah, bh, idesc, 1);
          accum = 1;
        }
        umma::mma_commit(empty + s);
        umma::mma_commit(t_full + s);
      }
      __syncwarp();
      stamp(it, 1, 3);
    }
  } else if (warp < 2 || warp >= 8) {
    // =============================== epilogue (one warp pair per accumulator stage) ===============================
    const int es = (warp >= 8) ? 1 : 0;      // stage served by this pair
    const int wq = warp & 3;                 // TMEM lane quarter: 0 or 1
    const int c = wq * 32 + lane;            // embedding column = TMEM lane
    int it = es, use = 0;
    for (int tile = t_begin + es; tile < t_end; tile += 2, it += 2, ++use) {
      const int cell0 = a.tile_cell0[tile], len = a.tile_len[tile];
      // the tile's Zo values do not depend on the MMA: fetch them while it runs
      float zo[AP_TN];
      if (c < d) {
#pragma unroll
        for (int j = 0; j < AP_TN; ++j) zo[j] = (j < len) ? ld_stream(a.Zo + (size_t)(cell0 + j) * DS + c) : 0.f;
      }
      stamp(it, 2, 0);
      umma::mbar_wait(t_full + es, use & 1);
      umma::fence_after_sync();
      stamp(it, 2, 1);
      const uint32_t trow = tmem + es * AP_TN + ((uint32_t)(wq * 32) << 16);
      float v[AP_TN];
#pragma unroll
      for (int j = 0; j < AP_TN; j += 16) {
        float t16[16];
        umma::tmem_ld16(trow + j, t16);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[j + i] = t16[i];
      }
      umma::tmem_ld_wait();
      umma::fence_before_sync();
      umma::mbar_arrive(t_empty + es);
      stamp(it, 2, 2);
      if (c < d) {
#pragma unroll
        for (int j = 0; j < AP_TN; ++j)
          if (j < len) a.Zc[(size_t)(cell0 + j) * DS + c] = zo[j] - v[j];
      }
      stamp(it, 2, 3);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 2) umma::tmem_dealloc(tmem, 128);
}

}  // namespace hb
