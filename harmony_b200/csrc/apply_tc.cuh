// apply_tc.cuh — K5 on the tensor cores: the correction of harmony.cpp:347 + :615 for all clusters at once,
//   Zc_i = Zo_i - sum_k R_ik V_q[k][:]        (q = covariate tuple of cell i, V_q from k_ridge_solve)
// computed transposed so that the per-tile operand is the small one:
//   D'[c][cell] = sum_k A[c][k] * B[cell][k],   A = V_q^T (static while the tuple does not change, M = 128
//   rows of which d are used),  B = a 64-cell tile of R (N = 64),  3xTF32, fp32 accumulators in TMEM.
// Persistent CTA per SM over a contiguous range of 64-cell tiles:
//   warp 2       producer + issuer : 1-D bulk (TMA) loads of the raw R tiles (2 stages), then per tile
//                                    3 x (K/8) tcgen05.mma, commits to the operand / accumulator mbarriers
//   warps 4-11   converters        : raw tile -> tf32 hi/lo -> canonical no-swizzle K-major B tile; reload A
//                                    (= V_q^T) when the tuple changes
//   warps 0-1, 12-13 epilogue      : one pair per accumulator; TMEM lane = embedding column c:
//                                    Zc[cell][c] = Zo[cell][c] - D'[c][cell], coalesced across the warp
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace hb {

constexpr int AP_TN = 64;         // cells per tile (= UMMA N), equals the static tile size TM
constexpr int AP_THREADS = 448;   // 14 warps
constexpr int AP_CONV = 256;      // converter threads (warps 4-11)

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
  long long* dbg;
};

__host__ __device__ inline size_t apply_tc_smem_bytes(int KD, int KS) {
  // A hi/lo: 2 x 128 x KD; B hi/lo (single stage): 2 x 64 x KD; raw R stages: 2 x 64 x KS
  return sizeof(float) * (2 * (size_t)128 * KD + 2 * (size_t)AP_TN * KD + 2 * (size_t)AP_TN * KS) + 1024;
}

__global__ void __launch_bounds__(AP_THREADS, 1) k_apply_tc(ApplyTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int KD = a.KD, K = a.K, d = a.d, KS = a.KS, DS = a.DS;
  float* Ahi = reinterpret_cast<float*>(smem_raw);  // [KD/4][128][4]   V_q^T
  float* Alo = Ahi + (size_t)128 * KD;
  float* Bhi = Alo + (size_t)128 * KD;               // [KD/4][64][4]
  float* Blo = Bhi + (size_t)AP_TN * KD;
  float* rawR = Blo + (size_t)AP_TN * KD;            // [2][64][KS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rawR + 2 * (size_t)AP_TN * KS);
  uint64_t* raw_full = bars + 0;   // [2]
  uint64_t* raw_empty = bars + 2;  // [2]
  uint64_t* b_full = bars + 4;
  uint64_t* b_empty = bars + 5;
  uint64_t* t_full = bars + 6;     // [2]
  uint64_t* t_empty = bars + 8;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(raw_full + i, 1);
      umma::mbar_init(raw_empty + i, AP_CONV);
      umma::mbar_init(t_full + i, 1);
      umma::mbar_init(t_empty + i, 64);
    }
    umma::mbar_init(b_full, AP_CONV);
    umma::mbar_init(b_empty, 1);
    umma::fence_barrier_init();
  }
  if (warp == 2) umma::tmem_alloc(tmem_slot, 128);  // two 64-column accumulators
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int t_end = (t_begin + a.tiles_per_cta < a.ntiles) ? t_begin + a.tiles_per_cta : a.ntiles;
  const int nch = KD >> 2;  // 16-byte chunks along K

  if (warp == 2) {
    // =============================== producer + MMA issuer (one thread) ===============================
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_tf32(128, AP_TN, 0, 0);
      const uint32_t lboA = 128 * 16, lboB = AP_TN * 16, sbo = 128;
      const uint32_t aH = umma::smem_u32(Ahi), aL = umma::smem_u32(Alo), bH = umma::smem_u32(Bhi), bL = umma::smem_u32(Blo);
      auto load = [&](int it) {
        const int tile = t_begin + it;
        if (tile >= t_end) return;
        const int s = it & 1, use = it >> 1;
        if (use >= 1) umma::mbar_wait(raw_empty + s, (use - 1) & 1);
        const uint32_t bytes = (uint32_t)a.tile_len[tile] * KS * 4;
        umma::mbar_arrive_expect_tx(raw_full + s, bytes);
        umma::bulk_load(rawR + (size_t)s * AP_TN * KS, a.R + (size_t)a.tile_cell0[tile] * KS, bytes, raw_full + s);
      };
      load(0);
      int it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        load(it + 1);
        const int acc = it & 1, use = it >> 1;
        umma::mbar_wait(b_full, it & 1);
        if (use >= 1) umma::mbar_wait(t_empty + acc, (use - 1) & 1);
        umma::fence_after_sync();
        const uint32_t dt = tmem + acc * AP_TN;
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
        umma::mma_commit(b_empty);
        umma::mma_commit(t_full + acc);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // =============================== converters ===============================
    const int ct = tid - 128;  // 0..255
    const int KS4 = KS >> 2;
    int cur_q = -1;
    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int len = a.tile_len[tile], q = a.tile_tuple[tile];
      const int s = it & 1, use = it >> 1;
      umma::mbar_wait(raw_full + s, use & 1);
      if (it >= 1) umma::mbar_wait(b_empty, (it - 1) & 1);  // MMAs of the previous tile are done with A and B
      if (q != cur_q) {
        // new tuple: A = V_q^T (k contiguous per embedding column c), tf32 hi/lo
        const float* Vq = a.V + (size_t)q * K * d;
        for (int idx = ct; idx < 128 * KD; idx += AP_CONV) {
          const int k = idx / 128, c = idx - k * 128;
          const float v = (k < K && c < d) ? Vq[(size_t)k * d + c] : 0.f;
          float hi, lo;
          umma::split_tf32(v, hi, lo);
          const int off = ((k >> 2) * 128 + c) * 4 + (k & 3);
          Ahi[off] = hi;
          Alo[off] = lo;
        }
        cur_q = q;
      }
      const float* rR = rawR + (size_t)s * AP_TN * KS;
      for (int item = ct; item < nch * AP_TN; item += AP_CONV) {
        const int c4 = item / AP_TN, cell = item - c4 * AP_TN;  // lanes along cells: conflict-free 16-byte stores
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cell < len && c4 < KS4) v = *reinterpret_cast<const float4*>(rR + (size_t)cell * KS + c4 * 4);
        float4 hi, lo;
        umma::split_tf32(v.x, hi.x, lo.x);
        umma::split_tf32(v.y, hi.y, lo.y);
        umma::split_tf32(v.z, hi.z, lo.z);
        umma::split_tf32(v.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(Bhi + (size_t)item * 4) = hi;
        *reinterpret_cast<float4*>(Blo + (size_t)item * 4) = lo;
      }
      umma::fence_proxy_async();
      umma::mbar_arrive(b_full);
      umma::mbar_arrive(raw_empty + s);
    }
  } else if (warp < 2 || warp >= 12) {
    // =============================== epilogue (one warp pair per accumulator) ===============================
    const int es = (warp >= 12) ? 1 : 0;     // accumulator served by this pair
    const int wq = warp & 3;                 // TMEM lane quarter: 0 or 1
    const int c = wq * 32 + lane;            // embedding column = TMEM lane
    int use = 0;
    for (int tile = t_begin + es; tile < t_end; tile += 2, ++use) {
      const int cell0 = a.tile_cell0[tile], len = a.tile_len[tile];
      // the tile's Zo values do not depend on the MMA: fetch them while it runs
      float zo[AP_TN];
      if (c < d) {
#pragma unroll
        for (int j = 0; j < AP_TN; ++j) zo[j] = (j < len) ? ld_stream(a.Zo + (size_t)(cell0 + j) * DS + c) : 0.f;
      }
      umma::mbar_wait(t_full + es, use & 1);
      umma::fence_after_sync();
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
      if (c < d) {
#pragma unroll
        for (int j = 0; j < AP_TN; ++j)
          if (j < len) a.Zc[(size_t)(cell0 + j) * DS + c] = zo[j] - v[j];
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 2) umma::tmem_dealloc(tmem, 128);
}

}  // namespace hb
