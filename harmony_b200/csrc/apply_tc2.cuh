// apply_tc2.cuh — EXPERIMENTAL variant of k_apply_tc (apply_tc.cuh), selected with HB_APPLY_V2=1.
// k_apply_tc keeps nothing saturated (ncu, profiles/r01_final_kernels.md: DRAM 19 %, tensor pipe 14 %, issue
// 16 %): with two raw-tile stages only ~1 tile (25 KB) of R is in flight per SM, far below the ~65 KB that the
// HBM latency-bandwidth product asks for.  This variant buys a deeper ring with the shared memory of A:
//   * A = V_q^T is stored with 64 rows per 16-byte K-chunk instead of 128 (d <= 64 of the M = 128 rows are
//     real).  The operand descriptor keeps M = 128 and simply uses LBO = 64 * 16 B, so rows 64..127 of chunk j
//     alias rows 0..63 of chunk j + 1: finite garbage in accumulator lanes 64..127, which nobody reads;
//   * the raw R ring has AP2_NST = 4 stages and the producer runs AP2_NST - 1 tiles ahead.
// Roles, tile shape, 3xTF32 scheme and epilogue are those of k_apply_tc.
#pragma once
#include "apply_tc.cuh"

namespace hb {

constexpr int AP2_NST = 4;    // raw R stages
constexpr int AP2_AROWS = 64; // stored rows of A per K-chunk

__host__ __device__ inline size_t apply_tc2_smem_bytes(int KD, int KS) {
  // A hi/lo: 2 x 64 x KD (+ one extra 64-row chunk each: the last chunk's aliased rows must stay inside the
  // array); B hi/lo: 2 x 64 x KD; raw R: NST x 64 x KS
  return sizeof(float) * (2 * (size_t)AP2_AROWS * (KD + 4) + 2 * (size_t)AP_TN * KD + (size_t)AP2_NST * AP_TN * KS) + 1024;
}

__global__ void __launch_bounds__(AP_THREADS, 1) k_apply_tc2(ApplyTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int KD = a.KD, K = a.K, d = a.d, KS = a.KS, DS = a.DS;
  const size_t a_floats = (size_t)AP2_AROWS * (KD + 4);
  float* Ahi = reinterpret_cast<float*>(smem_raw);  // [KD/4 + 1][64][4]   V_q^T (rows c < d), zero tail chunk
  float* Alo = Ahi + a_floats;
  float* Bhi = Alo + a_floats;                       // [KD/4][64][4]
  float* Blo = Bhi + (size_t)AP_TN * KD;
  float* rawR = Blo + (size_t)AP_TN * KD;            // [NST][64][KS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rawR + (size_t)AP2_NST * AP_TN * KS);
  uint64_t* raw_full = bars + 0;          // [NST]
  uint64_t* raw_empty = bars + AP2_NST;   // [NST]
  uint64_t* b_full = bars + 2 * AP2_NST;
  uint64_t* b_empty = b_full + 1;
  uint64_t* t_full = b_full + 2;          // [2]
  uint64_t* t_empty = b_full + 4;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < AP2_NST; ++i) {
      umma::mbar_init(raw_full + i, 1);
      umma::mbar_init(raw_empty + i, AP_CONV);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(t_full + i, 1);
      umma::mbar_init(t_empty + i, 64);
    }
    umma::mbar_init(b_full, AP_CONV);
    umma::mbar_init(b_empty, 1);
    umma::fence_barrier_init();
  }
  // the zero tail chunk of A (read through the aliased rows of the last real chunk)
  for (int i = tid; i < AP2_AROWS * 4; i += AP_THREADS) {
    Ahi[(size_t)AP2_AROWS * KD + i] = 0.f;
    Alo[(size_t)AP2_AROWS * KD + i] = 0.f;
  }
  if (warp == 2) umma::tmem_alloc(tmem_slot, 128);  // two 64-column accumulators
  umma::fence_proxy_async();
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
      const uint32_t lboA = AP2_AROWS * 16, lboB = AP_TN * 16, sbo = 128;
      const uint32_t aH = umma::smem_u32(Ahi), aL = umma::smem_u32(Alo), bH = umma::smem_u32(Bhi), bL = umma::smem_u32(Blo);
      auto load = [&](int it) {
        const int tile = t_begin + it;
        if (tile >= t_end) return;
        const int s = it % AP2_NST, use = it / AP2_NST;
        if (use >= 1) umma::mbar_wait(raw_empty + s, (use - 1) & 1);
        const uint32_t bytes = (uint32_t)a.tile_len[tile] * KS * 4;
        umma::mbar_arrive_expect_tx(raw_full + s, bytes);
        umma::bulk_load(rawR + (size_t)s * AP_TN * KS, a.R + (size_t)a.tile_cell0[tile] * KS, bytes, raw_full + s);
      };
      for (int i = 0; i < AP2_NST - 1; ++i) load(i);
      int it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        load(it + AP2_NST - 1);
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
      const int s = it % AP2_NST, use = it / AP2_NST;
      umma::mbar_wait(raw_full + s, use & 1);
      if (it >= 1) umma::mbar_wait(b_empty, (it - 1) & 1);  // MMAs of the previous tile are done with A and B
      if (q != cur_q) {
        // new tuple: A = V_q^T (k contiguous per embedding column c), tf32 hi/lo, 64 stored rows
        const float* Vq = a.V + (size_t)q * K * d;
        for (int idx = ct; idx < AP2_AROWS * KD; idx += AP_CONV) {
          const int k = idx / AP2_AROWS, c = idx - k * AP2_AROWS;
          const float v = (k < K && c < d) ? Vq[(size_t)k * d + c] : 0.f;
          float hi, lo;
          umma::split_tf32(v, hi, lo);
          const int off = ((k >> 2) * AP2_AROWS + c) * 4 + (k & 3);
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
