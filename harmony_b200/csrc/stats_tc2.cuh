// stats_tc2.cuh — EXPERIMENTAL variant of k_stats_tc (stats_tc.cuh), selected with HB_STATS_V2=1: identical roles
// and arithmetic, but ST2_NST = 3 raw-tile stages instead of 2.  k_stats_tc moves 0.6 GB in 0.31 ms (24 % of HBM,
// profiles/r01_final_kernels.md) because only one 39 KB tile is in flight per SM while the other is converted;
// the third stage keeps two in flight, above the latency-bandwidth product of one SM's HBM share.
#pragma once
#include "stats_tc.cuh"

namespace hb {

constexpr int ST2_NST = 3;

__host__ __device__ inline size_t stats_tc2_smem_bytes(int KS, int DS) {
  return sizeof(float) * ((size_t)ST2_NST * ST_TN * (KS + DS) + 2 * 128 * ST_TN + 2 * 64 * ST_TN) + 1024;
}

__global__ void __launch_bounds__(ST_THREADS, 1) k_stats_tc2(StatsTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = a.K, d = a.d, KS = a.KS, DS = a.DS, D1 = d + 1;
  float* rawR = reinterpret_cast<float*>(smem_raw);            // [NST][64][KS]
  float* rawZ = rawR + (size_t)ST2_NST * ST_TN * KS;           // [NST][64][DS]
  float* Ahi = rawZ + (size_t)ST2_NST * ST_TN * DS;              // [16 cell chunks][128 clusters][4 cells]
  float* Alo = Ahi + 128 * ST_TN;
  float* Bhi = Alo + 128 * ST_TN;                          // [16][64 columns][4]
  float* Blo = Bhi + 64 * ST_TN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Blo + 64 * ST_TN);
  uint64_t* raw_full = bars + 0;              // [NST]
  uint64_t* raw_empty = bars + ST2_NST;       // [NST]
  uint64_t* ab_full = bars + 2 * ST2_NST;
  uint64_t* ab_empty = ab_full + 1;
  uint64_t* acc_full = ab_full + 2;
  uint64_t* acc_empty = ab_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ab_full + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < ST2_NST; ++i) {
      umma::mbar_init(raw_full + i, 1);
      umma::mbar_init(raw_empty + i, ST_CONV);
    }
    umma::mbar_init(ab_full, ST_CONV);
    umma::mbar_init(ab_empty, 1);
    umma::mbar_init(acc_full, 1);
    umma::mbar_init(acc_empty, 128);
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc(tmem_slot, 64);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int t_end = (t_begin + a.tiles_per_cta < a.ntiles) ? t_begin + a.tiles_per_cta : a.ntiles;

  if (warp == 0) {
    // =============================== producer ===============================
    if (lane == 0) {
      int it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const int s = it % ST2_NST, use = it / ST2_NST;
        if (use >= 1) umma::mbar_wait(raw_empty + s, (use - 1) & 1);
        const int cell0 = a.tile_cell0[tile], len = a.tile_len[tile];
        const uint32_t bR = (uint32_t)len * KS * 4, bZ = (uint32_t)len * DS * 4;
        umma::mbar_arrive_expect_tx(raw_full + s, bR + bZ);
        umma::bulk_load(rawR + (size_t)s * ST_TN * KS, a.R + (size_t)cell0 * KS, bR, raw_full + s);
        umma::bulk_load(rawZ + (size_t)s * ST_TN * DS, a.Zo + (size_t)cell0 * DS, bZ, raw_full + s);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = umma::make_idesc_tf32(128, 64, 0, 0);
    const uint32_t lboA = 128 * 16, lboB = 64 * 16, sbo = 128;
    const uint32_t aH = umma::smem_u32(Ahi), aL = umma::smem_u32(Alo), bH = umma::smem_u32(Bhi), bL = umma::smem_u32(Blo);
    int it = 0, runs = 0;
    uint32_t accum = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int q = a.tile_tuple[tile];
      const bool last_of_run = (tile + 1 == t_end) || (a.tile_tuple[tile + 1] != q);
      umma::mbar_wait(ab_full, it & 1);
      if (accum == 0 && runs >= 1) umma::mbar_wait(acc_empty, (runs - 1) & 1);  // previous run flushed
      umma::fence_after_sync();
      if (lane == 0) {
        for (int ks = 0; ks < ST_TN / 8; ++ks) {
          const uint64_t ah = umma::make_desc(aH + ks * 2 * lboA, lboA, sbo);
          const uint64_t al = umma::make_desc(aL + ks * 2 * lboA, lboA, sbo);
          const uint64_t bh = umma::make_desc(bH + ks * 2 * lboB, lboB, sbo);
          const uint64_t bl = umma::make_desc(bL + ks * 2 * lboB, lboB, sbo);
          umma::mma_tf32(tmem, al, bh, idesc, accum);
          umma::mma_tf32(tmem, ah, bl, idesc, 1);
          umma::mma_tf32(tmem, ah, bh, idesc, 1);
          accum = 1;
        }
        umma::mma_commit(ab_empty);
        if (last_of_run) umma::mma_commit(acc_full);
      }
      accum = 1;
      __syncwarp();
      if (last_of_run) {
        accum = 0;
        ++runs;
      }
    }
  } else if (warp >= 4) {
    // =============================== converters (+ TMEM flush) ===============================
    const int ct = tid - 128;  // 0..255
    int it = 0, runs = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int len = a.tile_len[tile], q = a.tile_tuple[tile];
      const bool last_of_run = (tile + 1 == t_end) || (a.tile_tuple[tile + 1] != q);
      const int s = it % ST2_NST, use = it / ST2_NST;
      umma::mbar_wait(raw_full + s, use & 1);
      if (it >= 1) umma::mbar_wait(ab_empty, (it - 1) & 1);  // MMAs of the previous tile done with A/B
      const float* rR = rawR + (size_t)s * ST_TN * KS;
      const float* rZ = rawZ + (size_t)s * ST_TN * DS;
      // A[k][cell]: item = (cell chunk cc, cluster k): 4 cells x 1 cluster -> one 16-byte store (hi and lo)
      for (int item = ct; item < 16 * 128; item += ST_CONV) {
        const int cc = item >> 7, k = item & 127;
        float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
        if (k < K) {
          const int c0 = cc * 4;
          const float v0 = (c0 + 0 < len) ? rR[(size_t)(c0 + 0) * KS + k] : 0.f;
          const float v1 = (c0 + 1 < len) ? rR[(size_t)(c0 + 1) * KS + k] : 0.f;
          const float v2 = (c0 + 2 < len) ? rR[(size_t)(c0 + 2) * KS + k] : 0.f;
          const float v3 = (c0 + 3 < len) ? rR[(size_t)(c0 + 3) * KS + k] : 0.f;
          umma::split_tf32(v0, hi.x, lo.x);
          umma::split_tf32(v1, hi.y, lo.y);
          umma::split_tf32(v2, hi.z, lo.z);
          umma::split_tf32(v3, hi.w, lo.w);
        }
        *reinterpret_cast<float4*>(Ahi + (size_t)item * 4) = hi;
        *reinterpret_cast<float4*>(Alo + (size_t)item * 4) = lo;
      }
      // B[c][cell]: item = (cell chunk cc, column c); column d is the ones column
      for (int item = ct; item < 16 * 64; item += ST_CONV) {
        const int cc = item >> 6, c = item & 63;
        float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
        if (c < D1) {
          const int c0 = cc * 4;
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = (c0 + j < len) ? ((c < d) ? rZ[(size_t)(c0 + j) * DS + c] : 1.f) : 0.f;
          umma::split_tf32(v[0], hi.x, lo.x);
          umma::split_tf32(v[1], hi.y, lo.y);
          umma::split_tf32(v[2], hi.z, lo.z);
          umma::split_tf32(v[3], hi.w, lo.w);
        }
        *reinterpret_cast<float4*>(Bhi + (size_t)item * 4) = hi;
        *reinterpret_cast<float4*>(Blo + (size_t)item * 4) = lo;
      }
      umma::fence_proxy_async();
      umma::mbar_arrive(ab_full);
      umma::mbar_arrive(raw_empty + s);
      if (last_of_run && warp < 8) {
        // flush the accumulator of this tuple: thread = cluster row (TMEM lane), 64 columns
        umma::mbar_wait(acc_full, runs & 1);
        umma::fence_after_sync();
        const int wq = warp & 3, k = wq * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);
        for (int c0 = 0; c0 < 64; c0 += 16) {
          float v[16];
          umma::tmem_ld16(trow + c0, v);
          umma::tmem_ld_wait();
          if (k < K) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < D1) atomicAdd(a.S + ((size_t)q * K + k) * D1 + c0 + i, v[i]);
          }
        }
        umma::fence_before_sync();
        umma::mbar_arrive(acc_empty);
      }
      if (last_of_run) ++runs;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 64);
}

}  // namespace hb
