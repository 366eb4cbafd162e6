// assign_tc2.cuh — EXPERIMENTAL rewrite of K1 (assign_tc.cuh), selected with HB_ASSIGN_V2=1.
//
// Same arithmetic as k_assign_tc (3xTF32 contraction on tcgen05, dist -> U -> softmax, column sums per level,
// objective partials; harmony.cpp:141-150 / :220-227) — different plumbing.  The stamp trace and the ncu source
// page of k_assign_tc show one serial chain per 128-cell tile as the critical path: the
// loader warps fetch the tile metadata, then their rows with per-thread global loads (latency exposed twice),
// split them while sharing a scheduler with two epilogue warps, wait for the single operand buffer, write the
// normalised rows back with per-thread stores and only then issue the MMAs; the two epilogue groups idle half
// of the time, 12 warps per SM.  Here
//   warp 0 (one lane)    producer : one 1-D bulk (TMA) load per raw Z tile, A2_NR stages ahead
//   warp 1 (one lane)    issuer   : 3 x (d/8) tcgen05.mma per tile into one of A2_NG TMEM accumulators
//   warps 2-5            convert  : own row from the staged raw tile -> L2-normalise -> tf32 hi/lo -> one of A2_NA
//                                   operand stages; the normalised tile leaves through one bulk store
//   warps 6-17           epilogue : A2_NG groups of 4 warps, TMEM lane (= cell) per thread, two sweeps over the
//                                   accumulator: (1) U rows + row sums / objective pieces, (2) R rows + column
//                                   sums by a butterfly reduction across the warp's 32 rows (no staged R tile:
//                                   its 2 x 51 KB of shared memory now hold the raw ring and the second operand
//                                   stage)
// Limits: clusters padded to NP <= 128 (three accumulators in 512 TMEM columns), d <= 64.
#pragma once
#include "assign_tc.cuh"

namespace hb {

constexpr int A2_THREADS = 576;  // 18 warps
constexpr int A2_NR = 2;         // raw Z stages
constexpr int A2_NA = 2;         // operand (A hi/lo) stages
constexpr int A2_NG = 3;         // epilogue groups = TMEM accumulators
constexpr int A2_CONV = 128;     // converter threads (warps 2-5)

__host__ __device__ inline size_t assign_tc2_smem_bytes(int KD, int NP, int DS) {
  // raw ring, A stages (hi + lo), B hi/lo, sigma / -1/sigma, column-sum scratch per group
  return sizeof(float) * ((size_t)A2_NR * TC_TM * DS + (size_t)A2_NA * 2 * TC_TM * KD + 2 * (size_t)NP * KD + 2 * (size_t)NP +
                          (size_t)A2_NG * NP) + 256;
}

__global__ void __launch_bounds__(A2_THREADS, 1) k_assign_tc2(AssignTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int KD = a.KD, NP = a.NP, KS = a.KS, K = a.K, d = a.d, DS = a.DS;
  float* raw = reinterpret_cast<float*>(smem_raw);               // [NR][128][DS]
  float* Aop = raw + (size_t)A2_NR * TC_TM * DS;                 // [NA][hi | lo][KD/4][128][4]
  float* Bhi = Aop + (size_t)A2_NA * 2 * TC_TM * KD;             // [KD/4][NP][4]
  float* Blo = Bhi + (size_t)NP * KD;
  float* sig = Blo + (size_t)NP * KD;                            // [NP]
  float* isig = sig + NP;                                        // [NP] -1/sigma
  float* colsum = isig + NP;                                     // [NG][NP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(colsum + (size_t)A2_NG * NP);
  uint64_t* raw_full = bars;                      // [NR]  producer -> converters (tx bytes)
  uint64_t* raw_empty = raw_full + A2_NR;         // [NR]  converters (128) -> producer
  uint64_t* a_full = raw_empty + A2_NR;           // [NA]  converters (128) -> issuer
  uint64_t* a_empty = a_full + A2_NA;             // [NA]  tcgen05.commit -> converters
  uint64_t* t_full = a_empty + A2_NA;             // [NG]  tcgen05.commit -> epilogue group
  uint64_t* t_empty = t_full + A2_NG;             // [NG]  epilogue group (128) -> issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + A2_NG);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: centroids -> tf32 hi/lo operand tiles, tables, barriers, TMEM ----
  for (int i = tid; i < NP * KD; i += A2_THREADS) {
    const int n = i / KD, k = i - n * KD;
    const float y = (n < K && k < d) ? a.Y[(size_t)n * d + k] : 0.f;
    float hi, lo;
    umma::split_tf32(y, hi, lo);
    const int off = ((k >> 2) * NP + n) * 4 + (k & 3);
    Bhi[off] = hi;
    Blo[off] = lo;
  }
  for (int k = tid; k < NP; k += A2_THREADS) {
    sig[k] = (k < K) ? a.sigma[k] : 1.f;
    isig[k] = (k < K) ? -1.f / a.sigma[k] : -5.0e29f;  // padding columns: dist = 2 -> u = U_PAD -> exp = 0
  }
  for (int k = tid; k < A2_NG * NP; k += A2_THREADS) colsum[k] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < A2_NR; ++i) {
      umma::mbar_init(raw_full + i, 1);
      umma::mbar_init(raw_empty + i, A2_CONV);
    }
    for (int i = 0; i < A2_NA; ++i) {
      umma::mbar_init(a_full + i, A2_CONV);
      umma::mbar_init(a_empty + i, 1);
    }
    for (int i = 0; i < A2_NG; ++i) {
      umma::mbar_init(t_full + i, 1);
      umma::mbar_init(t_empty + i, 128);
    }
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc(tmem_slot, 512);  // three 128-column accumulators (power-of-two allocation)
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const int my_first = blockIdx.x, stride = gridDim.x;
  const uint32_t row_bytes = (uint32_t)DS * 4u;

  if (warp == 0) {
    // =============================== producer ===============================
    if (lane == 0) {
      int it = 0;
      for (int tile = my_first; tile < a.ntiles; tile += stride, ++it) {
        const int s = it % A2_NR, use = it / A2_NR;
        const int len = a.tile_len[tile], cell0 = a.tile_cell0[tile];  // in flight while the stage drains
        if (use >= 1) umma::mbar_wait(raw_empty + s, (use - 1) & 1);
        const uint32_t bytes = (uint32_t)len * row_bytes;
        umma::mbar_arrive_expect_tx(raw_full + s, bytes);
        umma::bulk_load(raw + (size_t)s * TC_TM * DS, a.Zc + (size_t)cell0 * DS, bytes, raw_full + s);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_tf32(TC_TM, NP, 0, 0);
      const uint32_t lboA = TC_TM * 16, lboB = NP * 16, sbo = 128;
      const uint32_t bH = umma::smem_u32(Bhi), bL = umma::smem_u32(Blo);
      int it = 0;
      for (int tile = my_first; tile < a.ntiles; tile += stride, ++it) {
        const int sa = it % A2_NA, acc = it % A2_NG;
        umma::mbar_wait(a_full + sa, (it / A2_NA) & 1);
        if (it >= A2_NG) umma::mbar_wait(t_empty + acc, ((it / A2_NG) - 1) & 1);
        umma::fence_after_sync();
        const uint32_t aH = umma::smem_u32(Aop + (size_t)sa * 2 * TC_TM * KD);
        const uint32_t aL = aH + (uint32_t)TC_TM * KD * 4u;
        const uint32_t dt = tmem + acc * 128;
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
        umma::mma_commit(a_empty + sa);   // operands consumed -> the converters may refill the stage
        umma::mma_commit(t_full + acc);   // accumulator ready -> epilogue group `acc`
      }
    }
  } else if (warp < 6) {
    // =============================== converters ===============================
    const int r = tid - 64;  // row of the tile, 0..127
    const int DS4 = DS >> 2, KD4 = KD >> 2;
    int it = 0;
    for (int tile = my_first; tile < a.ntiles; tile += stride, ++it) {
      const int cell0 = a.tile_cell0[tile], len = a.tile_len[tile];
      const int sr = it % A2_NR, sa = it % A2_NA;
      float* rawt = raw + (size_t)sr * TC_TM * DS;
      float* rowp = rawt + (size_t)r * DS;
      umma::mbar_wait(raw_full + sr, (it / A2_NR) & 1);
      float4 z[TC_DS4MAX];
#pragma unroll
      for (int c = 0; c < TC_DS4MAX; ++c)
        z[c] = (c < DS4 && r < len) ? *reinterpret_cast<const float4*>(rowp + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const bool wb = (r < len) && a.normalise;
      if (wb) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < TC_DS4MAX; ++c) ss += (z[c].x * z[c].x + z[c].y * z[c].y) + (z[c].z * z[c].z + z[c].w * z[c].w);
        float nrm = sqrtf(ss);
        if (nrm == 0.f) nrm = 1.f;
        const float rn = 1.f / nrm;
#pragma unroll
        for (int c = 0; c < TC_DS4MAX; ++c) {
          z[c].x *= rn;
          z[c].y *= rn;
          z[c].z *= rn;
          z[c].w *= rn;
        }
      }
      if (!a.normalise) umma::mbar_arrive(raw_empty + sr);  // the row lives in registers now
      // the operand stage is free once the MMAs of tile it - NA have completed
      if (it >= A2_NA) umma::mbar_wait(a_empty + sa, ((it / A2_NA) - 1) & 1);
      float* Ahi = Aop + (size_t)sa * 2 * TC_TM * KD;
      float* Alo = Ahi + (size_t)TC_TM * KD;
#pragma unroll
      for (int c = 0; c < TC_DS4MAX + 1; ++c)
        if (c < KD4) {
          float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
          if (c < TC_DS4MAX && c < DS4) {
            umma::split_tf32(z[c].x, hi.x, lo.x);
            umma::split_tf32(z[c].y, hi.y, lo.y);
            umma::split_tf32(z[c].z, hi.z, lo.z);
            umma::split_tf32(z[c].w, hi.w, lo.w);
          }
          *reinterpret_cast<float4*>(Ahi + ((size_t)c * TC_TM + r) * 4) = hi;
          *reinterpret_cast<float4*>(Alo + ((size_t)c * TC_TM + r) * 4) = lo;
        }
      umma::fence_proxy_async();
      umma::mbar_arrive(a_full + sa);
      if (a.normalise) {
        // write the normalised rows back (harmony.cpp:220): in place in the raw stage, then one bulk store
        if (wb) {
#pragma unroll
          for (int c = 0; c < TC_DS4MAX; ++c)
            if (c < DS4) *reinterpret_cast<float4*>(rowp + c * 4) = z[c];
        }
        umma::fence_proxy_async();
        umma::named_sync(1, A2_CONV);
        if (r == 0) {
          umma::bulk_store(a.Zc + (size_t)cell0 * DS, rawt, (uint32_t)len * row_bytes);
          umma::bulk_commit();
          umma::bulk_wait_read();  // the stage may be refilled once the store has read it
        }
        umma::mbar_arrive(raw_empty + sr);
      }
    }
    if (r == 0) umma::bulk_wait_all();
  } else {
    // =============================== epilogue ===============================
    const int g = (warp - 6) >> 2;     // group = TMEM accumulator
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // row of the tile = TMEM lane
    const int et = ((warp - 6) & 3) * 32 + lane;  // 0..127 within the group
    const int bar_id = 2 + g;
    float* cs = colsum + (size_t)g * NP;
    // column owned by this lane after the butterfly (bit i of the column = bit i + 1 of the lane)
    const int bcol = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    float okd = 0.f, oent = 0.f;
    int use = 0;
    for (int it = g; my_first + it * stride < a.ntiles; it += A2_NG, ++use) {
      const int tile = my_first + it * stride;
      const int cell0 = a.tile_cell0[tile], len = a.tile_len[tile], tq = a.tile_tuple[tile];
      umma::mbar_wait(t_full + g, use & 1);
      umma::fence_after_sync();
      const uint32_t trow = tmem + g * 128 + ((uint32_t)(q * 32) << 16);
      const bool live = r < len;
      float* ug = a.U + (size_t)(cell0 + r) * KS;
      float* rg = a.R + (size_t)(cell0 + r) * KS;
      // ---- sweep 1: dist -> u (stored), row sums and objective pieces ----
      float ssum = 0.f, A1 = 0.f, B1 = 0.f, S1 = 0.f;  // sum e, sum e*dist, sum sigma*e*u, sum sigma*e
      for (int c = 0; c < NP; c += 16) {
        float v[16];
        umma::tmem_ld16(trow + c, v);
        umma::tmem_ld_wait();
        float uu[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float dist = 2.f * (1.f - v[i]);
          uu[i] = dist * isig[c + i];  // -dist / sigma (U_PAD in the padding columns)
          const float e = __expf(uu[i]);
          ssum += e;
          A1 = fmaf(e, dist, A1);
          const float se = sig[c + i] * e;
          B1 = fmaf(se, uu[i], B1);
          S1 += se;
        }
        if (live) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            if (c + i < KS) *reinterpret_cast<float4*>(ug + c + i) = make_float4(uu[i], uu[i + 1], uu[i + 2], uu[i + 3]);
        }
      }
      // R.each_row() /= sum(R, 0) (no zero guard in the reference); rows beyond the tile weigh 0
      const float inv = live ? 1.f / ssum : 0.f;
      if (live) {
        const float ls = __logf(ssum);
        okd += A1 * inv;                 // sum_k R dist
        oent += inv * (B1 - ls * S1);    // sum_k sigma R log R,  log R = u - log(sum)
      }
      // ---- sweep 2: R rows (stored) and the column sums of the warp's 32 rows ----
      for (int c = 0; c < NP; c += 16) {
        float v[16];
        umma::tmem_ld16(trow + c, v);
        umma::tmem_ld_wait();
        float rr[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float dist = 2.f * (1.f - v[i]);
          rr[i] = __expf(dist * isig[c + i]) * inv;
        }
        if (live) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            if (c + i < KS) *reinterpret_cast<float4*>(rg + c + i) = make_float4(rr[i], rr[i + 1], rr[i + 2], rr[i + 3]);
        }
        // butterfly: after the five exchanges lane L holds the 32-row sum of column c + bcol(L)
        float w8[8], w4[4], w2[2];
        {
          const bool up = (lane & 16) != 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float send = up ? rr[j] : rr[j + 8];
            const float keep = up ? rr[j + 8] : rr[j];
            w8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool up = (lane & 8) != 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float send = up ? w8[j] : w8[j + 4];
            const float keep = up ? w8[j + 4] : w8[j];
            w4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
        {
          const bool up = (lane & 4) != 0;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float send = up ? w4[j] : w4[j + 2];
            const float keep = up ? w4[j + 2] : w4[j];
            w2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
        }
        float w1;
        {
          const bool up = (lane & 2) != 0;
          const float send = up ? w2[0] : w2[1];
          const float keep = up ? w2[1] : w2[0];
          w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
        if ((lane & 1) == 0 && c + bcol < K) atomicAdd(cs + c + bcol, w1);
      }
      // TMEM accumulator fully read -> the issuer may overwrite it
      umma::fence_before_sync();
      umma::mbar_arrive(t_empty + g);
      // column sums of the tile -> O[level], row sums
      umma::named_sync(bar_id, 128);
      for (int k = et; k < K; k += 128) {
        const float t = cs[k];
        cs[k] = 0.f;
        atomicAdd(a.rs_acc + k, t);
        for (int c = 0; c < a.C; ++c) atomicAdd(a.O_acc + (size_t)a.tuple_levels[tq * a.C + c] * KS + k, t);
      }
      umma::named_sync(bar_id, 128);
    }
    okd = warp_sum(okd);
    oent = warp_sum(oent);
    if (lane == 0) {
      atomicAdd(a.obj_acc + 0, (double)okd);
      atomicAdd(a.obj_acc + 1, (double)oent);
    }
  }
  // ---- teardown ----
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

}  // namespace hb
