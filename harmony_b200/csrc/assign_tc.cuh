// assign_tc.cuh — K1 on the 5th-generation tensor cores: the assignment step of harmony.cpp:141-150 / :220-227
//   dist = 2 (1 - Y^T z),  U = -dist / sigma,  R = softmax_k(U),  O[level] += column sums,  objective sums
// for tiles of 128 cells, with the K x d contraction as tcgen05.mma.kind::tf32 in 3xTF32 error-compensated
// form (a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi; fp32-grade, see scripts/mb/umma_test.cu) and the fp32
// accumulator tile (128 cells x N clusters) in TMEM.
//
// Persistent CTA (one per SM), warp-specialised:
//   warps 0-3  loader   : one cell row per thread: global -> (L2-normalise, write back) -> tf32 hi/lo split
//                         -> canonical no-swizzle K-major operand tiles in shared memory; one elected thread
//                         of warp 0 then issues the 3 x (d/8) MMAs of the tile and commits to mbarriers
//   warps 4-11 epilogue : two warpgroups, one per TMEM accumulator (tiles alternate): one TMEM lane (= cell)
//                         per thread: softmax pieces in registers, U rows stored directly, R rows staged in
//                         shared memory -> column sums per cluster + one 1-D bulk (TMA) store of the R tile
// The two accumulators let the epilogue of tile i overlap the load + MMA of tiles i+1, i+2.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace hb {

constexpr int TC_TM = 128;        // cells per tile (= UMMA M)
constexpr int TC_THREADS = 384;   // 12 warps: 4 loader (warp 0 also issues the MMAs), 2 x 4 epilogue
constexpr int TC_DS4MAX = 16;     // float4 per embedding row held in registers by the loader (d <= 64)

struct AssignTcArgs {
  float* Zc;             // [n][DS]
  const float* Y;        // [K][d]
  const float* sigma;    // [K]
  float* U;              // [n][KS]
  float* R;              // [n][KS]
  const int* tile_cell0;  // TC_TM-cell tiles that do not straddle covariate tuples
  const int* tile_len;
  const int* tile_tuple;
  const int* tuple_levels;  // [J][C]
  float* O_acc;             // [B][KS]
  float* rs_acc;            // [KS]
  double* obj_acc;          // [2]
  int ntiles, d, K, C, DS, KS;
  int KD;   // reduction length padded to a multiple of 8 (tf32 UMMA K)
  int NP;   // clusters padded to a multiple of 16 (UMMA N)
  int normalise;
  long long* dbg;  // optional: [tiles of CTA 0][3 roles][8] globaltimer stamps
};

__host__ __device__ inline size_t assign_tc_smem_bytes(int KD, int NP, int KS) {
  return sizeof(float) * (2 * (size_t)TC_TM * KD + 2 * (size_t)NP * KD + 2 * (size_t)TC_TM * KS + 2 * (size_t)NP) + 256;  // A hi/lo, B hi/lo, 2 R stages
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_assign_tc(AssignTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int KD = a.KD, NP = a.NP, KS = a.KS, K = a.K, d = a.d;
  float* Ahi = reinterpret_cast<float*>(smem_raw);   // [KD/4][128][4]
  float* Alo = Ahi + (size_t)TC_TM * KD;
  float* Bhi = Alo + (size_t)TC_TM * KD;             // [KD/4][NP][4]
  float* Blo = Bhi + (size_t)NP * KD;
  float* Rt0 = Blo + (size_t)NP * KD;                // [2][128][KS] staging of the R tile, one per epilogue group
  float* sig = Rt0 + 2 * (size_t)TC_TM * KS;         // [NP]
  float* isig = sig + NP;                            // [NP] -1/sigma
  uint64_t* bars = reinterpret_cast<uint64_t*>(isig + NP);  // a_full, a_empty, t_full[2], t_empty[2]
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* t_full = bars + 2;
  uint64_t* t_empty = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: centroids -> tf32 hi/lo operand tiles, barriers, TMEM ----
  for (int i = tid; i < NP * KD; i += TC_THREADS) {
    const int n = i / KD, k = i - n * KD;
    const float y = (n < K && k < d) ? a.Y[(size_t)n * d + k] : 0.f;
    float hi, lo;
    umma::split_tf32(y, hi, lo);
    const int off = ((k >> 2) * NP + n) * 4 + (k & 3);
    Bhi[off] = hi;
    Blo[off] = lo;
  }
  for (int k = tid; k < NP; k += TC_THREADS) {
    sig[k] = (k < K) ? a.sigma[k] : 1.f;
    isig[k] = (k < K) ? -1.f / a.sigma[k] : -5.0e29f;  // padding columns: dist = 2 -> u = U_PAD -> exp = 0
  }
  if (tid == 0) {
    umma::mbar_init(a_full, 128);
    umma::mbar_init(a_empty, 1);
    umma::mbar_init(t_full + 0, 1);
    umma::mbar_init(t_full + 1, 1);
    umma::mbar_init(t_empty + 0, 128);
    umma::mbar_init(t_empty + 1, 128);
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 256);  // two 128-column accumulators
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const int my_first = blockIdx.x, stride = gridDim.x;
  auto stamp = [&](int it, int role, int slot) {
    if (a.dbg && blockIdx.x == 0 && (tid & 127) == 0 && it < 64) {
      long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      a.dbg[((size_t)it * 3 + role) * 8 + slot] = tns;
    }
  };

  if (warp < 4) {
    // =============================== loader (+ MMA issue by warp 0) ===============================
    const int r = tid;  // row of the tile
    const int DS4 = a.DS >> 2, KD4 = KD >> 2;
    const uint32_t idesc = umma::make_idesc_tf32(TC_TM, NP, 0, 0);
    const uint32_t lboA = TC_TM * 16, lboB = NP * 16, sbo = 128;
    const uint32_t aH = umma::smem_u32(Ahi), aL = umma::smem_u32(Alo), bH = umma::smem_u32(Bhi), bL = umma::smem_u32(Blo);
    float4 z[TC_DS4MAX];
    auto fetch = [&](int tile) {
      if (tile < a.ntiles) {
        const int cell0 = a.tile_cell0[tile], len = a.tile_len[tile];
        if (r < len) {
          const float4* zp = reinterpret_cast<const float4*>(a.Zc + (size_t)(cell0 + r) * a.DS);
#pragma unroll
          for (int c = 0; c < TC_DS4MAX; ++c)
            if (c < DS4) z[c] = ld_stream4(zp + c);
        }
      }
    };
    fetch(my_first);
    int it = 0;
    for (int tile = my_first; tile < a.ntiles; tile += stride, ++it) {
      const int cell0 = a.tile_cell0[tile], len = a.tile_len[tile];
      stamp(it, 0, 0);
      const bool wb = (r < len) && a.normalise;
      if (wb) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < TC_DS4MAX; ++c)
          if (c < DS4) ss += (z[c].x * z[c].x + z[c].y * z[c].y) + (z[c].z * z[c].z + z[c].w * z[c].w);
        float nrm = sqrtf(ss);
        if (nrm == 0.f) nrm = 1.f;
        const float rn = 1.f / nrm;  // one division per row (the loader is on the critical path)
#pragma unroll
        for (int c = 0; c < TC_DS4MAX; ++c)
          if (c < DS4) {
            z[c].x *= rn;
            z[c].y *= rn;
            z[c].z *= rn;
            z[c].w *= rn;
          }
      }
      // the single A stage is free once the MMAs of the previous tile have completed
      stamp(it, 0, 1);
      if (it > 0) umma::mbar_wait(a_empty, (it - 1) & 1);
      stamp(it, 0, 2);
#pragma unroll
      for (int c = 0; c < TC_DS4MAX + 1; ++c)
        if (c < KD4) {
          float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
          if (r < len && c < DS4 && c < TC_DS4MAX) {
            umma::split_tf32(z[c].x, hi.x, lo.x);
            umma::split_tf32(z[c].y, hi.y, lo.y);
            umma::split_tf32(z[c].z, hi.z, lo.z);
            umma::split_tf32(z[c].w, hi.w, lo.w);
          }
          *reinterpret_cast<float4*>(Ahi + ((size_t)c * TC_TM + r) * 4) = hi;
          *reinterpret_cast<float4*>(Alo + ((size_t)c * TC_TM + r) * 4) = lo;
        }
      umma::fence_proxy_async();
      umma::mbar_arrive(a_full);
      stamp(it, 0, 3);
      if (wb) {  // write the normalised row back (harmony.cpp:220) — off the MMA's critical path
        float4* zw = reinterpret_cast<float4*>(a.Zc + (size_t)(cell0 + r) * a.DS);
#pragma unroll
        for (int c = 0; c < TC_DS4MAX; ++c)
          if (c < DS4) zw[c] = z[c];
      }
      fetch(tile + stride);  // next tile's rows: in flight while this tile is multiplied and reduced
      if (warp == 0) {
        // ---- MMA issue: operands of all four loader warps are in place, the accumulator is drained ----
        const int acc = it & 1;
        umma::mbar_wait(a_full, it & 1);
        if (it >= 2) umma::mbar_wait(t_empty + acc, ((it >> 1) - 1) & 1);
        umma::fence_after_sync();
        stamp(it, 0, 4);
        if (lane == 0) {
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
          umma::mma_commit(a_empty);       // operands consumed -> the loader may refill
          umma::mma_commit(t_full + acc);  // accumulator ready -> epilogue group `acc`
        }
        __syncwarp();
        stamp(it, 0, 5);
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int wg = (warp - 4) >> 2;    // epilogue group = TMEM accumulator it serves
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // row of the tile = TMEM lane
    const int et = tid & 127;          // 0..127 within the group
    const int bar_id = 1 + wg;
    float* Rt = Rt0 + (size_t)wg * TC_TM * KS;
    float okd = 0.f, oent = 0.f;
    int it = wg, use = 0;
    for (int tile = my_first + wg * stride; tile < a.ntiles; tile += 2 * stride, it += 2, ++use) {
      const int cell0 = a.tile_cell0[tile], len = a.tile_len[tile], tq = a.tile_tuple[tile];
      stamp(it, 1 + wg, 0);
      umma::mbar_wait(t_full + wg, use & 1);
      umma::fence_after_sync();
      stamp(it, 1 + wg, 1);
      // the staging buffer: the bulk store of this group's previous tile must have finished reading it
      if (et == 0) umma::bulk_wait_read();
      umma::named_sync(bar_id, 128);
      stamp(it, 1 + wg, 2);
      const uint32_t trow = tmem + wg * 128 + ((uint32_t)(q * 32) << 16);
      // pass 1: dist -> u (stored straight to global), e = exp(u) (staged), row sums
      float ssum = 0.f, A1 = 0.f, B1 = 0.f, S1 = 0.f;  // sum e, sum e*dist, sum sigma*e*u, sum sigma*e
      float* rr = Rt + (size_t)r * KS;
      float* ug = a.U + (size_t)(cell0 + r) * KS;
      const bool live = r < len;
      for (int c = 0; c < NP; c += 16) {
        float v[16];
        umma::tmem_ld16(trow + c, v);
        umma::tmem_ld_wait();
        float uu[16], ee[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float dist = 2.f * (1.f - v[i]);
          uu[i] = dist * isig[c + i];  // -dist / sigma (U_PAD in the padding columns)
          ee[i] = __expf(uu[i]);
          ssum += ee[i];
          A1 = fmaf(ee[i], dist, A1);
          const float se = sig[c + i] * ee[i];
          B1 = fmaf(se, uu[i], B1);
          S1 += se;
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          if (c + i < KS) {
            *reinterpret_cast<float4*>(rr + c + i) = make_float4(ee[i], ee[i + 1], ee[i + 2], ee[i + 3]);
            if (live) *reinterpret_cast<float4*>(ug + c + i) = make_float4(uu[i], uu[i + 1], uu[i + 2], uu[i + 3]);
          }
        }
      }
      // TMEM accumulator fully read -> the issuer may overwrite it
      umma::fence_before_sync();
      umma::mbar_arrive(t_empty + wg);
      stamp(it, 1 + wg, 3);
      // R.each_row() /= sum(R, 0) (no zero guard in the reference); rows beyond the tile are zeroed
      const float inv = live ? 1.f / ssum : 0.f;
      if (live) {
        const float ls = __logf(ssum);
        okd += A1 * inv;                 // sum_k R dist
        oent += inv * (B1 - ls * S1);    // sum_k sigma R log R,  log R = u - log(sum)
      }
      {
        const int KS4 = KS >> 2;
        for (int c4 = 0; c4 < KS4; ++c4) {
          float4 e4 = *reinterpret_cast<float4*>(rr + c4 * 4);
          e4.x *= inv;
          e4.y *= inv;
          e4.z *= inv;
          e4.w *= inv;
          *reinterpret_cast<float4*>(rr + c4 * 4) = e4;
        }
      }
      umma::fence_proxy_async();   // the staged tile is read by the bulk-copy engine
      umma::named_sync(bar_id, 128);
      stamp(it, 1 + wg, 4);
      if (et == 0) {               // R rows of a tile are contiguous in global memory
        umma::bulk_store(a.R + (size_t)cell0 * KS, Rt, (uint32_t)len * KS * 4);
        umma::bulk_commit();
      }
      // column sums of the R tile -> O[level], row sums
      for (int k = et; k < K; k += 128) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
        int rr2 = 0;
        for (; rr2 + 3 < len; rr2 += 4) {
          t0 += Rt[(size_t)(rr2 + 0) * KS + k];
          t1 += Rt[(size_t)(rr2 + 1) * KS + k];
          t2 += Rt[(size_t)(rr2 + 2) * KS + k];
          t3 += Rt[(size_t)(rr2 + 3) * KS + k];
        }
        for (; rr2 < len; ++rr2) t0 += Rt[(size_t)rr2 * KS + k];
        const float t = (t0 + t1) + (t2 + t3);
        atomicAdd(a.rs_acc + k, t);
        for (int c = 0; c < a.C; ++c) atomicAdd(a.O_acc + (size_t)a.tuple_levels[tq * a.C + c] * KS + k, t);
      }
      stamp(it, 1 + wg, 5);
    }
    if (et == 0) umma::bulk_wait_all();
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
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace hb
