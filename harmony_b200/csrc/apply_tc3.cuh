// apply_tc3.cuh — K5 on the tensor cores: the correction of harmony.cpp:347 + :615 for all clusters at once,
//   Zc_i = Zo_i - sum_k R_ik V_q[k][:]        (q = covariate tuple of cell i, V_q from k_ridge_solve)
// computed transposed so that the per-tile operand is the small one:
//   D'[c][cell] = sum_k A[c][k] * B[cell][k],   A = V_q^T (static while the tuple does not change, M = 128 rows of
//   which d are used),  B = a 64-cell tile of R (N = 64),  3xTF32, fp32 accumulators in TMEM.
// Second generation.  The first one (bulk-copied raw tiles -> converter warps -> ONE converted operand stage)
// spent 5.4 us per tile on a chain load -> convert -> MMA -> convert ... with a single tile in flight.  Here, as in
// assign_tc3.cuh:
//   * a loader thread owns half a row of the tile: it copies its 16-byte pieces with cp.async straight into the
//     canonical K-major operand layout (the raw fp32 tile IS the `hi` operand: the tensor core reads the upper 19
//     bits), waits for its OWN copies (cp.async.wait_group, no mbarrier between loading and converting), and writes
//     the `lo` pieces; two operand stages, so the rows of tile t + 1 land while tile t is multiplied and stored;
//   * A = V_q^T is rebuilt by the same threads when the tuple changes (after the previous tile's MMAs);
//   * epilogue as before: TMEM lane = embedding column c, Zc[cell][c] = Zo[cell][c] - D'[c][cell], coalesced
//     across the warp; one warp pair per accumulator, tile records fetched one tile ahead.
// Warp roles (448 threads): 0 MMA issuer, 2-5 loaders / converters, 8-9 and 12-13 epilogue (TMEM lane quarters
// 0 and 1 of accumulator 0 / 1), the others idle.  Limits: d <= 64, K <= 256 (shared memory).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace hb {

constexpr int AP_TN = 64;         // cells per tile (= UMMA N), equals the static tile size TM
constexpr int AP_THREADS = 448;   // 14 warps
constexpr int AP_LOAD = 128;      // loader / converter threads (warps 2-5)

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
  // A hi/lo: 2 x 128 x KD; B hi/lo, two stages: 2 x 2 x 64 x KD
  (void)KS;
  return sizeof(float) * (2 * (size_t)128 * KD + 4 * (size_t)AP_TN * KD) + 1024;
}

__global__ void __launch_bounds__(AP_THREADS, 1) k_apply_tc(ApplyTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int KD = a.KD, K = a.K, d = a.d, KS = a.KS, DS = a.DS;
  float* Ahi = reinterpret_cast<float*>(smem_raw);  // [KD/4][128][4]   V_q^T
  float* Alo = Ahi + (size_t)128 * KD;
  float* Bhi = Alo + (size_t)128 * KD;               // [2][KD/4][64][4]  raw R rows = hi operand
  float* Blo = Bhi + 2 * (size_t)AP_TN * KD;         // [2][KD/4][64][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Blo + 2 * (size_t)AP_TN * KD);
  uint64_t* lo_full = bars + 0;    // [2]  loaders (128): operands of the stage complete
  uint64_t* st_empty = bars + 2;   // [2]  tcgen05.commit: operands consumed
  uint64_t* t_full = bars + 4;     // [2]  tcgen05.commit: accumulator ready
  uint64_t* t_empty = bars + 6;    // [2]  epilogue pair (64): accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nch = KD >> 2;   // 16-byte chunks along K in the operand tiles
  const int KS4 = KS >> 2;   // ... in a row of R

  // the padding chunks of B (k >= KS) are never written by the loaders and must be zero (0 x NaN would poison D')
  for (int i = tid; i < 4 * AP_TN * KD; i += AP_THREADS) Bhi[i] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(lo_full + i, AP_LOAD);
      umma::mbar_init(st_empty + i, 1);
      umma::mbar_init(t_full + i, 1);
      umma::mbar_init(t_empty + i, 64);
    }
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 128);  // two 64-column accumulators
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int t_end = (t_begin + a.tiles_per_cta < a.ntiles) ? t_begin + a.tiles_per_cta : a.ntiles;
  auto stamp = [&](int it, int slot) {
    if (a.dbg && blockIdx.x == 0 && it < 32) {
      long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      a.dbg[it * 16 + slot] = tns;
    }
  };

  if (warp == 0) {
    // =============================== MMA issuer (one thread) ===============================
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_tf32(128, AP_TN, 0, 0);
      const uint32_t lboA = 128 * 16, lboB = AP_TN * 16, sbo = 128;
      const uint32_t aH = umma::smem_u32(Ahi), aL = umma::smem_u32(Alo);
      int it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const int s = it & 1, use = it >> 1;
        umma::mbar_wait(lo_full + s, use & 1);
        stamp(it, 5);
        if (use >= 1) umma::mbar_wait(t_empty + s, (use - 1) & 1);
        stamp(it, 6);
        umma::fence_after_sync();
        const uint32_t bH = umma::smem_u32(Bhi + (size_t)s * AP_TN * KD), bL = umma::smem_u32(Blo + (size_t)s * AP_TN * KD);
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
        umma::mma_commit(st_empty + s);
        umma::mma_commit(t_full + s);
        stamp(it, 7);
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // =============================== loaders / converters ===============================
    const int lt = tid - 64;             // 0..127
    const int cell = lt & (AP_TN - 1);   // row of the tile
    const int half = lt >> 6;            // which half of the row's 16-byte pieces
    const int hsplit = (KS4 + 1) >> 1;
    const int c_lo = half ? hsplit : 0, c_hi = half ? KS4 : hsplit;
    auto meta = [&](int it, int& cell0, int& len, int& q) {
      const int tile = t_begin + it;
      cell0 = 0;
      len = 0;
      q = -1;
      if (tile < t_end) {
        cell0 = __ldg(a.tile_cell0 + tile);
        len = __ldg(a.tile_len + tile);
        q = __ldg(a.tile_tuple + tile);
      }
    };
    auto issue_row = [&](int it, int cell0, int len) {  // tile `it` -> stage it & 1; always commits
      if (cell < len) {
        const float* src = a.R + (size_t)(cell0 + cell) * KS;
        const unsigned dst = umma::smem_u32(Bhi + (size_t)(it & 1) * AP_TN * KD + (size_t)cell * 4);
        for (int c = c_lo; c < c_hi; ++c)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)c * (AP_TN * 16u)), "l"(src + 4 * c) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int c0_0, len0, q0, c0_1, len1, q1, c0_2, len2, q2;
    meta(0, c0_0, len0, q0);
    meta(1, c0_1, len1, q1);
    meta(2, c0_2, len2, q2);
    issue_row(0, c0_0, len0);
    int cur_q = -1;
    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int s = it & 1;
      if (lt == 0) stamp(it, 0);
      bool prev_done = (it == 0);  // the MMAs of tile it - 1 are known to be complete
      if (tile + 1 < t_end) {
        if (it >= 1) {
          umma::mbar_wait(st_empty + ((it + 1) & 1), ((it - 1) >> 1) & 1);  // MMAs of tile it - 1: its stage is free
          prev_done = true;
        }
        if (lt == 0) stamp(it, 1);
        issue_row(it + 1, c0_1, len1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      if (lt == 0) stamp(it, 2);
      if (q0 != cur_q) {
        // new tuple: A = V_q^T (k contiguous per embedding column c), tf32 hi/lo — after the previous tile's MMAs
        if (!prev_done) umma::mbar_wait(st_empty + ((it + 1) & 1), ((it - 1) >> 1) & 1);
        const float* Vq = a.V + (size_t)q0 * K * d;
        for (int idx = lt; idx < 128 * KD; idx += AP_LOAD) {
          const int k = idx / 128, c = idx - k * 128;
          const float v = (k < K && c < d) ? Vq[(size_t)k * d + c] : 0.f;
          float hi, lo;
          umma::split_tf32(v, hi, lo);
          const int off = ((k >> 2) * 128 + c) * 4 + (k & 3);
          Ahi[off] = hi;
          Alo[off] = lo;
        }
        cur_q = q0;
      }
      // `lo` pieces of this thread's half row (rows beyond the tile keep whatever the stage holds: their columns of D'
      // are not read)
      {
        const float* hi = Bhi + (size_t)s * AP_TN * KD + (size_t)cell * 4;
        float* lo = Blo + (size_t)s * AP_TN * KD + (size_t)cell * 4;
        if (cell < len0) {
          for (int c = c_lo; c < c_hi; ++c) {
            const float4 z = *reinterpret_cast<const float4*>(hi + (size_t)c * AP_TN * 4);
            float4 l4;
            // the tensor core reads trunc_tf32(z); the remainder is exact in fp32 and is rounded to tf32 here
            l4.x = umma::round_tf32(z.x - __uint_as_float(__float_as_uint(z.x) & 0xffffe000u));
            l4.y = umma::round_tf32(z.y - __uint_as_float(__float_as_uint(z.y) & 0xffffe000u));
            l4.z = umma::round_tf32(z.z - __uint_as_float(__float_as_uint(z.z) & 0xffffe000u));
            l4.w = umma::round_tf32(z.w - __uint_as_float(__float_as_uint(z.w) & 0xffffe000u));
            *reinterpret_cast<float4*>(lo + (size_t)c * AP_TN * 4) = l4;
          }
        }
      }
      umma::fence_proxy_async();  // this thread's cp.async pieces (observed above), `lo` and A writes -> tensor core
      umma::mbar_arrive(lo_full + s);
      if (lt == 0) stamp(it, 3);
      c0_0 = c0_1;
      len0 = len1;
      q0 = q1;
      c0_1 = c0_2;
      len1 = len2;
      q1 = q2;
      meta(it + 3, c0_2, len2, q2);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == 8 || warp == 9 || warp == 12 || warp == 13) {
    // =============================== epilogue (one warp pair per accumulator) ===============================
    const int es = (warp >= 12) ? 1 : 0;     // accumulator served by this pair
    const int wq = warp & 3;                 // TMEM lane quarter: 0 or 1
    const int c = wq * 32 + lane;            // embedding column = TMEM lane
    int use = 0;
    int cell0_n = 0, len_n = 0;
    if (t_begin + es < t_end) {
      cell0_n = __ldg(a.tile_cell0 + t_begin + es);
      len_n = __ldg(a.tile_len + t_begin + es);
    }
    for (int tile = t_begin + es; tile < t_end; tile += 2, ++use) {
      const int cell0 = cell0_n, len = len_n;
      if (tile + 2 < t_end) {
        cell0_n = __ldg(a.tile_cell0 + tile + 2);
        len_n = __ldg(a.tile_len + tile + 2);
      }
      // the tile's Zo values do not depend on the MMA: fetch them while it runs
      float zo[AP_TN];
      if (c < d) {
#pragma unroll
        for (int j = 0; j < AP_TN; ++j) zo[j] = (j < len) ? ld_stream(a.Zo + (size_t)(cell0 + j) * DS + c) : 0.f;
      }
      umma::mbar_wait(t_full + es, use & 1);
      umma::fence_after_sync();
      if (lane == 0 && wq == 0) stamp(2 * use + es, 8);
      const uint32_t trow = tmem + es * AP_TN + ((uint32_t)(wq * 32) << 16);
      // 16 cells at a time (TMEM -> registers -> store): the whole accumulator would not fit beside zo[]
#pragma unroll
      for (int j = 0; j < AP_TN; j += 16) {
        float t16[16];
        umma::tmem_ld16(trow + j, t16);
        umma::tmem_ld_wait();
        if (j + 16 == AP_TN) {  // accumulator fully read -> the issuer may overwrite it
          umma::fence_before_sync();
          umma::mbar_arrive(t_empty + es);
        }
        if (c < d) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (j + i < len) a.Zc[(size_t)(cell0 + j + i) * DS + c] = zo[j + i] - t16[i];
        }
      }
      if (lane == 0 && wq == 0) stamp(2 * use + es, 9);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

}  // namespace hb
