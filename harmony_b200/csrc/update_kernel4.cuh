// update_kernel4.cuh — single-pass update_R (harmony.cpp:269-342): every U row is read from HBM exactly ONCE per
// clustering round.
//
// The reference walks the blocks of a random partition sequentially (Gauss-Seidel on the K x B tables O, E):
//     O -= colsum(R_blk);  P = ((2E+1)/(O+E+1))^theta;  R_blk = L1norm(exp(U_blk) * P[level]);  O += colsum(R_blk)
// The removal term of a block, rem_{t,j} = sum of the CURRENT R over the cells of block j of round t, is the sum of
// rows that were last written in round t-1.  All T update orders of a cluster_cpp call are known before the first
// step runs, so the round that produces a row also files it under the block that will remove it: while round t
// updates a cell it adds the new row to add_{t,j} (its own block) and to remT[block of the cell in round t+1].
// Nothing is ever re-read: no look-ahead pass, no second (L2) touch, no saved penalty tables.  Round 0 takes its
// removal sums from k_rem_sums (one gather pass over the R of the assignment step / of the user).
//
// This header holds the algorithm's global protocol (argument block, table derivation, exchange words, the small
// helper kernels); the persistent kernel itself — how a CTA moves and reduces its rows — is update_kernel5.cuh.
//
// Global tables (same accumulator-slot scheme as the first generation): acc slot(s) = [add_{s-1} | rem_s],
//   O_s = (O_{s-1} - rem_{s-1}) + add_{s-1},  E likewise with row sums * Pr_b,
//   P_s = ((2 (E_s - rs_rem_s Pr_b) + 1) / ((O_s - rem_s) + (E_s - rs_rem_s Pr_b) + 1)) ^ theta.
// remT[parity][nb][J][KS] collects the next round's removal sums per (block, tuple); the first CTAs to arrive in a
// new round fold it into the rem halves of that round's slots (one extra grid-wide counter per round).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace hb {

constexpr int U4_MAXNV = 2;  // K <= 128 * U4_MAXNV (wider rows leave no room for the row rings)

struct Upd4Args {
  const float* U;        // [n][KS]
  float* R;              // [n][KS]
  const int* order;      // [T][n]  rows sorted by (block, tuple, block in the next round, cell)
  const int* next_at;    // [T][n]  block in round t+1 of the cell at each position of `order`
  const int4* ranges;    // [T*nb][grid] (lo, hi, tuple, 0): each CTA's slice of a block lies inside ONE tuple
  const int* tuple_levels;  // [J][C]
  const int* lvl_ptr;       // [B + 1]  CSR level -> tuples that contain it (ascending tuple order)
  const int* lvl_tup;       // [J C]
  const int* lvl_first1;    // [1] first level of covariate 1 (= B_vec[0])
  const float* sigma;    // [K]
  const float* theta;    // [B]
  const float* Pr_b;     // [B]
  float* ring;           // [2 parity][2 (O,E)][B][KS]
  float* acc;            // [(S+2)][SL]  slot(s) at (s+1)*SL: [add_O B*KS | add_rs KS | rem_O B*KS | rem_rs KS]
  float* remT;           // [2 parity][nb][J][KS]
  float* remS;           // [nb][J][KS]  sharded cells: sum of the ranks' remT of the round being folded
  float* OEend;          // [T][2][B][KS]  tables at the end of each round (for the objective)
  double* obj;           // [T][2]
  unsigned* bar;         // cntU[s + 1] for s = -1 .. S, then cntF[t] for t = 0 .. T
  int64_t n;
  int K, KS, C, J, B, nb, T;
  int s_begin, s_end;    // steps [s_begin, s_end) of this launch
  int write_from;        // rounds t >= write_from store R
  int has_next_from;     // rounds t < has_next_from file their rows under the next round's blocks
  int sigma_uniform;
  float sigma0;
  int ring_rows;         // update_kernel5.cuh: rows of every warp's private ring
  int dbg_flags;         // timing experiments only (HB_U5_FLAGS): 1 = no remT reductions, 2 = no R stores (results invalid)
  int coop;              // 1: one launch covers many steps (counters + in-kernel fold); 0: single-step launch
  long long* dbg;        // optional [steps][8] globaltimer stamps of CTA dbg_cta (null = off)
  int dbg_cta;
};
// Sharded cells on one node (one process per GPU): the per-step sums cross GPUs through peer memory instead of a
// collective per block step.  Every rank owns an exchange area that all ranks of the node have mapped (CUDA IPC
// over NVLink):  inbox[slot][src][XH]  (XH = B KS + KS values: the add half of an accumulator slot) and the rank's
// remT tables.  The CTA that completes the LOCAL add half of a slot copies it into entry src = rank of every
// rank's inbox; readers add the entries in rank order, so every rank derives bit-identical tables.  The next
// round's removal sums are read from the peers' remT directly when a round is folded.
constexpr int U4_MAXWORLD = 8;
struct Upd4Xch {
  int world = 1, rank = 0;
  unsigned epoch = 0;
  int XH = 0;
  uint2* inbox = nullptr;       // local area: [slot][src][XH] of (value bits, epoch)
  uint2* peer_inbox[U4_MAXWORLD] = {};
  float* peer_remT[U4_MAXWORLD] = {};  // every rank's remT (own entry = local pointer)
};

// Low-latency exchange words: every value travels with the epoch of the running cluster_cpp call in ONE 8-byte
// store (single-copy atomic), so a reader needs neither a flag nor a fence: it polls the word it wants until the
// epoch matches.  One NVLink one-way latency per block step instead of data + system fence + flag.
__device__ __forceinline__ uint2 u4_ld_ll(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void u4_st_ll(uint2* p, float val, unsigned epoch) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(__float_as_uint(val)), "r"(epoch) : "memory");
}
// two consecutive words in one 16-byte store (p 16-byte aligned): each 64-bit element is single-copy atomic and
// carries its own epoch, so the pair needs no atomicity as a whole — half the packets on the fabric
__device__ __forceinline__ void u4_st_ll2(uint2* p, float v0, float v1, unsigned epoch) {
  const unsigned long long w0 = ((unsigned long long)epoch << 32) | __float_as_uint(v0);
  const unsigned long long w1 = ((unsigned long long)epoch << 32) | __float_as_uint(v1);
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
// sum over the ranks' entries of element `off` of the add half of accumulator slot index `sl`, in rank order
// (waits for entries that have not arrived yet)
__device__ __forceinline__ float u4_xsum(const Upd4Xch& x, int sl, int off) {
  const uint2* q = x.inbox + ((size_t)sl * x.world) * x.XH + off;
  float t = 0.f;
  for (int r = 0; r < x.world; ++r) {
    uint2 v = u4_ld_ll(q + (size_t)r * x.XH);
    while (v.y != x.epoch) {
      __nanosleep(20);
      v = u4_ld_ll(q + (size_t)r * x.XH);
    }
    t += __uint_as_float(v.x);
  }
  return t;
}

struct Upd4Launch {
  Upd4Args a;
  Upd4Xch x;
};

__device__ __forceinline__ unsigned u4_ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void u4_red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ((2E+1)/(O+E+1))^theta (harmony_pow, utils.cpp:84-90) as ex2(theta * lg2(x)); x > 0, theta = 0 gives exactly 1.
__device__ __forceinline__ float u4_penalty_pow(float o_eff, float e_eff, float th) {
  const float x = ((2.f * e_eff) + 1.f) / (o_eff + e_eff + 1.f);
  float l, y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(th * l));
  return y;
}

struct U4Tables {
  const float* ringO;   // O_{s-1}
  const float* ringE;
  const float* prev;    // slot(s-1)
  const float* cur;     // slot(s)
  int BK, KS;
};
// O_s, E_s and P_s of element (b, k); every input was completed before the counter this CTA waited for.
// xs >= 0: sharded cells, the add half of slot index xs is the sum of the ranks' inbox entries (xs < 0: local / none)
__device__ __forceinline__ void u4_derive(const U4Tables& tv, const float* Pr_b, const float* theta, int b, int k, float& o,
                                          float& e, float& p, const Upd4Xch* x = nullptr, int xs = -1) {
  const int idx = b * tv.KS + k;
  const float prb = __ldg(Pr_b + b);
  const float* prev_rem_O = tv.prev + tv.BK + tv.KS;
  const float* prev_rem_rs = prev_rem_O + tv.BK;
  const float* cur_add_O = tv.cur;
  const float* cur_add_rs = tv.cur + tv.BK;
  const float* cur_rem_O = tv.cur + tv.BK + tv.KS;
  const float* cur_rem_rs = cur_rem_O + tv.BK;
  float add_o, add_rs;
  if (x) {
    add_o = (xs >= 0) ? u4_xsum(*x, xs, idx) : 0.f;
    add_rs = (xs >= 0) ? u4_xsum(*x, xs, tv.BK + k) : 0.f;
  } else {
    add_o = __ldcg(cur_add_O + idx);
    add_rs = __ldcg(cur_add_rs + k);
  }
  o = (__ldcg(tv.ringO + idx) - __ldcg(prev_rem_O + idx)) + add_o;
  e = (__ldcg(tv.ringE + idx) - __ldcg(prev_rem_rs + k) * prb) + add_rs * prb;
  const float e_eff = e - __ldcg(cur_rem_rs + k) * prb;
  const float o_eff = o - __ldcg(cur_rem_O + idx);
  p = u4_penalty_pow(o_eff, e_eff, __ldg(theta + b));
}

__host__ __device__ inline int upd4_nv(int KS) {
  int nv = 1;
  while (128 * nv < KS) nv <<= 1;
  return nv;
}

// Removal sums of a round from R in memory (round 0 of every cluster_cpp call: the R of the assignment step, or
// the R the user wrote): slot(t0 nb + j).rem += column sums of the rows of block j, per level.
//   grid = (G, nb): CTA (x, j) reduces range x of block j of the plan (the ranges of the update kernel).
__global__ void __launch_bounds__(256) k_rem_sums(const float* __restrict__ R, const int* __restrict__ order,
                                                  const int4* __restrict__ ranges, const int* __restrict__ tuple_levels,
                                                  float* __restrict__ acc, int s0, int G, int K, int KS, int C, int B) {
  extern __shared__ __align__(16) float sm_part[];  // [8][KS]
  const int j = blockIdx.y, x = blockIdx.x;
  const int4 rg = ranges[(size_t)j * G + x];
  const int lo = rg.x, hi = rg.y, q = rg.z;
  if (hi <= lo) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KS4 = KS >> 2;
  const int BK = B * KS, SL = 2 * (BK + KS);
  for (int c4 = lane; c4 < KS4; c4 += 32) {
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = lo + warp; r < hi; r += 8) {
      const float4 v = ld_stream4(reinterpret_cast<const float4*>(R + (size_t)__ldg(order + r) * KS) + c4);
      s4.x += v.x;
      s4.y += v.y;
      s4.z += v.z;
      s4.w += v.w;
    }
    *reinterpret_cast<float4*>(sm_part + (size_t)warp * KS + 4 * c4) = s4;
  }
  __syncthreads();
  float* slot = acc + (size_t)(s0 + j + 1) * SL;
  float* rem_O = slot + BK + KS;
  float* rem_rs = rem_O + BK;
  for (int k = tid; k < K; k += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm_part[(size_t)w * KS + k];
    atomicAdd(rem_rs + k, t);
    for (int c = 0; c < C; ++c) atomicAdd(rem_O + (size_t)tuple_levels[q * C + c] * KS + k, t);
  }
}

// Fold of remT (the next round's removal sums per (block, tuple), filed during round t - 1) into the rem halves of
// round t's accumulator slots, one (block j, cluster k) column per call: rem_O[b][k] = sum over the tuples that
// contain level b, in ascending tuple order (the same order on every rank and in every launch mode); the row sum is
// the sum over the levels of covariate 0 (every tuple has exactly one).  Plain stores: the column has one owner.
// `table`: the [nb][J][KS] sums of round t — this rank's remT parity, or (sharded cells) the rank-ordered sum of all
// ranks' tables that u4_gather_remT left in a.remS.
__device__ __forceinline__ void u4_fold_column(const Upd4Args& a, const float* table, int t, int j, int k) {
  const int KS = a.KS, J = a.J, nb = a.nb, B = a.B;
  const int BK = B * KS, SL = 2 * (BK + KS);
  const size_t joff = (size_t)j * J * KS + k;  // within `table` = the [nb][J][KS] sums of round t
  float* Tz = a.remT + (size_t)((t + 1) & 1) * nb * J * KS + (size_t)j * J * KS + k;
  float* slot = a.acc + (size_t)(t * nb + j + 1) * SL;
  float* rem_O = slot + BK + KS;
  float* rem_rs = rem_O + BK;
  auto val = [&](int q) -> float { return __ldcg(table + joff + (size_t)q * KS); };
  const int B0 = (a.C > 1) ? __ldg(a.lvl_first1) : B;  // levels of covariate 0 come first
  float rs = 0.f;
  for (int b0 = 0; b0 < B; b0 += 4) {  // four levels at a time: their loads are in flight together
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + i;
      if (b < B)
        for (int e = __ldg(a.lvl_ptr + b); e < __ldg(a.lvl_ptr + b + 1); ++e) o[i] += val(__ldg(a.lvl_tup + e));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + i;
      if (b < B) {
        rem_O[(size_t)b * KS + k] = o[i];
        if (b < B0) rs += o[i];
      }
    }
  }
  rem_rs[k] = rs;
  for (int q = 0; q < J; ++q) Tz[(size_t)q * KS] = 0.f;  // the other parity collects round t + 1's sums
}
// Stand-alone fold (per-step launches: sharded cells without the peer exchange, where remT is all-reduced by the
// host in between).  Same arithmetic as the in-kernel fold.
__global__ void k_fold_round(Upd4Args a, int t) {
  const float* table = a.remT + (size_t)(t & 1) * a.nb * a.J * a.KS;
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < a.nb * a.K; item += gridDim.x * blockDim.x)
    u4_fold_column(a, table, t, item / a.K, item % a.K);
}

// After the last executed step S: O = O_S, E = E_S into the handle's tables.
__global__ void k_update_finalize4(Upd4Args a, Upd4Xch x, int S, float* __restrict__ O, float* __restrict__ E) {
  const int KS = a.KS, BK = a.B * KS, SL = 2 * (BK + KS);
  const bool multi = x.world > 1;
  // sharded cells: u4_xsum waits for the other ranks' last add halves (their kernels may still be running)
  const float* ringO = a.ring + (size_t)((S - 1) & 1) * 2 * BK;
  const float* ringE = ringO + BK;
  const float* prev = a.acc + (size_t)(S)*SL;
  const float* cur = a.acc + (size_t)(S + 1) * SL;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < BK; idx += gridDim.x * blockDim.x) {
    const int b = idx / KS, k = idx - b * KS;
    float o = 0.f, e = 0.f;
    if (k < a.K) {
      // same arithmetic as u4_derive() without the removal of step S (which never runs)
      const float prb = a.Pr_b[b];
      const float* prev_rem_O = prev + BK + KS;
      const float* prev_rem_rs = prev_rem_O + BK;
      const float add_o = multi ? (S >= 1 ? u4_xsum(x, S + 1, idx) : 0.f) : cur[idx];
      const float add_rs = multi ? (S >= 1 ? u4_xsum(x, S + 1, BK + k) : 0.f) : cur[BK + k];
      o = (ringO[idx] - prev_rem_O[idx]) + add_o;
      e = (ringE[idx] - prev_rem_rs[k] * prb) + add_rs * prb;
    }
    O[idx] = o;
    E[idx] = e;
  }
}

}  // namespace hb
