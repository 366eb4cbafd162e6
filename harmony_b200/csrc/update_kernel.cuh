// update_kernel.cuh — persistent cooperative kernel for harmony::update_R (harmony.cpp:269-342).
//
// One launch covers a range of global block steps s = round * nb + block of one cluster_cpp call.  Steps
// are separated by a grid-wide barrier (the Gauss-Seidel dependency on O, E).  Within step s every CTA
//   A. (owners) materialises the K x B tables O_s, E_s and saves the penalty table P_s,
//   B. updates its share of the rows of block s:    R = L1norm(exp(U) * sum_c P_s[level_c])   (:318-323)
//      accumulating the new column sums per level (add_s, :329-330) and the objective partials,
//   C. looks ahead: the column sums of the *current* R of its share of block s+1 (rem_{s+1}, :312-313),
//      recomputed from U and the penalty tables of the previous round (R itself is only stored in the
//      last round; in the first round of a call it is read from memory).
// HBM traffic per cell and round: one read of the U row (the second touch, one step later, hits L2).
//
// Per-step accumulators live in `acc`: slot(s) = [add_{s-1} | rem_s], s = -1 .. S, zeroed by the host
// before the first launch of a call.  Derivation used by everybody (reference order of operations):
//   O_s = (O_{s-1} - rem_{s-1}) + add_{s-1}        E likewise with rowsums * Pr_b
//   P_s = ((2 (E_s - rs_rem_s Pr_b) + 1) / ((O_s - rem_s) + (E_s - rs_rem_s Pr_b) + 1)) ^ theta
#pragma once
#include "common.cuh"

namespace hb {

constexpr int UPD_THREADS = 512;
constexpr int UPD_WARPS = UPD_THREADS / 32;
constexpr float U_PAD = -1.0e30f;  // logit of the padding columns (exp -> 0)

struct UpdArgs {
  const float* U;   // [n][KS]
  float* R;         // [n][KS]
  const int* order;      // [T][n]   rows sorted by (block, tuple, cell) per round
  const int* seg_start;  // [T][nb*J + 1]
  const int* blk_of;     // [T][n]   block of each cell per round
  const int* tuple_levels;  // [J][C]
  const float* sigma;    // [K]
  const float* theta;    // [B]
  const float* Pr_b;     // [B]
  float* ring;           // [2 parity][2 (O,E)][B][KS]
  float* acc;            // [(S+2)][SL]  slot(s) at (s+1)*SL: [addprev_O B*KS | addprev_rs KS | rem_O B*KS | rem_rs KS]
  float* Psave;          // [2 round parity][nb][B][KS]
  float* OEend;          // [T][2][B][KS]  tables at the end of each round (for the objective)
  double* obj;           // [T][2]
  unsigned* bar;         // [2] grid barrier state
  int64_t n;
  int K, KS, C, J, B, nb, T;
  int s_begin, s_end;    // steps [s_begin, s_end)
  int prologue;          // run the look-ahead for s_begin before the first step (then barrier)
  int first_round_from_R;  // look-ahead of round 0 reads R from memory
  unsigned write_R_mask;   // bit t: store R in round t
  int use_barrier;         // 0 when the launch covers a single step and no prologue (no co-residency needed)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned gen = ld_acquire_u32(bar + 1);
    unsigned arrived;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(arrived) : "l"(bar) : "memory");
    if (arrived == nblocks - 1) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + 1) : "memory");
    } else {
      while (ld_acquire_u32(bar + 1) == gen) __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ float penalty_pow(float o_eff, float e_eff, float th) {
  return powf(((2.f * e_eff) + 1.f) / (o_eff + e_eff + 1.f), th);
}

struct TableView {
  const float* ringO;   // O_{s-1}
  const float* ringE;
  const float* prev;    // slot(s-1)
  const float* cur;     // slot(s)
  int BK, KS;
};
// O_s, E_s and P_s of element (b, k); all inputs were completed before the last barrier -> L2 loads (.cg)
__device__ __forceinline__ void derive(const TableView& tv, const float* Pr_b, const float* theta, int b, int k,
                                       float& o, float& e, float& p) {
  const int idx = b * tv.KS + k;
  const float prb = Pr_b[b];
  const float* prev_rem_O = tv.prev + tv.BK + tv.KS;
  const float* prev_rem_rs = prev_rem_O + tv.BK;
  const float* cur_add_O = tv.cur;
  const float* cur_add_rs = tv.cur + tv.BK;
  const float* cur_rem_O = tv.cur + tv.BK + tv.KS;
  const float* cur_rem_rs = cur_rem_O + tv.BK;
  o = (__ldcg(tv.ringO + idx) - __ldcg(prev_rem_O + idx)) + __ldcg(cur_add_O + idx);
  e = (__ldcg(tv.ringE + idx) - __ldcg(prev_rem_rs + k) * prb) + __ldcg(cur_add_rs + k) * prb;
  const float e_eff = e - __ldcg(cur_rem_rs + k) * prb;
  const float o_eff = o - __ldcg(cur_rem_O + idx);
  p = penalty_pow(o_eff, e_eff, theta[b]);
}

// NV = float4 per lane (8 lanes per row): K <= 32*NV
template <int NV>
__global__ void __launch_bounds__(UPD_THREADS, 1) k_update_steps(UpdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int K = a.K, KS = a.KS, C = a.C, J = a.J, B = a.B, nb = a.nb;
  const int BK = B * KS;
  const int SL = 2 * (BK + KS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 3, gl = lane & 7;  // 4 row groups of 8 lanes
  const int KS4 = KS >> 2;
  // shared memory carve-up
  float* Psum = smem;                       // [KS]   sum_c P_s[level_c] of the current tuple run
  float* lP = Psum + KS;                    // [KS]   log of it
  float* sig = lP + KS;                     // [KS]
  float* part = sig + KS;                   // [UPD_WARPS][KS] column-sum partials
  float* Pprev = part + UPD_WARPS * KS;     // [nb][KS] previous-round penalty sums of the current tuple run
  int* segs = reinterpret_cast<int*>(Pprev + (size_t)nb * KS);  // [J + 1] segment starts of the current block
  __shared__ double sh_obj[2];

  for (int k = tid; k < KS; k += UPD_THREADS) sig[k] = (k < K) ? a.sigma[k] : 0.f;
  __syncthreads();

  float okd = 0.f, oent = 0.f;  // objective partials of the round being processed (flushed at round ends)

  // ---- helpers -------------------------------------------------------------------------------
  auto tables_for = [&](int s) {
    TableView tv;
    const int par = (s - 1) & 1;
    tv.ringO = a.ring + (size_t)par * 2 * BK;
    tv.ringE = tv.ringO + BK;
    tv.prev = a.acc + (size_t)(s) * SL;      // slot(s-1) at index s
    tv.cur = a.acc + (size_t)(s + 1) * SL;   // slot(s)
    tv.BK = BK;
    tv.KS = KS;
    return tv;
  };
  // CTA-wide flush of per-lane column sums cs[NV][4] into dst_O[level][k] (+ dst_rs[k]) for tuple q
  auto flush = [&](float (&cs)[NV][4], int q, float* dst_O, float* dst_rs) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float x = cs[v][c];
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        cs[v][c] = x;
      }
    if (grp == 0) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        int q4 = gl + 8 * v;
        if (q4 < KS4) *reinterpret_cast<float4*>(part + warp * KS + q4 * 4) = make_float4(cs[v][0], cs[v][1], cs[v][2], cs[v][3]);
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) cs[v][c] = 0.f;
    __syncthreads();
    for (int k = tid; k < K; k += UPD_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < UPD_WARPS; ++w) t += part[w * KS + k];
      atomicAdd(dst_rs + k, t);
      for (int c = 0; c < C; ++c) atomicAdd(dst_O + (size_t)a.tuple_levels[q * C + c] * KS + k, t);
    }
    __syncthreads();
  };
  // rows [lo, hi) of block (t, j) assigned to this CTA
  auto my_range = [&](int t, int j, int& lo, int& hi) {
    const int* ss = a.seg_start + (size_t)t * (nb * J + 1) + (size_t)j * J;
    for (int i = tid; i <= J; i += UPD_THREADS) segs[i] = ss[i];
    __syncthreads();
    const int b0 = segs[0], b1 = segs[J];
    const int64_t len = b1 - b0;
    lo = b0 + (int)((len * blockIdx.x) / gridDim.x);
    hi = b0 + (int)((len * (blockIdx.x + 1)) / gridDim.x);
  };
  auto first_seg = [&](int p) {  // last q with segs[q] <= p  (p < segs[J])
    int l = 0, h = J;
    while (h - l > 1) {
      int m = (l + h) >> 1;
      if (segs[m] <= p) l = m; else h = m;
    }
    return l;
  };

  // ---- phase C: look-ahead column sums of block (t, j) = step s, into slot(s).rem -------------------
  auto lookahead = [&](int s) {
    const int t = s / nb, j = s - t * nb;
    int lo, hi;
    my_range(t, j, lo, hi);
    if (lo >= hi) return;
    const bool fromR = (t == 0) && a.first_round_from_R;
    float* slot = a.acc + (size_t)(s + 1) * SL;
    float* rem_O = slot + BK + KS;
    float* rem_rs = rem_O + BK;
    const int* order = a.order + (size_t)t * a.n;
    const int* blkprev = a.blk_of + (size_t)(t > 0 ? t - 1 : 0) * a.n;
    float cs[NV][4];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) cs[v][c] = 0.f;
    int p = lo;
    int q = first_seg(p);
    while (p < hi) {
      while (segs[q + 1] <= p) ++q;  // skip empty segments
      const int run_end = min(hi, segs[q + 1]);
      if (!fromR) {
        // previous-round penalty sums of tuple q for every block jp: Pprev[jp][k]
        const int sp_cur = s - 1;                 // the step running right now (its P is not saved yet)
        const int tp = t - 1;
        for (int idx = tid; idx < nb * KS; idx += UPD_THREADS) {
          const int jp = idx / KS, k = idx - jp * KS;
          float v = 0.f;
          if (k < K) {
            const int sp = tp * nb + jp;
            if (sp == sp_cur) {
              TableView tv = tables_for(sp);
              for (int c = 0; c < C; ++c) {
                float o, e, pp;
                derive(tv, a.Pr_b, a.theta, a.tuple_levels[q * C + c], k, o, e, pp);
                v += pp;
              }
            } else {
              const float* Ps = a.Psave + ((size_t)(tp & 1) * nb + jp) * BK;
              for (int c = 0; c < C; ++c) v += __ldcg(Ps + (size_t)a.tuple_levels[q * C + c] * KS + k);
            }
          }
          Pprev[idx] = v;
        }
        __syncthreads();
      }
      for (int r0 = p + warp * 4; r0 < run_end; r0 += 4 * UPD_WARPS) {
        const int row = r0 + grp;
        const bool valid = row < run_end;
        const int cell = valid ? order[row] : 0;
        if (fromR) {
          const float4* rp = reinterpret_cast<const float4*>(a.R + (size_t)cell * KS);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            int q4 = gl + 8 * v;
            if (valid && q4 < KS4) {
              float4 x = __ldcg(rp + q4);
              cs[v][0] += x.x;
              cs[v][1] += x.y;
              cs[v][2] += x.z;
              cs[v][3] += x.w;
            }
          }
        } else {
          const int jp = valid ? blkprev[cell] : 0;
          const float4* up = reinterpret_cast<const float4*>(a.U + (size_t)cell * KS);
          const float4* pp = reinterpret_cast<const float4*>(Pprev + (size_t)jp * KS);
          float e[NV][4];
          float ssum = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            int q4 = gl + 8 * v;
            if (valid && q4 < KS4) {
              float4 u = ld_stream4(up + q4);
              float4 pw = pp[q4];
              e[v][0] = expf(u.x) * pw.x;
              e[v][1] = expf(u.y) * pw.y;
              e[v][2] = expf(u.z) * pw.z;
              e[v][3] = expf(u.w) * pw.w;
            } else {
              e[v][0] = e[v][1] = e[v][2] = e[v][3] = 0.f;
            }
            ssum += (fabsf(e[v][0]) + fabsf(e[v][1])) + (fabsf(e[v][2]) + fabsf(e[v][3]));
          }
          ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
          ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
          ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);
          const float sdiv = (ssum == 0.f) ? 1.f : ssum;
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) cs[v][c] += e[v][c] / sdiv;
        }
      }
      flush(cs, q, rem_O, rem_rs);
      p = run_end;
    }
  };

  // ---- phase B: update the rows of block (t, j) = step s ---------------------------------------------
  auto update = [&](int s) {
    const int t = s / nb, j = s - t * nb;
    int lo, hi;
    my_range(t, j, lo, hi);
    if (lo >= hi) return;
    const bool writeR = (a.write_R_mask >> t) & 1u;
    TableView tv = tables_for(s);
    float* nslot = a.acc + (size_t)(s + 2) * SL;  // slot(s+1): add_s goes to its addprev part
    float* add_O = nslot;
    float* add_rs = nslot + BK;
    const int* order = a.order + (size_t)t * a.n;
    float cs[NV][4];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) cs[v][c] = 0.f;
    int p = lo;
    int q = first_seg(p);
    while (p < hi) {
      while (segs[q + 1] <= p) ++q;
      const int run_end = min(hi, segs[q + 1]);
      for (int k = tid; k < KS; k += UPD_THREADS) {
        float v = 0.f;
        if (k < K)
          for (int c = 0; c < C; ++c) {
            float o, e, pp;
            derive(tv, a.Pr_b, a.theta, a.tuple_levels[q * C + c], k, o, e, pp);
            v += pp;
          }
        Psum[k] = v;
        lP[k] = (k < K) ? logf(v) : 0.f;
      }
      __syncthreads();
      for (int r0 = p + warp * 4; r0 < run_end; r0 += 4 * UPD_WARPS) {
        const int row = r0 + grp;
        const bool valid = row < run_end;
        const int cell = valid ? order[row] : 0;
        const float4* up = reinterpret_cast<const float4*>(a.U + (size_t)cell * KS);
        float e[NV][4];
        float ssum = 0.f, Aacc = 0.f, Bacc = 0.f, Sacc = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          int q4 = gl + 8 * v;
          if (valid && q4 < KS4) {
            float4 u = ld_stream4(up + q4);
            float4 pw = *reinterpret_cast<const float4*>(Psum + q4 * 4);
            float4 lp = *reinterpret_cast<const float4*>(lP + q4 * 4);
            float4 sg = *reinterpret_cast<const float4*>(sig + q4 * 4);
            float uu[4] = {u.x, u.y, u.z, u.w}, pv[4] = {pw.x, pw.y, pw.z, pw.w};
            float lv[4] = {lp.x, lp.y, lp.z, lp.w}, sv[4] = {sg.x, sg.y, sg.z, sg.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float ee = expf(uu[c]) * pv[c];
              e[v][c] = ee;
              float tt = sv[c] * ee;
              Aacc = fmaf(tt, uu[c], Aacc);
              Bacc = fmaf(tt, lv[c], Bacc);
              Sacc += tt;
            }
          } else {
            e[v][0] = e[v][1] = e[v][2] = e[v][3] = 0.f;
          }
          ssum += (fabsf(e[v][0]) + fabsf(e[v][1])) + (fabsf(e[v][2]) + fabsf(e[v][3]));
        }
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);
        const float sdiv = (ssum == 0.f) ? 1.f : ssum;  // arma::normalise(.., 1, 0): zero norm divides by 1
        const float inv = 1.f / sdiv;
        const float ls = logf(sdiv);
        float4* rp = reinterpret_cast<float4*>(a.R + (size_t)cell * KS);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          int q4 = gl + 8 * v;
          float4 r;
          r.x = e[v][0] / sdiv;
          r.y = e[v][1] / sdiv;
          r.z = e[v][2] / sdiv;
          r.w = e[v][3] / sdiv;
          cs[v][0] += r.x;
          cs[v][1] += r.y;
          cs[v][2] += r.z;
          cs[v][3] += r.w;
          if (writeR && valid && q4 < KS4) rp[q4] = r;
        }
        // sum_k R dist = -sum sigma R U ;  sum_k sigma R log R = sum sigma R (U + log Psum - log s)
        okd -= inv * Aacc;
        oent += inv * (Aacc + Bacc - ls * Sacc);
      }
      flush(cs, q, add_O, add_rs);
      p = run_end;
    }
  };

  // ---- phase A: owners materialise O_s, E_s (ring), the end-of-round tables and save P_s -------------
  auto owners = [&](int s) {
    const int t = s / nb, j = s - t * nb;
    TableView tv = tables_for(s);
    float* outO = a.ring + (size_t)(s & 1) * 2 * BK;
    float* outE = outO + BK;
    float* Ps = a.Psave + ((size_t)(t & 1) * nb + j) * BK;
    for (int idx = blockIdx.x * UPD_THREADS + tid; idx < BK; idx += gridDim.x * UPD_THREADS) {
      const int b = idx / KS, k = idx - b * KS;
      float o = 0.f, e = 0.f, pp = 0.f;
      if (k < K) derive(tv, a.Pr_b, a.theta, b, k, o, e, pp);
      outO[idx] = o;
      outE[idx] = e;
      Ps[idx] = pp;
      if (j == 0 && t > 0) {
        float* oe = a.OEend + (size_t)(t - 1) * 2 * BK;
        oe[idx] = o;
        oe[BK + idx] = e;
      }
    }
  };
  auto flush_objective = [&](int t) {
    okd = warp_sum(okd);
    oent = warp_sum(oent);
    if (tid == 0) {
      sh_obj[0] = 0.0;
      sh_obj[1] = 0.0;
    }
    __syncthreads();
    if (lane == 0) {
      atomicAdd(&sh_obj[0], (double)okd);
      atomicAdd(&sh_obj[1], (double)oent);
    }
    __syncthreads();
    if (tid == 0) {
      atomicAdd(a.obj + 2 * t + 0, sh_obj[0]);
      atomicAdd(a.obj + 2 * t + 1, sh_obj[1]);
    }
    okd = 0.f;
    oent = 0.f;
  };

  const int S_total = a.T * nb;
  if (a.prologue) {
    lookahead(a.s_begin);
    grid_barrier(a.bar, gridDim.x);
  }
  for (int s = a.s_begin; s < a.s_end; ++s) {
    owners(s);
    update(s);
    if ((s + 1) % nb == 0) flush_objective(s / nb);
    if (s + 1 < S_total) lookahead(s + 1);
    if (a.use_barrier && s + 1 < a.s_end) grid_barrier(a.bar, gridDim.x);
  }
  if (a.s_end % nb != 0 && a.s_end > a.s_begin) flush_objective((a.s_end - 1) / nb);  // partial round (per-step launches)
}

// After the last executed step S: O = O_S, E = E_S into the handle's tables (and the objective tables).
__global__ void k_update_finalize(UpdArgs a, int S, float* __restrict__ O, float* __restrict__ E) {
  const int KS = a.KS, BK = a.B * KS, SL = 2 * (BK + KS);
  TableView tv;
  const int par = (S - 1) & 1;
  tv.ringO = a.ring + (size_t)par * 2 * BK;
  tv.ringE = tv.ringO + BK;
  tv.prev = a.acc + (size_t)(S)*SL;
  tv.cur = a.acc + (size_t)(S + 1) * SL;
  tv.BK = BK;
  tv.KS = KS;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < BK; idx += gridDim.x * blockDim.x) {
    const int b = idx / KS, k = idx - b * KS;
    float o = 0.f, e = 0.f;
    if (k < a.K) {
      // same arithmetic as derive() without the removal of step S (which never runs)
      const float prb = a.Pr_b[b];
      const float* prev_rem_O = tv.prev + BK + KS;
      const float* prev_rem_rs = prev_rem_O + BK;
      o = (tv.ringO[idx] - prev_rem_O[idx]) + tv.cur[idx];
      e = (tv.ringE[idx] - prev_rem_rs[k] * prb) + tv.cur[BK + k] * prb;
    }
    O[idx] = o;
    E[idx] = e;
  }
}

}  // namespace hb
