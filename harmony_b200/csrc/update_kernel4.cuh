// update_kernel4.cuh — persistent kernel for harmony::update_R (harmony.cpp:269-342): every U row is read from
// HBM exactly ONCE per clustering round.
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
// Data movement inside a CTA (persistent, one per SM, cooperative launch):
//   * NP producer warps gather the rows of the CTA's share of every block step with 16-byte cp.async (the LSU
//     gathers 400-byte rows at ~5.3 TB/s; one 1-D bulk/TMA copy per row tops out at 1.3 TB/s, scripts/mb/tma_rows.cu)
//     into a ring of 8-row batches.  Loads do not depend on the step's tables, so the producers run ahead of the
//     consumers by the whole ring (>= one block step at K = 100) and HBM streams without gaps across steps.
//     A batch's arrival is tracked by its slot's mbarrier (cp.async.mbarrier.arrive.noinc: the producers never
//     block on their own loads, the whole ring can be in flight); slots are handed back with monotonic row counts
//     and a consumer looks at a slot's mbarrier only after the producer's monotonic `issued` counter says the slot
//     has entered the use it waits for — a phase parity alone could not tell the wanted refill from the previous one.
//   * NW consumer warps: one row = one warp (lane l owns columns 4(l+32v)..+3), RU rows in flight per warp.  The
//     step's penalty row sum_c P[level_c] lives in registers; column sums accumulate in registers and leave through
//     vector reductions (red.global.add.v4.f32) whenever the next-round block of the rows changes — the plan sorts
//     a CTA's rows by that block, so this happens ~nb times per CTA and step.
// Per step the critical path is: completion counter of step s-1 -> derive the penalty row (K x C table entries,
// L2) -> reduce ~n/(nb*grid) rows from shared memory -> flush K column sums -> release the counter of step s.
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

constexpr int U4_THREADS = 512;
constexpr int U4_NP = 2;                       // producer warps
constexpr int U4_NW = U4_THREADS / 32 - U4_NP;  // consumer warps
constexpr int U4_GT = U4_NW * 32;              // consumer threads
constexpr int U4_BR = 8;                       // rows per ring batch
constexpr int U4_MINBATCH = 16;                // smallest ring worth running (slots)
constexpr int U4_MAXNV = 2;                    // K <= 128 * U4_MAXNV (wider rows leave no room for a ring)

struct Upd4Args {
  const float* U;        // [n][KS]
  float* R;              // [n][KS]
  const int* order;      // [T][n]  rows sorted by (block, tuple, block in the next round, cell)
  const int* next_at;    // [T][n]  block in round t+1 of the cell at each position of `order`
  const int4* ranges;    // [T*nb][grid] (lo, hi, tuple, 0): each CTA's slice of a block lies inside ONE tuple
  const int* tuple_levels;  // [J][C]
  const float* sigma;    // [K]
  const float* theta;    // [B]
  const float* Pr_b;     // [B]
  float* ring;           // [2 parity][2 (O,E)][B][KS]
  float* acc;            // [(S+2)][SL]  slot(s) at (s+1)*SL: [add_O B*KS | add_rs KS | rem_O B*KS | rem_rs KS]
  float* remT;           // [2 parity][nb][J][KS]
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
  int nbatch;            // ring slots (batches of U4_BR rows)
  int ring_rows;         // update_kernel5.cuh: rows of every warp's private ring
  int dbg_flags;         // timing experiments only (HB_U5_FLAGS): 1 = no remT reductions, 2 = no R stores (results invalid)
  int coop;              // 1: one launch covers many steps (counters + in-kernel fold); 0: single-step launch
  long long* dbg;        // optional [steps][8] globaltimer stamps of CTA dbg_cta (null = off)
  int dbg_cta;
};
// Sharded cells on one node (one process per GPU): the per-step sums cross GPUs through peer memory instead of a
// collective per block step.  Every rank owns an exchange area that all ranks of the node have mapped (CUDA IPC
// over NVLink):  inbox[slot][src][XH]  (XH = B KS + KS floats: the add half of an accumulator slot),
// flags[slot][src], and the rank's remT tables.  The CTA that completes the LOCAL add half of a slot copies it into
// entry src = rank of every rank's inbox and then raises the flags (value = epoch of the running cluster_cpp call);
// readers wait for the `world` flags of a slot and add the entries in rank order, so every rank derives bit-identical
// tables.  The next round's removal sums are read from the peers' remT directly when a round is folded.
constexpr int U4_MAXWORLD = 8;
struct Upd4Xch {
  int world = 1, rank = 0;
  unsigned epoch = 0;
  int XH = 0;
  float* inbox = nullptr;       // local area
  unsigned* flags = nullptr;
  float* peer_inbox[U4_MAXWORLD] = {};
  unsigned* peer_flags[U4_MAXWORLD] = {};
  float* peer_remT[U4_MAXWORLD] = {};  // every rank's remT (own entry = local pointer)
};

__device__ __forceinline__ unsigned u4_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void u4_st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// sum over the ranks' entries of element `off` of the add half of accumulator slot index `sl`, in rank order
__device__ __forceinline__ float u4_xsum(const Upd4Xch& x, int sl, int off) {
  const float* q = x.inbox + ((size_t)sl * x.world) * x.XH + off;
  float t = 0.f;
  for (int r = 0; r < x.world; ++r) t += __ldcg(q + (size_t)r * x.XH);
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
// Shared-memory hand-shake words.  Deliberately WITHOUT acquire / release qualifiers: a release here would fence the
// thread's outstanding global reductions and stores (~1 us each time).  What the hand-shakes order is shared-memory
// traffic only: a slot is handed back after the arithmetic that consumed its rows (data dependence), and a batch's
// arrival is observed through its mbarrier (acquire by definition).
__device__ __forceinline__ int u4_ld_volatile(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void u4_st_volatile(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void u4_red_add_shared(int* p, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void u4_mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void u4_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// arrives once all cp.async operations this thread has issued so far have landed
__device__ __forceinline__ void u4_cpasync_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool u4_mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void u4_red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void u4_gsync() { asm volatile("bar.sync 1, %0;" ::"n"(U4_GT) : "memory"); }

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

// shared-memory carve-up: tab[2 KP4] | sig[KP4] | part[NW][KP4] | ring[nbatch BR][KS] | full[nbatch] (mbarriers) |
//   cellid[nbatch BR] | nxt[nbatch BR] | consumed[nbatch] | issued[8]      (KP4 = 128 NV floats)
__host__ __device__ inline size_t upd4_fixed_bytes(int NV, int nbatch) {
  return sizeof(float) * ((size_t)128 * NV * (3 + U4_NW)) + 8 * (size_t)nbatch + sizeof(int) * ((size_t)nbatch * (2 * U4_BR + 1) + 8);
}
__host__ __device__ inline size_t upd4_smem_bytes(int NV, int nbatch, int KS) {
  return upd4_fixed_bytes(NV, nbatch) + sizeof(float) * (size_t)nbatch * U4_BR * KS;
}
__host__ __device__ inline int upd4_nv(int KS) {
  int nv = 1;
  while (128 * nv < KS) nv <<= 1;
  return nv;
}
// ring slots that fit `limit` bytes of shared memory (a multiple of U4_NP: a slot always belongs to the same
// producer); 0 = the kernel cannot run this row width
inline int upd4_nbatch(int KS, size_t limit) {
  const int nv = upd4_nv(KS);
  if (nv > U4_MAXNV) return 0;
  int nbt = 512;
  while (nbt >= U4_MINBATCH && upd4_smem_bytes(nv, nbt, KS) > limit) nbt -= U4_NP;
  return nbt >= U4_MINBATCH ? nbt : 0;
}

template <int NV, bool SIGU>
__global__ void __launch_bounds__(U4_THREADS, 1) k_update_steps4(Upd4Launch lp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Upd4Args& a = lp.a;
  const Upd4Xch& x = lp.x;
  const bool multi = x.world > 1;  // sharded cells with the peer-memory exchange
  constexpr int KP4 = 128 * NV;
  constexpr int RU = (NV <= 2) ? 2 : 1;  // rows in flight per consumer warp
  const int K = a.K, KS = a.KS, C = a.C, J = a.J, B = a.B, nb = a.nb;
  const int KS4 = KS >> 2;
  const int BK = B * KS;
  const int SL = 2 * (BK + KS);
  const int NBT = a.nbatch;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, grid = gridDim.x;

  float* tab = reinterpret_cast<float*>(smem_raw);         // Psum | log Psum
  float* sig = tab + 2 * KP4;
  float* part = sig + KP4;                                  // [NW][KP4]
  float* ringbuf = part + (size_t)U4_NW * KP4;              // [NBT][BR][KS]
  uint64_t* full = reinterpret_cast<uint64_t*>(ringbuf + (size_t)NBT * U4_BR * KS);  // [NBT] one phase per use of the slot
  int* cellid = reinterpret_cast<int*>(full + NBT);
  int* nxt = cellid + (size_t)NBT * U4_BR;
  int* consumed = nxt + (size_t)NBT * U4_BR;                // rows handed back per slot, cumulative
  int* issued_w = consumed + NBT;                           // batches issued per producer warp, cumulative
  __shared__ double sh_obj[2];
  __shared__ int sh_last;

  // stale ring rows are read (with weight 0) by the tail of a row group: they must be finite
  for (float* q = ringbuf + tid; q < ringbuf + (size_t)NBT * U4_BR * KS; q += U4_THREADS) *q = 0.f;
  for (int i = tid; i < KP4; i += U4_THREADS) sig[i] = (i < K) ? a.sigma[i] : 0.f;
  for (int i = tid; i < 2 * KP4; i += U4_THREADS) tab[i] = 0.f;
  for (int i = tid; i < NBT; i += U4_THREADS) consumed[i] = 0;
  if (tid < 8) issued_w[tid] = 0;
  for (int i = tid; i < NBT; i += U4_THREADS) u4_mbar_init(full + i, 33);  // 32 cp.async arrivals + the meta writer
  if (tid == 0) {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sh_obj[0] = 0.0;
    sh_obj[1] = 0.0;
  }
  __syncthreads();

  auto range_of = [&](int s, int& lo, int& n) {
    const int4 rg = __ldg(a.ranges + (size_t)s * grid + cta);
    lo = rg.x;
    n = rg.y > rg.x ? rg.y - rg.x : 0;
    return rg.z;
  };

  if (warp < U4_NP) {
    // ======================================= producers =======================================
    const int p = warp;
    bool lane_ok[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) lane_ok[v] = (lane + 32 * v) < KS4;
    int issued = 0;  // batches of this warp
    int gb0 = 0;     // global batch index of the first batch of the step
    int slot = p % NBT, need = 0;  // ring slot of this warp's next batch (global batch p, p + NP, ..) / rows its earlier uses owe
    for (int s = a.s_begin; s < a.s_end; ++s) {
      int lo, n;
      range_of(s, lo, n);
      const int t = s / nb;
      const int* order = a.order + (size_t)t * a.n + lo;
      const int* next_at = a.next_at + (size_t)t * a.n + lo;
      const int nbt = (n + U4_BR - 1) / U4_BR;
      int b = ((p - gb0) % U4_NP + U4_NP) % U4_NP;  // first batch of the step that is this warp's
      // software pipeline over the warp's batches: the plan entries of batch b + NP are loaded while b is issued
      int cell_n = 0, nx_n = 0;
      if (b < nbt && lane < U4_BR && b * U4_BR + lane < n) {
        cell_n = __ldg(order + b * U4_BR + lane);
        nx_n = __ldg(next_at + b * U4_BR + lane);
      }
      for (; b < nbt; b += U4_NP) {
        const int cell = cell_n, nx = nx_n;
        const int bn = b + U4_NP;
        if (bn < nbt && lane < U4_BR && bn * U4_BR + lane < n) {
          cell_n = __ldg(order + bn * U4_BR + lane);
          nx_n = __ldg(next_at + bn * U4_BR + lane);
        }
        while (u4_ld_volatile(consumed + slot) < need) __nanosleep(32);  // every earlier use handed back U4_BR rows
        const int nr = min(U4_BR, n - b * U4_BR);
        float* dst = ringbuf + (size_t)slot * U4_BR * KS;
#pragma unroll
        for (int r = 0; r < U4_BR; ++r) {
          const int cr = __shfl_sync(0xffffffffu, cell, r);
          if (r < nr) {
            const float* src = a.U + (size_t)cr * KS;
#pragma unroll
            for (int v = 0; v < NV; ++v)
              if (lane_ok[v]) {
                const unsigned sp = (unsigned)__cvta_generic_to_shared(dst + (size_t)r * KS + 4 * (lane + 32 * v));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp), "l"(src + 4 * (lane + 32 * v)) : "memory");
              }
          }
        }
        u4_cpasync_arrive(full + slot);
        if (lane < U4_BR) {
          cellid[slot * U4_BR + lane] = cell;
          nxt[slot * U4_BR + lane] = nx;
        }
        __syncwarp();
        ++issued;
        if (lane == 0) {
          u4_mbar_arrive(full + slot);          // release: the meta words above are visible with the phase
          u4_st_volatile(issued_w + p, issued);  // the slot has entered this use: its parity may be looked at
        }
        slot += U4_NP;                           // this warp's next batch is global batch + NP
        if (slot >= NBT) {
          slot -= NBT;
          need += U4_BR;
        }
      }
      gb0 += nbt;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }

  // ======================================= consumers =======================================
  const int gw = warp - U4_NP;      // consumer warp
  const int gt = gw * 32 + lane;    // consumer thread
  const int S_total = a.T * nb;
  unsigned* cntU = a.bar + 1;               // cntU[s], s >= -1
  unsigned* cntF = a.bar + (S_total + 2);   // cntF[t]
  bool lane_ok[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) lane_ok[v] = (lane + 32 * v) < KS4;

  // signal(c, sl): this CTA is done with the step counted by c.  Sharded cells: the CTA that completes the count
  // publishes the finished add half of accumulator slot index sl to every rank (sl < 0: nothing to publish).
  auto signal = [&](unsigned* c, int sl) {
    u4_gsync();
    if (!multi || sl < 0) {
      if (gt == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
      }
      return;
    }
    if (gt == 0) {
      __threadfence();
      const unsigned old = atomicAdd(c, 1u);
      __threadfence();
      sh_last = (old == (unsigned)grid - 1u) ? 1 : 0;
    }
    u4_gsync();
    if (sh_last) {
      const float* src = a.acc + (size_t)sl * SL;  // complete: every CTA's atomics preceded its count
      const size_t entry = (size_t)sl * x.world + x.rank;
      for (int r = 0; r < x.world; ++r) {
        float* dst = x.peer_inbox[r] + entry * x.XH;
        for (int i = gt * 4; i < x.XH; i += U4_GT * 4)
          *reinterpret_cast<float4*>(dst + i) = __ldcg(reinterpret_cast<const float4*>(src + i));
      }
      __threadfence_system();
      u4_gsync();
      if (gt < x.world) u4_st_release_sys(x.peer_flags[gt] + entry, x.epoch);
    }
  };
  auto wait_for = [&](const unsigned* c) {
    if (gt == 0) {
      while (u4_ld_acquire_gpu(c) < (unsigned)grid) __nanosleep(20);
      __threadfence();
    }
    u4_gsync();
  };
  // sharded cells: every rank's entry of the add half of slot index sl has arrived in the local inbox
  auto wait_slot = [&](int sl) {
    if (gt < x.world) {
      const unsigned* f = x.flags + (size_t)sl * x.world + gt;
      while (u4_ld_acquire_sys(f) != x.epoch) __nanosleep(40);
      __threadfence();
    }
    u4_gsync();
  };
  auto stamp = [&](int s, int slot_id) {
    if (a.dbg && cta == a.dbg_cta && gt == 0) {
      long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      a.dbg[(size_t)(s - a.s_begin) * 8 + slot_id] = tns;
    }
  };
  auto tables_for = [&](int s) {
    U4Tables tv;
    const int par = (s - 1) & 1;
    tv.ringO = a.ring + (size_t)par * 2 * BK;
    tv.ringE = tv.ringO + BK;
    tv.prev = a.acc + (size_t)(s)*SL;       // slot(s-1)
    tv.cur = a.acc + (size_t)(s + 1) * SL;  // slot(s)
    tv.BK = BK;
    tv.KS = KS;
    return tv;
  };
  // removal sums of round t: remT[t & 1][j][q][:] (filed during round t-1; sharded cells: summed over the ranks'
  // tables in rank order) -> rem halves of slot(t nb + j); each (j, k) column is owned by one thread of the grid, which
  // also clears the other parity for round t (every rank has read it: all of them have finished round t-1)
  auto fold_round = [&](int t) {
    const size_t par_off = (size_t)(t & 1) * nb * J * KS;
    float* Tz = a.remT + (size_t)((t + 1) & 1) * nb * J * KS;
    for (int item = cta * U4_GT + gt; item < nb * K; item += grid * U4_GT) {
      const int j = item / K, k = item - j * K;
      const size_t joff = par_off + (size_t)j * J * KS + k;
      float* slot = a.acc + (size_t)(t * nb + j + 1) * SL;
      float* rem_O = slot + BK + KS;
      float* rem_rs = rem_O + BK;
      float rs = 0.f;
      for (int q = 0; q < J; ++q) {
        float v;
        if (multi) {
          v = 0.f;
          for (int r = 0; r < x.world; ++r) v += __ldcg(x.peer_remT[r] + joff + (size_t)q * KS);
        } else {
          v = __ldcg(a.remT + joff + (size_t)q * KS);
        }
        if (v != 0.f) {
          rs += v;
          for (int c = 0; c < C; ++c) {
            float* o = rem_O + (size_t)__ldg(a.tuple_levels + q * C + c) * KS + k;
            *o = *o + v;
          }
        }
        Tz[((size_t)j * J + q) * KS + k] = 0.f;
      }
      rem_rs[k] = rs;
    }
  };

  float okd = 0.f, oent = 0.f;
  auto flush_objective = [&](int t) {
    okd = warp_sum(okd);
    oent = warp_sum(oent);
    if (lane == 0) {
      atomicAdd(&sh_obj[0], (double)okd);
      atomicAdd(&sh_obj[1], (double)oent);
    }
    u4_gsync();
    if (gt == 0) {
      atomicAdd(a.obj + 2 * t + 0, sh_obj[0]);
      atomicAdd(a.obj + 2 * t + 1, sh_obj[1]);
      sh_obj[0] = 0.0;
      sh_obj[1] = 0.0;
    }
    u4_gsync();
    okd = 0.f;
    oent = 0.f;
  };

  int gb0 = 0;  // global batch index of the first batch of the step (same sequence as the producers')
  for (int s = a.s_begin; s < a.s_end; ++s) {
    stamp(s, 0);
    int lo, n;
    const int q = range_of(s, lo, n);
    const int t = s / nb, j = s - t * nb;
    const int nbt = (n + U4_BR - 1) / U4_BR;
    const bool writeR = t >= a.write_from;
    const bool has_next = t < a.has_next_from;
    if (a.coop) {
      if (s > a.s_begin) wait_for(cntU + s - 1);  // add_{s-1}, ring(s-1); at j == 0 also: round t-1 is complete
      if (multi && s >= 1) wait_slot(s + 1);      // ... on every rank: add_{s-1} lives in slot(s) = index s + 1
      if (j == 0 && t > 0) {
        fold_round(t);
        signal(cntF + t, -1);
        wait_for(cntF + t);
      }
    }
    const Upd4Xch* xp = multi ? &x : nullptr;
    const int xs = (s >= 1) ? s + 1 : -1;
    stamp(s, 1);
    // ---- tables of the step: the penalty row of this CTA's tuple, and this CTA's share of O_s, E_s ----
    {
      const U4Tables tv = tables_for(s);
      if (n > 0) {
        for (int k = gt; k < K; k += U4_GT) {
          float v = 0.f;
          for (int c = 0; c < C; ++c) {
            float o, e, pp;
            u4_derive(tv, a.Pr_b, a.theta, __ldg(a.tuple_levels + q * C + c), k, o, e, pp, xp, xs);
            v += pp;
          }
          tab[k] = v;
          tab[KP4 + k] = fast_log(v);
        }
      }
      float* outO = a.ring + (size_t)(s & 1) * 2 * BK;
      float* outE = outO + BK;
      for (int idx = cta + grid * (U4_GT - 1 - gt); idx < BK; idx += grid * U4_GT) {  // the last threads first: they idle above
        const int b = idx / KS, k = idx - b * KS;
        float o = 0.f, e = 0.f, pp = 0.f;
        if (k < K) u4_derive(tv, a.Pr_b, a.theta, b, k, o, e, pp, xp, xs);
        outO[idx] = o;
        outE[idx] = e;
        if (j == 0 && t > 0) {
          float* oe = a.OEend + (size_t)(t - 1) * 2 * BK;
          oe[idx] = o;
          oe[BK + idx] = e;
        }
      }
    }
    u4_gsync();
    stamp(s, 2);
    if (n > 0) {
      float4 pP[NV], pL[NV], sg[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        pP[v] = *reinterpret_cast<const float4*>(tab + 4 * (lane + 32 * v));
        pL[v] = *reinterpret_cast<const float4*>(tab + KP4 + 4 * (lane + 32 * v));
        sg[v] = *reinterpret_cast<const float4*>(sig + 4 * (lane + 32 * v));
      }
      float4 cs[NV], cs2[NV];  // column sums of the step / of the rows filed under next-round block cur_nb
#pragma unroll
      for (int v = 0; v < NV; ++v) cs[v] = cs2[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      int cur_nb = -1;
      float* remT_next = a.remT + (size_t)((t + 1) & 1) * nb * J * KS + (size_t)q * KS;
      auto flush_next = [&]() {
        if (cur_nb >= 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (has_next && lane_ok[v]) u4_red_add_v4(remT_next + (size_t)cur_nb * J * KS + 4 * (lane + 32 * v), cs2[v]);
            cs[v].x += cs2[v].x;
            cs[v].y += cs2[v].y;
            cs[v].z += cs2[v].z;
            cs[v].w += cs2[v].w;
            cs2[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      // this warp's rows of the step: a contiguous run [r0, r1), walked batch by batch, RU rows per iteration
      int c_rows = (n + U4_NW - 1) / U4_NW;
      c_rows = (c_rows + RU - 1) / RU * RU;   // even starts: a row pair never straddles two batches
      const int r0 = gw * c_rows, r1 = min(n, r0 + c_rows);
      if (r0 < r1) {
        int b = r0 / U4_BR;                    // batch of the step
        int gbl = gb0 + b;                     // global batch index
        int slot = gbl % NBT;
        unsigned par = (unsigned)(gbl / NBT) & 1u;
        bool first = true;
        for (int rb = r0; rb < r1;) {
          const int bend = min(r1, (b + 1) * U4_BR);
          {  // the batch has entered this use of its slot (issued), then: its rows have landed (mbarrier phase)
            const int* iw = issued_w + (gbl % U4_NP);
            const int want = gbl / U4_NP + 1;
            // back off between polls: a spinning consumer warp takes issue slots and LSU queue entries away from the
            // producer warp that shares its scheduler (measured: 14 spinning warps cut the gather rate to a fifth)
            while (u4_ld_volatile(iw) < want) __nanosleep(64);
            while (!u4_mbar_try_wait(full + slot, par)) {  // try_wait suspends the warp in hardware
            }
          }
          if (first) {
            stamp(s, 3);
            first = false;
          }
          const float* bbase = ringbuf + (size_t)slot * U4_BR * KS;
          const int* bcell = cellid + slot * U4_BR;
          const int* bnxt = nxt + slot * U4_BR;
          const int handed = bend - rb + ((bend == n) ? nbt * U4_BR - n : 0);  // + the step's padding rows
          for (int k = rb - b * U4_BR; rb < bend; rb += RU, k += RU) {
            float4 u[RU][NV], e[RU][NV];
            float ssum[RU], Aacc[RU], Bacc[RU], Sacc[RU];
            int cellr[RU], nbr[RU];
            bool valid[RU];
#pragma unroll
            for (int i = 0; i < RU; ++i) {
              valid[i] = (i == 0) || (rb + i < bend);
              const int kk = valid[i] ? k + i : k;
              cellr[i] = bcell[kk];
              nbr[i] = bnxt[kk];
              const float* rp = bbase + (size_t)kk * KS;
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                u[i][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane_ok[v]) u[i][v] = *reinterpret_cast<const float4*>(rp + 4 * (lane + 32 * v));
              }
            }
#pragma unroll
            for (int i = 0; i < RU; ++i) {
              ssum[i] = Aacc[i] = Bacc[i] = Sacc[i] = 0.f;
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                const float uu[4] = {u[i][v].x, u[i][v].y, u[i][v].z, u[i][v].w};
                const float pp[4] = {pP[v].x, pP[v].y, pP[v].z, pP[v].w};
                const float ll[4] = {pL[v].x, pL[v].y, pL[v].z, pL[v].w};
                const float ss[4] = {sg[v].x, sg[v].y, sg[v].z, sg[v].w};
                float ee[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  ee[c] = fast_exp(uu[c]) * pp[c];  // un-normalised R (>= 0; exactly 0 in the padding columns)
                  if (SIGU) {
                    Aacc[i] = fmaf(ee[c], uu[c], Aacc[i]);
                    Bacc[i] = fmaf(ee[c], ll[c], Bacc[i]);
                  } else {
                    const float tt = ss[c] * ee[c];
                    Aacc[i] = fmaf(tt, uu[c], Aacc[i]);
                    Bacc[i] = fmaf(tt, ll[c], Bacc[i]);
                    Sacc[i] += tt;
                  }
                }
                ssum[i] += (ee[0] + ee[1]) + (ee[2] + ee[3]);
                e[i][v] = make_float4(ee[0], ee[1], ee[2], ee[3]);
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
              for (int i = 0; i < RU; ++i) ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], o);
#pragma unroll
            for (int i = 0; i < RU; ++i) {
              const float sdiv = (ssum[i] == 0.f) ? 1.f : ssum[i];  // arma::normalise(.., 1, 0): zero norm divides by 1
              const float inv = valid[i] ? fast_rcp(sdiv) : 0.f;
              if (valid[i] && nbr[i] != cur_nb) {  // warp-uniform
                flush_next();
                cur_nb = nbr[i];
              }
              float* rp = a.R + (size_t)cellr[i] * KS;
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                float4 rr;
                rr.x = e[i][v].x * inv;
                rr.y = e[i][v].y * inv;
                rr.z = e[i][v].z * inv;
                rr.w = e[i][v].w * inv;
                cs2[v].x += rr.x;
                cs2[v].y += rr.y;
                cs2[v].z += rr.z;
                cs2[v].w += rr.w;
                if (writeR && valid[i] && lane_ok[v]) *reinterpret_cast<float4*>(rp + 4 * (lane + 32 * v)) = rr;
              }
              // sum_k R dist = -sum sigma R U ;  sum_k sigma R log R = sum sigma R (U + log Psum - log s)
              const float ls = fast_log(sdiv);
              if (SIGU) {
                const float srow = (lane == 0) ? ssum[i] : 0.f;  // the row total once per row
                const float w = a.sigma0 * inv;
                okd = fmaf(-w, Aacc[i], okd);
                oent = fmaf(w, (Aacc[i] + Bacc[i]) - ls * srow, oent);
              } else {
                okd = fmaf(-inv, Aacc[i], okd);
                oent = fmaf(inv, (Aacc[i] + Bacc[i]) - ls * Sacc[i], oent);
              }
            }
          }
          // hand the rows of this batch back to the producers
          __syncwarp();
          if (lane == 0) u4_red_add_shared(consumed + slot, handed);
          ++b;
          ++gbl;
          if (++slot == NBT) {
            slot = 0;
            par ^= 1u;
          }
        }
      }
      flush_next();
      stamp(s, 4);
      // ---- add_s: this CTA's column sums -> slot(s+1) ----
#pragma unroll
      for (int v = 0; v < NV; ++v) *reinterpret_cast<float4*>(part + (size_t)gw * KP4 + 4 * (lane + 32 * v)) = cs[v];
      u4_gsync();
      float* nslot = a.acc + (size_t)(s + 2) * SL;
      for (int k = gt; k < K; k += U4_GT) {
        float tsum = 0.f;
#pragma unroll
        for (int w = 0; w < U4_NW; ++w) tsum += part[(size_t)w * KP4 + k];
        atomicAdd(nslot + BK + k, tsum);
        for (int c = 0; c < C; ++c) atomicAdd(nslot + (size_t)__ldg(a.tuple_levels + q * C + c) * KS + k, tsum);
      }
      stamp(s, 5);
    }
    if ((s + 1) % nb == 0) flush_objective(t);
    stamp(s, 6);
    if (a.coop) signal(cntU + s, s + 2);  // add_s lives in slot(s + 1) = index s + 2
    gb0 += nbt;
  }
  if (a.s_end % nb != 0 && a.s_end > a.s_begin) flush_objective((a.s_end - 1) / nb);  // partial round
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

// Stand-alone fold of remT into the rem halves of round t's slots (per-step launches: sharded cells without the
// peer exchange, where remT is all-reduced by the host in between).  Same arithmetic as the in-kernel fold.
__global__ void k_fold_round(Upd4Args a, int t) {
  const int K = a.K, KS = a.KS, C = a.C, J = a.J, nb = a.nb;
  const int BK = a.B * KS, SL = 2 * (BK + KS);
  const float* T0 = a.remT + (size_t)(t & 1) * nb * J * KS;
  float* Tz = a.remT + (size_t)((t + 1) & 1) * nb * J * KS;
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < nb * K; item += gridDim.x * blockDim.x) {
    const int j = item / K, k = item - j * K;
    float* slot = a.acc + (size_t)(t * nb + j + 1) * SL;
    float* rem_O = slot + BK + KS;
    float* rem_rs = rem_O + BK;
    float rs = 0.f;
    for (int q = 0; q < J; ++q) {
      const float v = T0[((size_t)j * J + q) * KS + k];
      if (v != 0.f) {
        rs += v;
        for (int c = 0; c < C; ++c) rem_O[(size_t)a.tuple_levels[q * C + c] * KS + k] += v;
      }
      Tz[((size_t)j * J + q) * KS + k] = 0.f;
    }
    rem_rs[k] = rs;
  }
}

// After the last executed step S: O = O_S, E = E_S into the handle's tables.
__global__ void k_update_finalize4(Upd4Args a, Upd4Xch x, int S, float* __restrict__ O, float* __restrict__ E) {
  const int KS = a.KS, BK = a.B * KS, SL = 2 * (BK + KS);
  const bool multi = x.world > 1;
  if (multi && S >= 1) {  // the other ranks' kernels may still be running: wait for their last add halves
    if ((int)threadIdx.x < x.world) {
      const unsigned* f = x.flags + (size_t)(S + 1) * x.world + threadIdx.x;
      while (u4_ld_acquire_sys(f) != x.epoch) __nanosleep(40);
      __threadfence();
    }
    __syncthreads();
  }
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
