// update_kernel3.cuh — second generation of the persistent update_R kernel (harmony.cpp:269-342).
//
// EXPERIMENTAL: selected only with HB_UPDATE_V3=1 (and tuple-aligned CTA ranges); the default path is
// k_update_steps (update_kernel.cuh).  Same algorithm, same global data structures (UpdArgs: per-step
// accumulator slots, completion counters, ring / Psave / OEend tables) — what changes is how a CTA moves and
// reduces its rows.  k_update_steps is issue-bound (1024 threads -> 64 registers, per-lane address arithmetic
// and cp.async issue in the consumer loop: ~120 instructions per row and group); here
//   * two producer warps (one per warp group) gather the rows with one 1-D bulk (TMA) copy per row into
//     shared-memory stage rings guarded by full/empty mbarriers, and run ahead across block steps — the
//     pipeline is already full when a step's tables arrive;
//   * 8 + 6 consumer warps with 128 registers reduce the staged rows: LPR lanes per row and NV float4 per lane
//     chosen so that LPR * NV covers the row exactly when possible (K = 100: 5 lanes x 5 float4, 6 rows per
//     warp iteration), loop-invariant penalty row in registers, no per-lane global addressing;
//   * R (rounds that store it) is normalised in place in the stage and leaves through bulk stores.
// All rows of a CTA in one step belong to one covariate tuple (UpdArgs::ranges), so the column sums stay in
// registers for the whole step and are flushed once.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "umma.cuh"
#include "update_kernel.cuh"

namespace hb {

constexpr int U3_THREADS = 512;
constexpr int U3_NWU = 8;    // update consumer warps (warps 2 .. 9)
constexpr int U3_NWL = 6;    // look-ahead consumer warps (warps 10 .. 15)
constexpr int U3_MAXD = 16;  // maximal ring depth
constexpr int U3_MAXNV = 5;

struct Upd3Geom {
  int NV;    // float4 per lane
  int LPR;   // lanes per row
  int RPI;   // rows per warp iteration = 32 / LPR
  int IT;    // warp iterations per stage
  int SR;    // rows per stage = RPI * IT (<= 32: one producer lane per row)
  int KP;    // padded row length in floats = 4 * NV * LPR (>= KS)
  int DU, DL;  // ring depths (stages) of the update / look-ahead group
};

// Multi-GPU step exchange through peer memory (one process per GPU, cells sharded).  Every rank
// owns an exchange area that all ranks of the node have mapped (CUDA IPC over NVLink):
//   data[((slot * 2 + half) * world + src) * XH + i]     flags[(slot * 2 + half) * world + src]
// slot = s + 1 for the accumulator slot of step s, half 0 = add_{s-1}, half 1 = rem_s (XH = B KS + KS floats).
// The CTA that completes a half of the LOCAL accumulator slot copies it into entry `src = rank` of every
// rank's area and then raises the flags (value = epoch of the running cluster_cpp call).  Readers wait for
// the `world` flags of a half and add the partial sums in rank order, so every rank derives bit-identical
// tables without a collective launch between block steps.
constexpr int U3_MAXWORLD = 8;
struct Upd3Xch {
  int world = 1, rank = 0;
  unsigned epoch = 0;
  int XH = 0;
  float* local_data = nullptr;
  unsigned* local_flag = nullptr;
  float* data[U3_MAXWORLD] = {};
  unsigned* flag[U3_MAXWORLD] = {};
};

struct Upd3Args {
  UpdArgs a;
  Upd3Geom g;
  Upd3Xch x;
  int l2_hints = 1;  // look-ahead rows are loaded evict_last (the update group re-reads them a few steps later),
                     // the update group's own loads evict_first (last use of the row in this round)
};

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// sum over the ranks' partial sums of element `off` of (slot, half), in rank order
__device__ __forceinline__ float u3_xsum(const Upd3Xch& x, int slot, int half, int off) {
  const float* q = x.local_data + ((size_t)(slot * 2 + half) * x.world) * x.XH + off;
  float t = 0.f;
  for (int r = 0; r < x.world; ++r) t += __ldcg(q + (size_t)r * x.XH);
  return t;
}
// derive() for sharded cells: the accumulator terms are sums over the ranks' exchange entries.  s is the
// step within the call: slot(s - 1) and the add half of slot(s) do not exist for s == 0 (zero).
__device__ __forceinline__ void derive_x(const Upd3Xch& x, const TableView& tv, const float* Pr_b, const float* theta, int s,
                                         int b, int k, bool with_removal, float& o, float& e, float& p) {
  const int idx = b * tv.KS + k;
  const float prb = Pr_b[b];
  float prev_rem_O = 0.f, prev_rem_rs = 0.f, cur_add_O = 0.f, cur_add_rs = 0.f;
  if (s >= 1) {
    prev_rem_O = u3_xsum(x, s, 1, idx);
    prev_rem_rs = u3_xsum(x, s, 1, tv.BK + k);
    cur_add_O = u3_xsum(x, s + 1, 0, idx);
    cur_add_rs = u3_xsum(x, s + 1, 0, tv.BK + k);
  }
  o = (__ldcg(tv.ringO + idx) - prev_rem_O) + cur_add_O;
  e = (__ldcg(tv.ringE + idx) - prev_rem_rs * prb) + cur_add_rs * prb;
  p = 0.f;
  if (with_removal) {
    const float e_eff = e - u3_xsum(x, s + 1, 1, tv.BK + k) * prb;
    const float o_eff = o - u3_xsum(x, s + 1, 1, idx);
    p = penalty_pow(o_eff, e_eff, theta[b]);
  }
}

// shared-memory carve-up (floats): sig[KP] | tabU[2 KP] | tabL[(nb+1) KP] | partU[NWU KP] | partL[NWL KP] |
//   metaU[DU 32] | metaL[DL 32] | mbarriers[2 DU + 2 DL] (8 B each) | ringU[DU SR KP] | ringL[DL SR KP]
__host__ __device__ inline size_t upd3_fixed_floats(const Upd3Geom& g, int nb) {
  return (size_t)g.KP * (1 + 2 + (nb + 1) + U3_NWU + U3_NWL) + 32 * (size_t)(g.DU + g.DL) + 2 * 2 * (size_t)(g.DU + g.DL);
}
__host__ __device__ inline size_t upd3_smem_bytes(const Upd3Geom& g, int nb) {
  return sizeof(float) * (upd3_fixed_floats(g, nb) + (size_t)(g.DU + g.DL) * g.SR * g.KP);
}

// Chooses the lane / stage geometry for a row of KS floats; false if the kernel cannot run this shape.
inline bool upd3_geometry(int KS, int nb, size_t smem_limit, Upd3Geom* out) {
  const int KS4 = KS >> 2;
  double best = 0.0;
  Upd3Geom g{};
  for (int nv = 1; nv <= U3_MAXNV; ++nv)
    for (int lpr = 1; lpr <= 32; ++lpr) {
      if (nv * lpr < KS4) continue;
      const int rpi = 32 / lpr;
      const double eff = ((double)KS4 / (nv * lpr)) * ((double)(rpi * lpr) / 32.0);
      if (eff > best + 1e-9 || (eff > best - 1e-9 && nv > g.NV)) {
        best = eff;
        g.NV = nv;
        g.LPR = lpr;
        g.RPI = rpi;
      }
      break;  // larger lpr for this nv only wastes more lanes
    }
  if (best == 0.0) return false;
  g.KP = 4 * g.NV * g.LPR;
  const int it_max = 32 / g.RPI;
  int it = (int)(16384 / ((size_t)g.RPI * g.KP * sizeof(float)));
  g.IT = it < 1 ? 1 : (it > it_max ? it_max : it);
  g.SR = g.RPI * g.IT;
  g.DU = g.DL = 2;
  if (upd3_smem_bytes(g, nb) > smem_limit) return false;
  while (g.DU < U3_MAXD) {
    Upd3Geom t = g;
    t.DU = t.DL = g.DU + 1;
    if (upd3_smem_bytes(t, nb) > smem_limit) break;
    g = t;
  }
  *out = g;
  return true;
}

__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Total of x over the LPR consecutive lanes of a row group; every lane of the group gets it.  Branch-free
// tree: level l adds the value 2^l lanes up with weight m[l] = (gl + 2^l < LPR) ? 1 : 0 (precomputed), the
// group leader (gl == 0) then holds the total and broadcasts it.
struct RowTree {
  float m[5];
  int leader;
};
__device__ __forceinline__ RowTree make_row_tree(int LPR, int gl, int lane) {
  RowTree t;
#pragma unroll
  for (int l = 0; l < 5; ++l) t.m[l] = (gl + (1 << l) < LPR) ? 1.f : 0.f;
  t.leader = lane - gl;
  return t;
}
__device__ __forceinline__ float u3_row_total(float x, const RowTree& t) {
#pragma unroll
  for (int l = 0; l < 5; ++l) x = fmaf(__shfl_down_sync(0xffffffffu, x, 1u << l), t.m[l], x);
  return __shfl_sync(0xffffffffu, x, t.leader);
}

template <int NV>
__global__ void __launch_bounds__(U3_THREADS, 1) k_update_steps3(Upd3Args p) {
  extern __shared__ __align__(16) float smem[];
  const UpdArgs& a = p.a;
  const Upd3Xch& x = p.x;
  const bool multi = x.world > 1;
  const int K = a.K, KS = a.KS, C = a.C, B = a.B, nb = a.nb;
  const int LPR = p.g.LPR, RPI = p.g.RPI, SR = p.g.SR, KP = p.g.KP, DU = p.g.DU, DL = p.g.DL;
  const int BK = B * KS;
  const int SL = 2 * (BK + KS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, grid = gridDim.x;

  float* sig = smem;
  float* tabU = sig + KP;                         // Psum | log Psum
  float* tabL = tabU + 2 * KP;                    // [nb + 1][KP]; row nb stays zero
  float* partU = tabL + (size_t)(nb + 1) * KP;    // [NWU][KP]
  float* partL = partU + (size_t)U3_NWU * KP;     // [NWL][KP]
  int* metaU = reinterpret_cast<int*>(partL + (size_t)U3_NWL * KP);  // [DU][32] cell of each staged row
  int* metaL = metaU + 32 * DU;                                      // [DL][32] previous-round block of each staged row
  uint64_t* fullU = reinterpret_cast<uint64_t*>(metaL + 32 * DL);
  uint64_t* emptyU = fullU + DU;
  uint64_t* fullL = emptyU + DU;
  uint64_t* emptyL = fullL + DL;
  float* ringU = reinterpret_cast<float*>(emptyL + DL);
  float* ringL = ringU + (size_t)DU * SR * KP;
  __shared__ double sh_obj[2];
  __shared__ int sh_last[2];  // per consumer group: this CTA completed the half it just signalled
  __shared__ int sh_tick[2][4];  // per group and step (mod 4): next stage of the step to hand to a consumer warp

  // ---- one-time initialisation: zero everything that is read before it is written (stale stage rows are
  // consumed with weight 0 and must be finite), tables' padding columns stay 0 for the whole kernel
  {
    float* endp = ringL + (size_t)DL * SR * KP;
    for (float* q = smem + tid; q < endp; q += U3_THREADS) *q = 0.f;
  }
  __syncthreads();
  for (int k = tid; k < KS; k += U3_THREADS) sig[k] = (k < K) ? a.sigma[k] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < DU; ++i) {
      umma::mbar_init(fullU + i, 1);
      umma::mbar_init(emptyU + i, 1);
    }
    for (int i = 0; i < DL; ++i) {
      umma::mbar_init(fullL + i, 1);
      umma::mbar_init(emptyL + i, 1);
    }
    umma::fence_barrier_init();
    sh_obj[0] = 0.0;
    sh_obj[1] = 0.0;
    for (int i = 0; i < 4; ++i) sh_tick[0][i] = sh_tick[1][i] = 0;
  }
  __syncthreads();

  const int S_total = a.T * nb;
  unsigned* cntU = a.bar + 1;  // cntU[s], s >= -1
  unsigned* cntL = a.bar + (S_total + 2) + 1;
  const int look_first = a.prologue ? a.s_begin : a.s_begin + 1;
  const int look_last = (a.s_end < S_total) ? a.s_end : S_total - 1;
  const uint32_t row_bytes = (uint32_t)KS * 4u;

  auto range_of = [&](int s, int& lo, int& hi, int& q) {
    const int4 rg = __ldg(a.ranges + (size_t)s * grid + cta);
    lo = rg.x;
    hi = rg.y > rg.x ? rg.y : rg.x;
    q = rg.z;
  };
  auto tables_for = [&](int s) {
    TableView tv;
    const int par = (s - 1) & 1;
    tv.ringO = a.ring + (size_t)par * 2 * BK;
    tv.ringE = tv.ringO + BK;
    tv.prev = a.acc + (size_t)(s)*SL;       // slot(s-1)
    tv.cur = a.acc + (size_t)(s + 1) * SL;  // slot(s)
    tv.BK = BK;
    tv.KS = KS;
    return tv;
  };

  if (warp < 2) {
    // =========================== producers: warp 0 feeds ringU, warp 1 feeds ringL ===========================
    const bool forU = (warp == 0);
    const int D = forU ? DU : DL;
    uint64_t* full = forU ? fullU : fullL;
    uint64_t* empty = forU ? emptyU : emptyL;
    float* ring = forU ? ringU : ringL;
    int* meta = forU ? metaU : metaL;
    const int s0 = forU ? a.s_begin : look_first;
    const int s1 = forU ? a.s_end - 1 : look_last;
    int slot = 0, use = 0;  // stage g lives in slot g % D; use = g / D
    const uint64_t policy = forU ? umma::l2_policy_evict_first() : umma::l2_policy_evict_last();
    for (int s = s0; s <= s1; ++s) {
      int lo, hi, q;
      range_of(s, lo, hi, q);
      const int t = s / nb;
      const int* order = a.order + (size_t)t * a.n;
      const int* prev_at = a.prev_at + (size_t)t * a.n;
      const float* src = (!forU && s < nb && a.first_round_from_R) ? a.R : a.U;
      for (int r0 = lo; r0 < hi; r0 += SR) {
        if (use >= 1) umma::mbar_wait(empty + slot, (use - 1) & 1);
        const int nr = (hi - r0 < SR) ? hi - r0 : SR;
        const bool valid = lane < nr;
        int cell = 0;
        if (valid) {
          cell = __ldg(order + r0 + lane);
          meta[slot * 32 + lane] = forU ? cell : ((t > 0) ? __ldg(prev_at + r0 + lane) : 0);
        }
        __syncwarp();
        if (lane == 0) umma::mbar_arrive_expect_tx(full + slot, (uint32_t)nr * row_bytes);
        __syncwarp();
        if (valid) {
          float* dstp = ring + ((size_t)slot * SR + lane) * KP;
          if (p.l2_hints)
            umma::bulk_load_hint(dstp, src + (size_t)cell * KS, row_bytes, full + slot, policy);
          else
            umma::bulk_load(dstp, src + (size_t)cell * KS, row_bytes, full + slot);
        }
        if (++slot == D) {
          slot = 0;
          ++use;
        }
      }
    }
    return;
  }

  // =========================== consumers ===========================
  const bool isU = warp < 2 + U3_NWU;
  const int gw = isU ? warp - 2 : warp - 2 - U3_NWU;     // warp within the group
  const int GT = (isU ? U3_NWU : U3_NWL) * 32;           // threads of the group
  const int gt = gw * 32 + lane;
  const int bar_id = isU ? 1 : 2;
  const int NW = isU ? U3_NWU : U3_NWL;
  const int rg = lane / LPR, gl = lane - rg * LPR;
  const bool lane_on = rg < RPI;        // lanes beyond RPI * LPR idle (they shadow row group 0 with weight 0)
  const int rgc = lane_on ? rg : 0;
  const RowTree tree = make_row_tree(LPR, gl, lane);
  const uint32_t vstride = (uint32_t)LPR * 16u;          // bytes between a lane's consecutive float4 slots
  const uint32_t lane_off = (uint32_t)(rgc * KP + gl * 4) * 4u;  // byte offset of the lane's first slot in an iteration
  const uint32_t it_stride = (uint32_t)(RPI * KP) * 4u;  // bytes between warp iterations of a stage
  auto gsync = [&]() { umma::named_sync(bar_id, GT); };
  // Stages of a step are handed out dynamically (the rows per CTA and step rarely divide evenly among the
  // warps).  The counter of step s is re-armed two steps earlier, when every warp of the group has left step s - 4.
  // Only min(warps, ring depth) warps take stages: with more, a warp could wait for the second refill of a slot
  // whose first refill is still outstanding, and an mbarrier phase parity cannot tell those two apart.
  int* tick = sh_tick[isU ? 0 : 1];
  const bool takes_stages = gw < (isU ? DU : DL);
  auto next_stage = [&](int s) {
    int i = 0x7fffffff;
    if (takes_stages) {
      if (lane == 0) i = atomicAdd(tick + (s & 3), 1);
      i = __shfl_sync(0xffffffffu, i, 0);
    }
    return i;
  };
  // signal(c, slot, half): this CTA's group is done with the step counted by c; with sharded cells the CTA that
  // completes the count publishes the finished half (slot, half) of the local accumulators to every rank
  auto signal = [&](unsigned* c, int slot, int half) {
    gsync();
    if (!multi) {
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
      sh_last[isU ? 0 : 1] = (old == (unsigned)grid - 1u) ? 1 : 0;
    }
    gsync();
    if (sh_last[isU ? 0 : 1]) {
      const float* src = a.acc + (size_t)slot * SL + (size_t)half * x.XH;  // complete: every CTA's atomics preceded its count
      const size_t entry = ((size_t)(slot * 2 + half) * x.world + x.rank);
      for (int r = 0; r < x.world; ++r) {
        float* dst = x.data[r] + entry * x.XH;
        for (int i = gt * 4; i < x.XH; i += GT * 4)
          *reinterpret_cast<float4*>(dst + i) = __ldcg(reinterpret_cast<const float4*>(src + i));
      }
      __threadfence_system();
      gsync();
      if (gt < x.world) st_release_sys_u32(x.flag[gt] + entry, x.epoch);
    }
  };
  auto wait_for = [&](const unsigned* c) {
    if (gt == 0) {
      while (ld_acquire_u32(c) < (unsigned)grid) __nanosleep(20);
      __threadfence();
    }
    gsync();
  };
  // sharded cells: all ranks' entries of (slot, half) have arrived in the local exchange area
  auto wait_half = [&](int slot, int half) {
    if (gt < x.world) {
      const unsigned* f = x.local_flag + (size_t)(slot * 2 + half) * x.world + gt;
      while (ld_acquire_sys_u32(f) != x.epoch) __nanosleep(40);
      __threadfence();
    }
    gsync();
  };
  // O_s, E_s (and P_s) of element (b, k) from the local or the exchanged accumulators
  auto derive_any = [&](const TableView& tv, int s, int b, int k, float& o, float& e, float& pp) {
    if (multi)
      derive_x(x, tv, a.Pr_b, a.theta, s, b, k, true, o, e, pp);
    else
      derive(tv, a.Pr_b, a.theta, b, k, o, e, pp);
  };
  auto stamp = [&](int s, int slot_id) {
    if (a.dbg && cta == a.dbg_cta && gt == 0) {
      long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      a.dbg[((size_t)(s - a.s_begin + 1) * 2 + (isU ? 0 : 1)) * 8 + slot_id] = tns;
    }
  };

  float cs[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) cs[v][c] = 0.f;

  // column sums of the step: registers -> per-warp rows of `part` -> global accumulators of tuple q
  auto flush = [&](float* part, int q, float* dst_O, float* dst_rs) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float col = lane_on ? cs[v][c] : 0.f;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
          if (m < RPI) {  // warp-uniform: fold row group rg + m onto rg
            const float y = __shfl_down_sync(0xffffffffu, col, (unsigned)(m * LPR) & 31u);
            if (rg + m < RPI) col += y;
          }
        }
        cs[v][c] = 0.f;
        if (lane < LPR) part[(size_t)gw * KP + (gl + LPR * v) * 4 + c] = col;
      }
    gsync();
    for (int k = gt; k < K; k += GT) {
      float tsum = 0.f;
      for (int w = 0; w < NW; ++w) tsum += part[(size_t)w * KP + k];
      atomicAdd(dst_rs + k, tsum);
      for (int c = 0; c < C; ++c) atomicAdd(dst_O + (size_t)__ldg(a.tuple_levels + q * C + c) * KS + k, tsum);
    }
  };

  if (isU) {
    // ------------------------------- update group: the critical path -------------------------------
    float okd = 0.f, oent = 0.f;
    const bool sig_u = a.sigma_uniform != 0;
    auto flush_objective = [&](int t) {
      okd = warp_sum(okd);
      oent = warp_sum(oent);
      if (lane == 0) {
        atomicAdd(&sh_obj[0], (double)okd);
        atomicAdd(&sh_obj[1], (double)oent);
      }
      gsync();
      if (gt == 0) {
        atomicAdd(a.obj + 2 * t + 0, sh_obj[0]);
        atomicAdd(a.obj + 2 * t + 1, sh_obj[1]);
        sh_obj[0] = 0.0;
        sh_obj[1] = 0.0;
      }
      gsync();
      okd = 0.f;
      oent = 0.f;
    };
    auto owners = [&](int s) {  // materialise O_s, E_s (ring), the end-of-round tables and save P_s
      const int t = s / nb, j = s - t * nb;
      TableView tv = tables_for(s);
      float* outO = a.ring + (size_t)(s & 1) * 2 * BK;
      float* outE = outO + BK;
      float* Ps = a.Psave + ((size_t)(t & 1) * nb + j) * BK;
      for (int idx = cta * GT + gt; idx < BK; idx += grid * GT) {
        const int b = idx / KS, k = idx - b * KS;
        float o = 0.f, e = 0.f, pp = 0.f;
        if (k < K) derive_any(tv, s, b, k, o, e, pp);
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

    int gbase = 0;  // stages consumed by the group before this step (same count as the producer's)
    for (int s = a.s_begin; s < a.s_end; ++s) {
      stamp(s, 0);
      if (gt == 0) tick[(s + 2) & 3] = 0;
      int lo, hi, q;
      range_of(s, lo, hi, q);
      const int nrows = hi - lo;
      const int nst = (nrows + SR - 1) / SR;
      const int t = s / nb;
      const bool writeR = (a.write_R_mask >> t) & 1u;
      if (a.use_barrier) {
        if (s > a.s_begin) wait_for(cntU + s - 1);            // add_{s-1}, ring, Psave of step s-1
        if (s > a.s_begin || a.prologue) wait_for(cntL + s);  // rem_s
        if (multi) {  // the other ranks' shares of add_{s-1}, rem_{s-1} and rem_s
          if (s >= 1) {
            wait_half(s + 1, 0);
            wait_half(s, 1);
          }
          wait_half(s + 1, 1);
        }
      }
      stamp(s, 1);
      if (nst > 0) {
        // penalty row of this CTA's tuple: Psum_k = sum_c P_s[level_c][k] and its logarithm
        {
          TableView tv = tables_for(s);
          for (int k = gt; k < KP; k += GT) {
            float v = 0.f;
            if (k < K)
              for (int c = 0; c < C; ++c) {
                float o, e, pp;
                derive_any(tv, s, __ldg(a.tuple_levels + q * C + c), k, o, e, pp);
                v += pp;
              }
            tabU[k] = v;
            tabU[KP + k] = (k < K) ? fast_log(v) : 0.f;
          }
        }
        gsync();
        stamp(s, 2);
        float4 pP[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          pP[v] = *reinterpret_cast<const float4*>(tabU + (gl + LPR * v) * 4);
          if (!lane_on) pP[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const uint32_t tabU_l = umma::smem_u32(tabU) + (uint32_t)gl * 16u;          // Psum slots of this lane
        const uint32_t tabLg_l = tabU_l + (uint32_t)KP * 4u;                         // log Psum slots
        const uint32_t sig_l = umma::smem_u32(sig) + (uint32_t)gl * 16u;
        // one warp iteration: RPI rows, LPR lanes each.  SIGU: all sigma_k equal (objective terms simplify)
        auto iteration = [&](uint32_t row_addr, bool ok, auto sigu_c) {
          constexpr bool SIGU = decltype(sigu_c)::value;
          float4 u4[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) u4[v] = lds4(row_addr + v * vstride);
          float ep[NV][4];
          float ssum = 0.f, Aacc = 0.f, Bacc = 0.f, Sacc = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const float4 l4 = lds4(tabLg_l + v * vstride);
            const float uu[4] = {u4[v].x, u4[v].y, u4[v].z, u4[v].w};
            const float pp[4] = {pP[v].x, pP[v].y, pP[v].z, pP[v].w};
            const float ll[4] = {l4.x, l4.y, l4.z, l4.w};
            float sg[4] = {1.f, 1.f, 1.f, 1.f};
            if constexpr (!SIGU) {
              const float4 s4 = lds4(sig_l + v * vstride);
              sg[0] = s4.x, sg[1] = s4.y, sg[2] = s4.z, sg[3] = s4.w;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float xr = fast_exp(uu[c]) * pp[c];  // un-normalised R (>= 0)
              ep[v][c] = xr;
              ssum += xr;
              if constexpr (SIGU) {
                Aacc = fmaf(xr, uu[c], Aacc);
                Bacc = fmaf(xr, ll[c], Bacc);
              } else {
                const float tt = sg[c] * xr;
                Aacc = fmaf(tt, uu[c], Aacc);
                Bacc = fmaf(tt, ll[c], Bacc);
                Sacc += tt;
              }
            }
          }
          ssum = u3_row_total(ssum, tree);
          const float sdiv = (ssum == 0.f) ? 1.f : ssum;  // arma::normalise(.., 1, 0): zero norm divides by 1
          const float inv = ok ? fast_rcp(sdiv) : 0.f;     // rows beyond the stage / idle lanes weigh 0
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) cs[v][c] = fmaf(ep[v][c], inv, cs[v][c]);
          if (writeR && ok) {
#pragma unroll
            for (int v = 0; v < NV; ++v)
              sts4(row_addr + v * vstride, make_float4(ep[v][0] * inv, ep[v][1] * inv, ep[v][2] * inv, ep[v][3] * inv));
          }
          // sum_k R dist = -sum sigma R U ;  sum_k sigma R log R = sum sigma R (U + log Psum - log s)
          const float ls = fast_log(sdiv);
          if constexpr (SIGU) {
            const float srow = (gl == 0) ? ssum : 0.f;  // the row total once per row
            const float w = a.sigma0 * inv;
            okd = fmaf(-w, Aacc, okd);
            oent = fmaf(w, (Aacc + Bacc) - ls * srow, oent);
          } else {
            okd = fmaf(-inv, Aacc, okd);
            oent = fmaf(inv, (Aacc + Bacc) - ls * Sacc, oent);
          }
        };
        for (int i = next_stage(s); i < nst; i = next_stage(s)) {
          const int g = gbase + i;
          const int slot = g % DU, use = g / DU;
          umma::mbar_wait(fullU + slot, use & 1);
          const int nr = (nrows - i * SR < SR) ? nrows - i * SR : SR;
          float* sbase = ringU + (size_t)slot * SR * KP;
          uint32_t row_addr = umma::smem_u32(sbase) + lane_off;
          int r = rgc;
          if (sig_u) {
            for (int r0 = 0; r0 < nr; r0 += RPI, r += RPI, row_addr += it_stride)
              iteration(row_addr, lane_on && r < nr, std::true_type());
          } else {
            for (int r0 = 0; r0 < nr; r0 += RPI, r += RPI, row_addr += it_stride)
              iteration(row_addr, lane_on && r < nr, std::false_type());
          }
          if (writeR) {
            umma::fence_proxy_async();  // the rows were rewritten through the generic proxy
            __syncwarp();
            if (lane < nr) umma::bulk_store(a.R + (size_t)metaU[slot * 32 + lane] * KS, sbase + (size_t)lane * KP, row_bytes);
            umma::bulk_commit();
            // A warp never holds a stage while it waits for another one (the producer refills in order, so a
            // held slot could be the very slot the next stage needs): wait until the stores have read the rows.
            umma::bulk_wait_read();
          }
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(emptyU + slot);
        }
        stamp(s, 4);
        float* nslot = a.acc + (size_t)(s + 2) * SL;  // slot(s+1): add_s goes to its addprev part
        flush(partU, q, nslot, nslot + BK);
        stamp(s, 5);
      }
      owners(s);
      if ((s + 1) % nb == 0) flush_objective(s / nb);
      stamp(s, 6);
      if (a.use_barrier) signal(cntU + s, s + 2, 0);  // add_s is half 0 of slot(s + 1)
      gbase += nst;
    }
    if (a.s_end % nb != 0 && a.s_end > a.s_begin) flush_objective((a.s_end - 1) / nb);  // partial round
    umma::bulk_wait_all();
  } else {
    // ------------------------------- look-ahead group -------------------------------
    // rem_s: column sums of the CURRENT R of block s, recomputed from U and the penalty rows the cells were
    // last updated with (previous round; round 0: the assignment's plain softmax, or R itself if the user set it)
    int gbase = 0;
    for (int s = look_first; s <= look_last; ++s) {
      int lo, hi, q;
      range_of(s, lo, hi, q);
      const int nrows = hi - lo;
      const int nst = (nrows + SR - 1) / SR;
      const int t = s / nb;
      const bool fromR = (s < nb) && a.first_round_from_R;
      if (gt == 0) tick[(s + 2) & 3] = 0;
      stamp(s - 1, 1);
      if (a.use_barrier) {
        int need = (t > 0) ? t * nb - 1 : -1;  // the previous round's tables must all be saved
        const int bound = s - UPD_RUNAHEAD;    // L2 footprint: stay <= UPD_RUNAHEAD steps ahead of the update group
        if (bound > need) need = bound;
        if (need >= a.s_begin) wait_for(cntU + need);
      }
      stamp(s - 1, 2);
      if (nst > 0) {
        if (!fromR) {
          const int sp_cur = s - 1;  // the step the update group is running (per-step launches: P not saved yet)
          const int tp = t - 1;
          const int njp = (t == 0) ? 1 : nb;
          for (int idx = gt; idx < njp * KP; idx += GT) {
            const int jp = idx / KP, k = idx - jp * KP;
            float v = 0.f;
            if (k < K) {
              if (t == 0) {
                v = (float)C;  // R of the assignment step: plain softmax(U)
              } else {
                const int sp = tp * nb + jp;
                if (sp == sp_cur && !a.use_barrier) {
                  TableView tv = tables_for(sp);
                  for (int c = 0; c < C; ++c) {
                    float o, e, pp;
                    derive(tv, a.Pr_b, a.theta, __ldg(a.tuple_levels + q * C + c), k, o, e, pp);
                    v += pp;
                  }
                } else {
                  const float* Ps = a.Psave + ((size_t)(tp & 1) * nb + jp) * BK;
                  for (int c = 0; c < C; ++c) v += __ldcg(Ps + (size_t)__ldg(a.tuple_levels + q * C + c) * KS + k);
                }
              }
            }
            tabL[idx] = v;
          }
        }
        gsync();
        stamp(s - 1, 3);
        const uint32_t tabL_l = umma::smem_u32(tabL) + (uint32_t)gl * 16u;
        const uint32_t tab_row = (uint32_t)KP * 4u;
        for (int i = next_stage(s); i < nst; i = next_stage(s)) {
          const int g = gbase + i;
          const int slot = g % DL, use = g / DL;
          umma::mbar_wait(fullL + slot, use & 1);
          const int nr = (nrows - i * SR < SR) ? nrows - i * SR : SR;
          uint32_t row_addr = umma::smem_u32(ringL + (size_t)slot * SR * KP) + lane_off;
          const int* mrow = metaL + slot * 32;
          int r = rgc;
          if (fromR) {
            for (int r0 = 0; r0 < nr; r0 += RPI, r += RPI, row_addr += it_stride) {
              const float w = (lane_on && r < nr) ? 1.f : 0.f;
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                const float4 u4 = lds4(row_addr + v * vstride);
                cs[v][0] = fmaf(u4.x, w, cs[v][0]);
                cs[v][1] = fmaf(u4.y, w, cs[v][1]);
                cs[v][2] = fmaf(u4.z, w, cs[v][2]);
                cs[v][3] = fmaf(u4.w, w, cs[v][3]);
              }
            }
          } else {
            for (int r0 = 0; r0 < nr; r0 += RPI, r += RPI, row_addr += it_stride) {
              const bool ok = lane_on && r < nr;
              const int prv = ok ? mrow[r] : nb;  // row nb of tabL is all zero
              const uint32_t pw = tabL_l + (uint32_t)prv * tab_row;
              float4 u4[NV];
#pragma unroll
              for (int v = 0; v < NV; ++v) u4[v] = lds4(row_addr + v * vstride);
              float ep[NV][4];
              float ssum = 0.f;
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                const float4 p4 = lds4(pw + v * vstride);
                ep[v][0] = fast_exp(u4[v].x) * p4.x;
                ep[v][1] = fast_exp(u4[v].y) * p4.y;
                ep[v][2] = fast_exp(u4[v].z) * p4.z;
                ep[v][3] = fast_exp(u4[v].w) * p4.w;
                ssum += (ep[v][0] + ep[v][1]) + (ep[v][2] + ep[v][3]);
              }
              ssum = u3_row_total(ssum, tree);
              const float sdiv = (ssum == 0.f) ? 1.f : ssum;
              const float inv = ok ? fast_rcp(sdiv) : 0.f;
#pragma unroll
              for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int c = 0; c < 4; ++c) cs[v][c] = fmaf(ep[v][c], inv, cs[v][c]);
            }
          }
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(emptyL + slot);
        }
        stamp(s - 1, 4);
        float* slot_s = a.acc + (size_t)(s + 1) * SL;  // slot(s): rem_s
        flush(partL, q, slot_s + BK + KS, slot_s + 2 * BK + KS);
        stamp(s - 1, 5);
      }
      if (a.use_barrier) signal(cntL + s, s + 1, 1);  // rem_s is half 1 of slot(s)
      gbase += nst;
    }
  }
}

// k_update_finalize for sharded cells with the peer-memory exchange: O = O_S, E = E_S from the exchanged
// partial sums (waits for the last halves of the other ranks, whose kernels may still be running).
__global__ void k_update_finalize3(UpdArgs a, Upd3Xch x, int S, float* __restrict__ O, float* __restrict__ E) {
  const int KS = a.KS, BK = a.B * KS;
  if (S >= 1 && (int)threadIdx.x < x.world) {
    const unsigned* f0 = x.local_flag + (size_t)((S + 1) * 2 + 0) * x.world + threadIdx.x;  // add_{S-1}
    const unsigned* f1 = x.local_flag + (size_t)(S * 2 + 1) * x.world + threadIdx.x;        // rem_{S-1}
    while (ld_acquire_sys_u32(f0) != x.epoch) __nanosleep(40);
    while (ld_acquire_sys_u32(f1) != x.epoch) __nanosleep(40);
    __threadfence();
  }
  __syncthreads();
  TableView tv;
  tv.ringO = a.ring + (size_t)((S - 1) & 1) * 2 * BK;
  tv.ringE = tv.ringO + BK;
  tv.prev = tv.cur = nullptr;
  tv.BK = BK;
  tv.KS = KS;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < BK; idx += gridDim.x * blockDim.x) {
    const int b = idx / KS, k = idx - b * KS;
    float o = 0.f, e = 0.f, pp;
    if (k < a.K) derive_x(x, tv, a.Pr_b, a.theta, S, b, k, false, o, e, pp);
    O[idx] = o;
    E[idx] = e;
  }
}

}  // namespace hb
