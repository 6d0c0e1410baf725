// scatter_tma.cuh - k_vmat_tma: the scatter contraction V_s = phi_s^T G + G^T phi_s with TMA-staged operands.
//
// Same mathematics and work decomposition as k_vmat (scatter_kernel.cuh; ScalarOperatorToMatrixAdder.cpp:225-303): persistent
// CTAs pull work items from a device queue; a block's upper triangle of 32 x 32 warp tiles U[I,J] = [phi_I | G_I] . [G_J | phi_J]^T
// is processed in host-scheduled rounds of <= 8 warp tiles on <= 6 staged 32-row groups; FP64 DMMA m8n8k4, accumulators leave
// through red.global.add.f64.  What is new:
//   * the producer is ONE elected thread: per K chunk it arms the stage's "full" mbarrier with the byte count
//     (mbarrier.arrive.expect_tx) and issues one cp.async.bulk.tensor.2d per staged box (32 rows x TKP points of the phi slot
//     and of the G slot of every staged group) from a 2-D tensor map over the whole tile workspace ([rows] x [128 points]);
//     the TMA unit writes the boxes with the hardware swizzle (64 B rows: SWIZZLE_64B, 128 B rows: SWIZZLE_128B), which is what
//     keeps the DMMA fragment loads bank-conflict free without padding - no LDGSTS, no address arithmetic in any warp;
//   * a CTA is 8 DMMA warps + 1 producer warp (288 threads), so ptxas gets 112 registers per thread;
//   * the block's compact->function map sits in shared memory and the next round's descriptor is fetched while the current
//     round multiplies: no dependent global load at a round boundary;
//   * K splits inside a round are per k-step of a chunk (kmask has TKP / 4 bits): idle warps of a partial round take k-steps of
//     the heaviest tiles (scatter schedule v2, sxc_api.cu).
#pragma once

#include "sxc_common.cuh"

namespace sxc {

// one round of the v2 schedule (64 bytes, read as 16 words by the lanes of a warp)
struct __align__(16) ScatterRound2 {
  unsigned char ngroups;   // staged groups (<= MAXG)
  unsigned char pad[7];
  unsigned char group[8];  // 32-row group index inside the block of every staged slot
  unsigned char ta[8];     // per warp: staged slot of the row group I (0xff: idle warp)
  unsigned char tb[8];     // per warp: staged slot of the column group J >= I
  unsigned char kmask[8];  // per warp: k-steps of a chunk it multiplies (bit ks)
  unsigned char ga[8];     // per warp: block group index of I
  unsigned char gb[8];     // per warp: block group index of J
  unsigned char pad2[8];
};
static_assert(sizeof(ScatterRound2) == 64, "ScatterRound2 layout");

namespace scat2 {
constexpr int WARPS = 8;
constexpr int THREADS = (WARPS + 1) * 32;
constexpr int MAXG = 6;
template <int TKP_, int STAGES_ = (TKP_ == 8 ? 4 : 2)>
struct Cfg {
  static constexpr int TKP = TKP_;
  static constexpr int KS = TKP / 4;                  // k-steps per chunk
  static constexpr int NKC = BP / TKP;                // chunks per round
  static constexpr int STAGES = STAGES_;
  static constexpr int BOX_ELEMS = 32 * TKP;          // 32 rows x TKP points
  static constexpr int GROUP_ELEMS = 2 * BOX_ELEMS;   // phi box, then G box
  static constexpr int STAGE_ELEMS = MAXG * GROUP_ELEMS;
  static constexpr unsigned BOX_BYTES = BOX_ELEMS * sizeof(double);
};
template <int TKP>
constexpr size_t smem_bytes(int sig_cap) {
  return (size_t)Cfg<TKP>::STAGES * Cfg<TKP>::STAGE_ELEMS * sizeof(double) + 2 * Cfg<TKP>::STAGES * sizeof(uint64_t) +
         (size_t)sig_cap * sizeof(int) + 16 + 1024;
}
}  // namespace scat2

// offset (doubles) inside a staged box row of the K index that lane column lc uses in k-step ks, for a lane whose fragment row is lr
//   TKP = 8,  SWIZZLE_64B : 16-byte chunk c of row r sits at c ^ ((r >> 1) & 3); k-step ks uses the points {0,1,4,5} + 2 ks
//   TKP = 16, SWIZZLE_128B: chunk c of row r sits at c ^ (r & 7);               k-step ks uses the points {0,1,8,9} + 2 ks
// (rows of a fragment are m * 8 + lr: m does not change the swizzle term).  A 64-bit shared load is served per half-warp (4
// fragment rows x 4 columns): with these point sets its 16 lanes hit 16 different 8-byte bank pairs - conflict free.
template <int TKP>
__device__ __forceinline__ int frag_col(int ks, int lr, int lc) {
  if (TKP == 8) return ((((lc >> 1) * 2 + ks) ^ ((lr >> 1) & 3)) << 1) | (lc & 1);
  return ((((lc >> 1) * 4 + ks) ^ lr) << 1) | (lc & 1);
}

// K loop of one round for a warp tile with MFR x NFR valid 8 x 8 fragments; DIAG: only fragments m <= nn are multiplied
template <class C, int MFR, int NFR, bool DIAG>
__device__ __forceinline__ void vmat2_round(double (&acc)[4][4][2], const double* __restrict__ stage_base, uint64_t* full,
                                            uint64_t* empty, int& stage, int& pass, int slot_a, int slot_b, int kmask,
                                            bool active, int lane) {
  constexpr int TKP = C::TKP;
  const int lr = lane >> 2, lc = lane & 3;
  int col[C::KS];
#pragma unroll
  for (int ks = 0; ks < C::KS; ++ks) col[ks] = lr * TKP + frag_col<TKP>(ks, lr, lc);
  for (int kc = 0; kc < C::NKC; ++kc) {
    mbar_wait(full + stage, pass & 1);
    if (active) {
      const double* sI = stage_base + stage * C::STAGE_ELEMS + slot_a * C::GROUP_ELEMS;  // phi_I box, then G_I box
      const double* sJ = stage_base + stage * C::STAGE_ELEMS + slot_b * C::GROUP_ELEMS;
#pragma unroll
      for (int ks = 0; ks < C::KS; ++ks) {
        if (!((kmask >> ks) & 1)) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // phi_I . G_J^T, then G_I . phi_J^T
          double a[MFR], bq[NFR];
#pragma unroll
          for (int m = 0; m < MFR; ++m) a[m] = sI[col[ks] + m * 8 * TKP + (half ? C::BOX_ELEMS : 0)];
#pragma unroll
          for (int nn = 0; nn < NFR; ++nn) bq[nn] = sJ[col[ks] + nn * 8 * TKP + (half ? 0 : C::BOX_ELEMS)];
#pragma unroll
          for (int m = 0; m < MFR; ++m)
#pragma unroll
            for (int nn = DIAG ? m : 0; nn < NFR; ++nn) dmma884(acc[m][nn][0], acc[m][nn][1], a[m], bq[nn]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    if (++stage == C::STAGES) {
      stage = 0;
      ++pass;
    }
  }
}

// producer side of one work item (one elected thread): stage the phi and G boxes of every round's groups, chunk by chunk
template <class C>
__device__ __forceinline__ void vmat2_produce_item(const CUtensorMap* tmap, const ScatterRound2* __restrict__ rounds, int nr,
                                                   int row0, int rowG, double* stage_base, uint64_t* full, uint64_t* empty,
                                                   int& stage, int& pass) {
  for (int r = 0; r < nr; ++r) {
    const uint2 grp = __ldg(reinterpret_cast<const uint2*>(rounds[r].group));
    const int ng = rounds[r].ngroups;
    for (int kc = 0; kc < C::NKC; ++kc) {
      if (pass > 0) mbar_wait(empty + stage, (pass - 1) & 1);
      double* st = stage_base + stage * C::STAGE_ELEMS;
      mbar_arrive_expect_tx(full + stage, (unsigned)ng * 2u * C::BOX_BYTES);
      for (int i = 0; i < ng; ++i) {
        const int gi = (int)(((i < 4 ? grp.x : grp.y) >> (8 * (i & 3))) & 0xffu);
        tma_load_2d(st + i * C::GROUP_ELEMS, tmap, kc * C::TKP, row0 + gi * 32, full + stage);
        tma_load_2d(st + i * C::GROUP_ELEMS + C::BOX_ELEMS, tmap, kc * C::TKP, rowG + gi * 32, full + stage);
      }
      if (++stage == C::STAGES) {
        stage = 0;
        ++pass;
      }
    }
  }
}

// DMMA side of one work item: the rounds of a block (or of a segment of them), accumulators into the upper triangle of W
template <class C>
__device__ __forceinline__ void vmat2_consume_item(const ScatterRound2* __restrict__ rounds, int nr, int s, int sp, int nbf,
                                                   const int* s_sig, const double* stage_base, uint64_t* full, uint64_t* empty,
                                                   int& stage, int& pass, int warp, int lane, double* __restrict__ W) {
  const int lr = lane >> 2, lc = lane & 3;
  const int s8 = (s + 7) & ~7;  // rows beyond s rounded up to 8 are zero padding: their 8 x 8 fragments are skipped
  double acc[4][4][2];
  // descriptor words of round r: lane l holds word (l & 15)
  unsigned dw = __ldg(reinterpret_cast<const unsigned*>(rounds) + (lane & 15));
  for (int r = 0; r < nr; ++r) {
    const unsigned w_ta = __shfl_sync(0xffffffffu, dw, 4 + (warp >> 2)), w_tb = __shfl_sync(0xffffffffu, dw, 6 + (warp >> 2));
    const unsigned w_km = __shfl_sync(0xffffffffu, dw, 8 + (warp >> 2)), w_ga = __shfl_sync(0xffffffffu, dw, 10 + (warp >> 2));
    const unsigned w_gb = __shfl_sync(0xffffffffu, dw, 12 + (warp >> 2));
    const int sh = 8 * (warp & 3);
    const int slot_a = (w_ta >> sh) & 0xff, slot_b = (w_tb >> sh) & 0xff, kmask = (w_km >> sh) & 0xff;
    const int gI = (w_ga >> sh) & 0xff, gJ = (w_gb >> sh) & 0xff;
    if (r + 1 < nr) dw = __ldg(reinterpret_cast<const unsigned*>(rounds + r + 1) + (lane & 15));  // lands during the K loop
    const bool active = slot_a != 0xff;
    // valid 8-row fragments of the row group I and the column group J (only the last group of a block is short)
    const int mfr = active ? min(4, (s8 - gI * 32) >> 3) : 0;
    const int nfr = active ? min(4, (s8 - gJ * 32) >> 3) : 0;
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) acc[m][nn][0] = acc[m][nn][1] = 0.0;
    if (active && slot_a == slot_b && mfr == 4) {  // full diagonal tile
      vmat2_round<C, 4, 4, true>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane);
    } else if (mfr == 4 || !active) {
      switch (active ? nfr : 4) {
        case 1: vmat2_round<C, 4, 1, false>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
        case 2: vmat2_round<C, 4, 2, false>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
        case 3: vmat2_round<C, 4, 3, false>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
        default: vmat2_round<C, 4, 4, false>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
      }
    } else {  // a short row group is the last one, so the tile is the last diagonal tile: nfr == mfr
      switch (mfr) {
        case 1: vmat2_round<C, 1, 1, true>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
        case 2: vmat2_round<C, 2, 2, true>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
        default: vmat2_round<C, 3, 3, true>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
      }
    }
    if (active) {
      // V += Proj V_s Proj^T (:301), upper triangle (compact i <= j <=> global sig[i] <= sig[j])
      const int i0 = gI * 32, j0 = gJ * 32;
      int rowi[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) rowi[m] = s_sig[min(i0 + m * 8 + lr, sp - 1)];
#pragma unroll
      for (int nn = 0; nn < 4; ++nn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = j0 + nn * 8 + 2 * lc + e;
          if (j >= s) continue;
          const size_t colo = (size_t)s_sig[j] * nbf;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int i = i0 + m * 8 + lr;
            if (i <= j) red_add_f64(W + colo + rowi[m], acc[m][nn][e]);
          }
        }
    }
  }
}

template <int TKP, int NSTAGES = (TKP == 8 ? 4 : 2)>
__global__ void __launch_bounds__(scat2::THREADS, 2)
k_vmat_tma(const __grid_constant__ CUtensorMap tmap, PlanView plan, int nbf, const WorkItem* __restrict__ items, int nitems,
           int* __restrict__ counter, const int* __restrict__ skip_flag, const ScatterRound2* __restrict__ tpl,
           const int* __restrict__ tpl_off, int sig_cap, double* __restrict__ W) {
  using namespace scat2;
  using C = Cfg<TKP, NSTAGES>;
  extern __shared__ unsigned char smem_raw[];
  // the hardware swizzle works on absolute shared-memory address bits: the boxes must start on a multiple of their pattern
  // (512 B / 1024 B), so the ring starts on the next 1 KB boundary (the launch reserves the slack)
  double* stage_base = reinterpret_cast<double*>(
      smem_raw + ((1024u - (static_cast<unsigned>(__cvta_generic_to_shared(smem_raw)) & 1023u)) & 1023u));
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + C::STAGES * C::STAGE_ELEMS);
  uint64_t* empty = full + C::STAGES;
  int* s_sig = reinterpret_cast<int*>(empty + C::STAGES);
  int* s_next = s_sig + sig_cap;  // sig_cap >= the largest s_pad of the plan

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(full + i, 1);       // the producer's arrive.expect_tx; the TMA unit completes the bytes
      mbar_init(empty + i, WARPS);  // one arrive per DMMA warp
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmap);
  }
  int stage = 0, pass = 0;  // ring position: both sides walk the stages in the same order over the whole life of the CTA

  for (;;) {
    __syncthreads();  // (orders the barrier initialisation before first use; s_sig / s_next of the previous item are free)
    if (tid == 0) *s_next = atomicAdd(counter, 1);
    __syncthreads();
    const int qi = *s_next;
    if (qi >= nitems) break;
    const WorkItem item = items[qi];
    const int q = item.q;
    if (skip_flag[q]) continue;
    const int s = plan.s[q];
    const int sp = plan.s_pad[q];
    const int n32 = sp >> 5;
    const ScatterRound2* __restrict__ rounds = tpl + tpl_off[n32] + item.begin;  // this item's segment of the rounds
    const int nr = item.end - item.begin;

    if (warp == WARPS) {
      // ---------------- producer: one thread drives the TMA unit (the other lanes of the warp take no part; lane 0 keeps the
      // ring position in its registers from item to item)
      if (lane == 0) {
        const int row0 = (int)(plan.phi_off[q] / BP);  // first row of the block's tile in the workspace
        vmat2_produce_item<C>(&tmap, rounds, nr, row0, row0 + 4 * sp, stage_base, full, empty, stage, pass);
      }
    } else {
      // ---------------- DMMA warps
      {  // compact -> function map of the block into shared memory (read by the RED epilogues of every round)
        const int* __restrict__ sig_g = plan.sig_bf + (size_t)q * plan.nbf_pad;
        for (int c = tid; c < sp; c += WARPS * 32) s_sig[c] = sig_g[c];
        __syncwarp();
        asm volatile("barrier.sync 1, %0;\n" ::"n"(WARPS * 32) : "memory");
      }
      vmat2_consume_item<C>(rounds, nr, s, sp, nbf, s_sig, stage_base, full, empty, stage, pass, warp, lane, W);
    }
  }
}

}  // namespace sxc
