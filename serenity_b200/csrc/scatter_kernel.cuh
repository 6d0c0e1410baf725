// scatter_kernel.cuh - v_xc on the grid -> Fock-matrix contribution: V_s = phi_s^T G + G^T phi_s on DMMA tiles.
//
// Row 8a-5 of SURVEY.md: ScalarOperatorToMatrixAdder::addBlock (src/data/grid/ScalarOperatorToMatrixAdder.cpp:225-303):
//     a = w o v_rho,  b_c = w o g_c;   skip the block if (sum|a| + sum|b_x| + sum|b_y| + sum|b_z|)/n < blockAveThreshold;
//     G = diag(b_x) d_x phi_s + diag(b_y) d_y phi_s + diag(b_z) d_z phi_s + 1/2 diag(a) phi_s;
//     T = phi_s^T G;  V_s = T + T^T;  V += Proj V_s Proj^T.
// The LDA variant (:179-223, V_s = phi_s^T diag(a) phi_s) is the same expression with b = 0.
// B200 design: k_form_g builds G in the fifth slot of the tile (bandwidth-bound: reads 4 tile components, writes 1;
// fused with the block-average test; phi and grad phi stay intact for the second spin of an UNRESTRICTED build).  k_vmat is a persistent kernel (8 DMMA warps + 2 producer warps, two CTAs per SM) that pulls blocks from a device
// work queue (largest first) and computes only the upper triangle of V_s, cut into 32 x 32 warp tiles
//          U[I,J] = [phi_I | G_I] . [G_J | phi_J]^T      (stacked K = 2 x 128 points, DMMA m8n8k4)
// in host-scheduled "rounds" (sxc_api.cu: scatter_schedule): <= 8 warp tiles that touch <= 6 distinct 32-row groups.
// The producer warps stage phi and G rows of those groups through a 4-stage cp.async ring over 8-point K chunks (the
// ring runs across round and block boundaries, so only the first round of a CTA pays the fill latency); every DMMA warp owns one tile
// (rounds with <= 4 tiles split the two k-steps of a chunk over two warps), and the accumulators go straight into
// the GPU-resident upper triangle with FP64 red.global (RED.E.ADD.F64).  k_mirror copies the strict upper triangle
// down once per build.  Scheduling at warp-tile granularity keeps > 90 % of the DMMA slots busy for any s (128 x 128
// CTA tiles left half of them idle at s ~ 280).  (Forming G inside k_vmat was tried and lost: with 256 threads per CTA
// the streaming phase cannot keep enough loads in flight, 4.3 ms against 0.74 + 2.9 ms.)
#pragma once

#include "sxc_common.cuh"

namespace sxc {

// ------------------------------------------------------------------------------------------------------------
// K3b: per block: weights x potential, block-average test, G into the fifth tile slot.  256 threads; gridDim.y CTAs share the
// function rows of a block (a small shard would otherwise leave too few bytes in flight to fill HBM).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 8)  // 32 registers: 8 CTAs per SM keep the streaming loads in flight
k_form_g(GridView g, PlanView plan, const int* __restrict__ order, double block_ave_thr, double a_scale,
         const double* __restrict__ v_rho, const double* __restrict__ v_gx, const double* __restrict__ v_gy,
         const double* __restrict__ v_gz, int npot, size_t pot_stride, double* __restrict__ phi_buf,
         int* __restrict__ skip_flag) {
  __shared__ double sa[BP], sx[BP], sy[BP], sz[BP];
  __shared__ double scratch[32];
  const int q = order[blockIdx.x];
  const int blk = plan.block_id[q];
  const long first = (long)blk * g.blocksize;
  const int n = (int)min((long)g.blocksize, g.npts - first);
  const int tid = threadIdx.x;
  // npot > 1: the potentials of several operators are scattered in one pass (sxc_build_nadd_multi, summed matrix); each operator
  // keeps ITS OWN block-average test, exactly as if it had been scattered alone (ScalarOperatorToMatrixAdder.cpp:262-268), and
  // only the ones that pass enter a, b
  bool any_pass = false;
  if (tid < BP) sa[tid] = sx[tid] = sy[tid] = sz[tid] = 0.0;
  for (int k = 0; k < npot; ++k) {
    double a = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
    if (tid < n) {
      const size_t o = (size_t)k * pot_stride + first + tid;
      const double wp = g.w[first + tid];
      a = wp * v_rho[o];
      if (v_gx) {
        bx = wp * v_gx[o];
        by = wp * v_gy[o];
        bz = wp * v_gz[o];
      }
    }
    const double total = block_sum(fabs(a) + fabs(bx) + fabs(by) + fabs(bz), scratch);  // (its barriers order the smem updates)
    if (!(total / (double)n < block_ave_thr)) {  // :262-268
      any_pass = true;
      if (tid < BP) {
        sa[tid] += a;
        sx[tid] += bx;
        sy[tid] += by;
        sz[tid] += bz;
      }
    }
  }
  __syncthreads();
  const int s = plan.s[q];
  const bool skip = !any_pass || s == 0;
  if (tid == 0 && blockIdx.y == 0) skip_flag[q] = skip ? 1 : 0;
  if (skip) return;
  const int sp = plan.s_pad[q];
  const size_t comp_stride = (size_t)sp * BP;
  double* __restrict__ tile = phi_buf + plan.phi_off[q];
  const int p = tid & (BP - 1);
  const double a = a_scale * sa[p], bx = sx[p], by = sy[p], bz = sz[p];  // a_scale = 1/2: the T + T^T trick of :281
  for (int c = (tid >> 7) + 2 * blockIdx.y; c < sp; c += 2 * gridDim.y) {
    const size_t i = (size_t)c * BP + p;
    tile[4 * comp_stride + i] = bx * tile[comp_stride + i] + by * tile[2 * comp_stride + i] +
                                bz * tile[3 * comp_stride + i] + a * tile[i];  // :276-281
  }
}

namespace scat {
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int MAXG = 6;              // distinct 32-row groups staged per round
constexpr int TKP = 8;               // points per K chunk
constexpr int NKC = BP / TKP;        // 16 chunks per round
constexpr int STRIDE = TKP + 4;      // 12 doubles: conflict-free fragment loads
constexpr int GROUP_ELEMS = 64 * STRIDE;        // 32 phi rows then 32 G rows of one group
constexpr int STAGE_ELEMS = MAXG * GROUP_ELEMS;  // 4608 doubles
constexpr int STAGES = 3;
constexpr size_t smem_bytes() { return (size_t)STAGES * STAGE_ELEMS * sizeof(double); }
}  // namespace scat

// one round of the schedule of a block with s_pad / 32 row groups: which groups to stage, which tile each warp owns
struct __align__(8) ScatterRound {
  unsigned char ngroups;       // staged groups (<= MAXG)
  unsigned char pad[7];
  unsigned char group[8];      // 32-row group index inside the block, [0, s_pad / 32)  (8-byte aligned: one load)
  unsigned char ta[8];         // per warp: staged slot of the row group I (0xff: idle warp)
  unsigned char tb[8];         // per warp: staged slot of the column group J >= I
  unsigned char kmask[8];      // per warp: k-steps of a chunk it multiplies (bit 0 / bit 1)
};
static_assert(sizeof(ScatterRound) == 40, "ScatterRound layout");

// ------------------------------------------------------------------------------------------------------------
// K4: persistent scatter.  grid <= 2 x SMs; work items (blocks, or segments of a block's rounds when the shard is small)
// are taken from `items` through the atomic `counter`.  8 DMMA warps + 2 producer warps: warps 8 and 9 alone issue the
// cp.async copies of a stage (warp 8 the phi rows, warp 9 the G rows of the staged groups: 24 x 16 B per lane and chunk)
// and a "full" mbarrier tracks their completion (cp.async.mbarrier.arrive.noinc); the DMMA warps wait on "full",
// multiply and arrive on the stage's "empty" mbarrier, which the producers await before refilling.  No CTA-wide barrier
// and no producer code in the K loop of the DMMA warps (together 23 % of the warp time of the barrier-synchronised ring
// this replaced, profiles/r01_ncu_source_density.md: 2.93 -> 2.82 ms); warps may drift by up to a stage.  10 warps cap
// ptxas at 96 registers, so the two products of a k-step are loaded and issued one after the other.
// ------------------------------------------------------------------------------------------------------------
namespace scat {
constexpr int PWARPS = 2;
constexpr int PTHREADS = (WARPS + PWARPS) * 32;
// dense, XOR-swizzled staging (no padding columns): row r of a group holds its 8 points at r * 8 + (col ^ 4 * ((r >> 1) & 1)),
// which keeps the fragment loads (8 rows x 4 columns per request) conflict free and makes room for a fourth stage
constexpr int PSTRIDE = TKP;                        // 8 doubles
constexpr int PGROUP_ELEMS = 64 * PSTRIDE;          // 512
constexpr int PSTAGE_ELEMS = MAXG * PGROUP_ELEMS;   // 3072 doubles = 24 KB
constexpr int PSTAGES = 4;
constexpr size_t smem_bytes_pipe() { return (size_t)PSTAGES * PSTAGE_ELEMS * sizeof(double) + 2 * PSTAGES * sizeof(uint64_t); }
}  // namespace scat

// K loop of one round for a warp tile with MFR x NFR valid 8 x 8 fragments (4 x 4 except where the last, short row group of
// a block is involved: rows beyond s rounded up to 8 are zero padding).  Compile-time fragment counts keep the DMMA stream
// free of predicates (predicating the 4 x 4 loop nest cost 14 %).
// DIAG: the tile sits on the diagonal of V_s (I == J); only the upper triangle is kept, so the 8 x 8 fragments strictly below
// the diagonal (m > nn: 6 of 16) are not multiplied at all.
template <int MFR, int NFR, bool DIAG = false>
__device__ __forceinline__ void vmat_round(double (&acc)[4][4][2], const double* __restrict__ stage_base, uint64_t* full,
                                           uint64_t* empty, int& stage, int& pass, int slot_a, int slot_b, int kmask,
                                           bool active, int lane) {
  using namespace scat;
  const int lr = lane >> 2, lc = lane & 3;
  for (int kc = 0; kc < NKC; ++kc) {
    mbar_wait(full + stage, pass & 1);
    if (active) {
      const double* sI = stage_base + stage * PSTAGE_ELEMS + slot_a * PGROUP_ELEMS;  // phi_I rows 0..31, G_I rows 32..63
      const double* sJ = stage_base + stage * PSTAGE_ELEMS + slot_b * PGROUP_ELEMS;
#pragma unroll
      for (int ks = 0; ks < TKP / 4; ++ks) {
        if (!((kmask >> ks) & 1)) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // phi_I . G_J^T, then G_I . phi_J^T
          double a[MFR], bq[NFR];
          const int o = lr * PSTRIDE + ((ks * 4 + lc) ^ (4 * ((lr >> 1) & 1)));
#pragma unroll
          for (int m = 0; m < MFR; ++m) a[m] = sI[o + m * 8 * PSTRIDE + (half ? 32 * PSTRIDE : 0)];
#pragma unroll
          for (int nn = 0; nn < NFR; ++nn) bq[nn] = sJ[o + nn * 8 * PSTRIDE + (half ? 0 : 32 * PSTRIDE)];
#pragma unroll
          for (int m = 0; m < MFR; ++m)
#pragma unroll
            for (int nn = DIAG ? m : 0; nn < NFR; ++nn) dmma884(acc[m][nn][0], acc[m][nn][1], a[m], bq[nn]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    if (++stage == PSTAGES) {
      stage = 0;
      ++pass;
    }
  }
}

__global__ void __launch_bounds__(scat::PTHREADS, 2)
k_vmat(PlanView plan, int nbf, const WorkItem* __restrict__ items, int nitems, int* __restrict__ counter,
            const int* __restrict__ skip_flag, const ScatterRound* __restrict__ tpl, const int* __restrict__ tpl_off,
            const double* __restrict__ phi_buf, double* __restrict__ W) {
  using namespace scat;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + PSTAGES * PSTAGE_ELEMS);
  uint64_t* empty = full + PSTAGES;
  __shared__ int s_next;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  if (tid < PSTAGES) {
    mbar_init(full + tid, PWARPS * 32);
    mbar_init(empty + tid, WARPS);
  }
  // ring position: both sides walk the stages in the same order over the whole life of the CTA
  int stage = 0, pass = 0;  // pass = completed trips around the ring

  for (;;) {
    __syncthreads();  // (also orders the mbarrier initialisation before their first use)
    if (tid == 0) s_next = atomicAdd(counter, 1);
    __syncthreads();
    const int qi = s_next;
    if (qi >= nitems) break;
    const WorkItem item = items[qi];
    const int q = item.q;
    if (skip_flag[q]) continue;
    const int s = plan.s[q];
    const int sp = plan.s_pad[q];
    const size_t comp_stride = (size_t)sp * BP;
    const double* __restrict__ tile = phi_buf + plan.phi_off[q];
    const int n32 = sp >> 5;
    const ScatterRound* __restrict__ rounds = tpl + tpl_off[n32] + item.begin;  // this item's segment of the rounds
    const int nr = item.end - item.begin;

    if (warp >= WARPS) {
      // ---------------- producers: warp 8 copies the 32 phi rows, warp 9 the 32 G rows of every staged group
      const double* src_base = tile + (warp == WARPS ? 0 : 4 * comp_stride) + (size_t)(lane >> 2) * BP + (lane & 3) * 2;
      const int dst_off = ((warp - WARPS) * 32 + (lane >> 2)) * PSTRIDE + (((lane & 3) * 2) ^ (4 * ((lane >> 3) & 1)));
      for (int r = 0; r < nr; ++r) {
        const uint2 grp = __ldg(reinterpret_cast<const uint2*>(rounds[r].group));
        const int ng = rounds[r].ngroups;
        for (int kc = 0; kc < NKC; ++kc) {
          if (pass > 0) mbar_wait(empty + stage, (pass - 1) & 1);
          double* st = stage_base + stage * PSTAGE_ELEMS + dst_off;
          const double* src = src_base + kc * TKP;
#pragma unroll
          for (int i = 0; i < MAXG; ++i)
            if (i < ng) {
              const unsigned gi = ((i < 4 ? grp.x : grp.y) >> (8 * (i & 3))) & 0xffu;
              const double* sg = src + (size_t)gi * (32 * BP);
#pragma unroll
              for (int t = 0; t < 4; ++t) cp_async16(st + i * PGROUP_ELEMS + t * 8 * PSTRIDE, sg + t * 8 * BP);
            }
          mbar_arrive_cp_async(full + stage);
          if (++stage == PSTAGES) {
            stage = 0;
            ++pass;
          }
        }
      }
    } else {
      // ---------------- DMMA warps
      const int* __restrict__ sig = plan.sig_bf + (size_t)q * plan.nbf_pad;
      const int s8 = (s + 7) & ~7;  // rows beyond s rounded up to 8 are zero padding: their 8 x 8 fragments are skipped
      double acc[4][4][2];
      for (int r = 0; r < nr; ++r) {
        const ScatterRound* rd = rounds + r;
        const int slot_a = rd->ta[warp], slot_b = rd->tb[warp], kmask = rd->kmask[warp];
        const bool active = slot_a != 0xff;
        // valid 8-row fragments of the row group I and the column group J (only the last group of a block is short)
        const int mfr = active ? min(4, (s8 - rd->group[slot_a] * 32) >> 3) : 0;
        const int nfr = active ? min(4, (s8 - rd->group[slot_b] * 32) >> 3) : 0;
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int nn = 0; nn < 4; ++nn) acc[m][nn][0] = acc[m][nn][1] = 0.0;
        if (active && slot_a == slot_b && mfr == 4) {  // full diagonal tile
          vmat_round<4, 4, true>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane);
        } else if (mfr == 4 || !active) {
          switch (active ? nfr : 4) {
            case 1: vmat_round<4, 1>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
            case 2: vmat_round<4, 2>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
            case 3: vmat_round<4, 3>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
            default: vmat_round<4, 4>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
          }
        } else {  // a short row group is the last one, so the tile is the last diagonal tile: nfr == mfr
          switch (mfr) {
            case 1: vmat_round<1, 1, true>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
            case 2: vmat_round<2, 2, true>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
            default: vmat_round<3, 3, true>(acc, stage_base, full, empty, stage, pass, slot_a, slot_b, kmask, active, lane); break;
          }
        }
        if (active) {
          // V += Proj V_s Proj^T (:301), upper triangle (compact i <= j <=> global sig[i] <= sig[j])
          const int i0 = rd->group[slot_a] * 32, j0 = rd->group[slot_b] * 32;
#pragma unroll
          for (int nn = 0; nn < 4; ++nn)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = j0 + nn * 8 + 2 * lc + e;
              if (j >= s) continue;
              const size_t col = (size_t)sig[j] * nbf;
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                const int i = i0 + m * 8 + lr;
                if (i <= j) atomicAdd(W + col + sig[i], acc[m][nn][e]);
              }
            }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// K4-AB: two-basis scatter: ScalarOperatorToMatrixAdder::addBlock for basis A != basis B
// (ScalarOperatorToMatrixAdder.cpp:216-220 LDA, :286-300 GGA), the operator of ABFuncPotential / ABNAddFuncPotential:
//     m_AB += pA [ phi_A^T diag(a) phi_B + phi_A^T grad_B + grad_A^T phi_B ] pB^T,   grad_X = sum_c diag(b_c) d_c phi_X.
// With G_A = grad_A (k_form_g, a_scale = 0) and G_B = a phi_B + grad_B (a_scale = 1) every 32 x 32 tile is the same stacked
// product as in k_vmat,  U[I, J] = [phi_A,I | G_A,I] . [G_B,J | phi_B,J]^T, over the full s_A x s_B rectangle.  Rounds are
// formed on the fly: a band of two A row groups against four B column groups = 8 warp tiles on 6 staged groups.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(scat::THREADS, 2)
k_vmat_ab(PlanView planA, PlanView planB, int nbfA, const int* __restrict__ order, int nitems, int* __restrict__ counter,
          const int* __restrict__ skipA, const int* __restrict__ skipB, const double* __restrict__ phiA,
          const double* __restrict__ phiB, double* __restrict__ W) {
  using namespace scat;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  __shared__ int s_next;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int wi = warp >> 2, wj = warp & 3;  // this warp's A slot (0, 1) and B slot (2 + wj)

  for (;;) {
    __syncthreads();  // the ring of the previous block is no longer read
    if (tid == 0) s_next = atomicAdd(counter, 1);
    __syncthreads();
    const int qi = s_next;
    if (qi >= nitems) break;
    const int q = order[qi];  // both plans own the same blocks in the same slots
    if (skipA[q] || skipB[q]) continue;
    const int sA = planA.s[q], sB = planB.s[q];
    const int spA = planA.s_pad[q], spB = planB.s_pad[q];
    const double* __restrict__ tA = phiA + planA.phi_off[q];
    const double* __restrict__ tB = phiB + planB.phi_off[q];
    const size_t gA = (size_t)4 * spA * BP, gB = (size_t)4 * spB * BP;  // offset of the G slot
    const int* __restrict__ sigA = planA.sig_bf + (size_t)q * planA.nbf_pad;
    const int* __restrict__ sigB = planB.sig_bf + (size_t)q * planB.nbf_pad;
    const int nA32 = spA >> 5, nB32 = spB >> 5;
    const int njb = (nB32 + 3) >> 2;
    const int nr = ((nA32 + 1) >> 1) * njb;

    // producer: thread (rr, c16) copies 16 B of row rr (32 phi rows, then 32 G rows) of every staged group
    const int rr = tid >> 2, c16 = tid & 3;
    const size_t row_off = (size_t)(rr & 31) * BP + c16 * 2;
    const double* srcA = tA + ((rr & 32) ? gA : 0) + row_off;
    const double* srcB = tB + ((rr & 32) ? gB : 0) + row_off;
    const int dst_off = rr * STRIDE + c16 * 2;
    int is_r = 0, is_kc = 0, is_stage = 0;
    auto issue = [&]() {
      if (is_r < nr) {
        const int ib = is_r / njb, jb = is_r - ib * njb;
        double* st = stage_base + is_stage * STAGE_ELEMS + dst_off;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (2 * ib + i < nA32) cp_async16(st + i * GROUP_ELEMS, srcA + (size_t)(2 * ib + i) * (32 * BP) + is_kc * TKP);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (4 * jb + j < nB32)
            cp_async16(st + (2 + j) * GROUP_ELEMS, srcB + (size_t)(4 * jb + j) * (32 * BP) + is_kc * TKP);
        if (++is_kc == NKC) {
          is_kc = 0;
          ++is_r;
        }
        is_stage = (is_stage + 1 == STAGES) ? 0 : is_stage + 1;
      }
      cp_async_commit();
    };

    double acc[4][4][2];
    issue();
    issue();
    int c_stage = 0;
    for (int r = 0; r < nr; ++r) {
      const int ib = r / njb, jb = r - ib * njb;
      const int gI = 2 * ib + wi, gJ = 4 * jb + wj;
      const bool active = gI < nA32 && gJ < nB32;
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) acc[m][nn][0] = acc[m][nn][1] = 0.0;
      for (int kc = 0; kc < NKC; ++kc) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue();
        const double* st = stage_base + c_stage * STAGE_ELEMS;
        c_stage = (c_stage + 1 == STAGES) ? 0 : c_stage + 1;
        if (active) {
          const double* sI = st + wi * GROUP_ELEMS;        // phi_A rows 0..31, G_A rows 32..63
          const double* sJ = st + (2 + wj) * GROUP_ELEMS;  // phi_B rows 0..31, G_B rows 32..63
#pragma unroll
          for (int ks = 0; ks < TKP / 4; ++ks) {
            double a1[4], a2[4], b1[4], b2[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int o = (m * 8 + lr) * STRIDE + ks * 4 + lc;
              a1[m] = sI[o];
              a2[m] = sI[o + 32 * STRIDE];
              b2[m] = sJ[o];
              b1[m] = sJ[o + 32 * STRIDE];
            }
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
              for (int nn = 0; nn < 4; ++nn) {
                dmma884(acc[m][nn][0], acc[m][nn][1], a1[m], b1[nn]);
                dmma884(acc[m][nn][0], acc[m][nn][1], a2[m], b2[nn]);
              }
          }
        }
      }
      if (active) {  // m_AB += pA U pB^T
        const int i0 = gI * 32, j0 = gJ * 32;
#pragma unroll
        for (int nn = 0; nn < 4; ++nn)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = j0 + nn * 8 + 2 * lc + e;
            if (j >= sB) continue;
            const size_t col = (size_t)sigB[j] * nbfA;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int i = i0 + m * 8 + lr;
              if (i < sA) atomicAdd(W + col + sigA[i], acc[m][nn][e]);
            }
          }
      }
    }
    cp_async_wait<0>();
  }
}

// V[j,i] = V[i,j] for i < j (column-major, upper triangle holds the sums)
__global__ void k_mirror(int nbf, double* __restrict__ V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < nbf && j < nbf && i < j) V[j + (size_t)i * nbf] = V[i + (size_t)j * nbf];
}

// The tail of a build in one launch: nmat matrices (back to back) get their strict upper triangle mirrored down, and CTA (0, 0, 0)
// sums nred arrays of per-block partial sums in a fixed order (deterministic): partial array r goes to
// out[(r >> 1) * pair_stride + (r & 1)] (E and N of the KS build; E[rho_tot], E[rho_act] per functional of the NAdd build).
__global__ void __launch_bounds__(256)
k_finish(int nbf, double* __restrict__ V, const double* __restrict__ part, int nlit, int nred, int pair_stride,
         double* __restrict__ out) {
  __shared__ double scratch[32];
  double* Vm = V + (size_t)blockIdx.z * nbf * nbf;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < nbf && j < nbf && i < j) Vm[j + (size_t)i * nbf] = Vm[i + (size_t)j * nbf];
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    for (int r = 0; r < nred; ++r) {
      double sum = 0.0;
      for (int k = t; k < nlit; k += 256) sum += part[(size_t)r * nlit + k];
      const int lane = t & 31, wid = t >> 5;
      sum = warp_sum(sum);
      __syncthreads();
      if (lane == 0) scratch[wid] = sum;
      __syncthreads();
      if (t == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += scratch[w];
        out[(size_t)(r >> 1) * pair_stride + (r & 1)] = tot;
      }
    }
  }
}

}  // namespace sxc
