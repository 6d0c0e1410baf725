// density_kernel.cuh - rho and grad rho on a block: B = phi_s P_s on the FP64 tensor cores, fused row sums.
//
// Row 8a-2 of SURVEY.md: MatrixOperatorToGridTransformer::transform
// (src/data/grid/MatrixOperatorToGridTransformer.cpp:103-165):
//     phi_s = phi Proj,  P_s = Proj^T P Proj,  B = phi_s P_s,
//     rho_p = sum_mu B_pmu phi_pmu,   grad rho_p = 2 sum_mu B_pmu grad phi_pmu.
// B200 design: one CTA (8 DMMA warps = 4 point groups x 2 function groups, warp tile 32 points x 32 functions, plus two
// producer warps; two CTAs per SM) per 128-point block.  The product runs as DMMA m8n8k4 tiles (mma.sync ... f64, the only
// FP64 tensor path on sm_100a): per j-tile of 64 functions the K loop streams 16-function chunks of the phi tile
// (cp.async, 16 B) and the matching gathered P_s chunk (cp.async, 8 B, straight from the L2-resident nb x nb matrix
// through the block's compact->function map) through a 4-stage shared-memory ring; B never leaves registers - the
// epilogue multiplies the accumulator fragments with phi / grad phi rows that the producers stream through the same
// ring and reduces over functions with warp shuffles.  Tiles are stored with s padded to 32, the product runs over s
// rounded up to 8; a last j-tile with <= 4 fragments of 8 functions is K-split over the two function-group warps (rho is
// linear in B, so the halves need no extra reduction).  (dens::THREADS / STAGES / B_STRIDE describe the barrier-synchronised ring that
// k_grad_contract, gradient_kernels.cuh, still uses.)
#pragma once

#include "sxc_common.cuh"

namespace sxc {

namespace dens {
constexpr int THREADS = 256;
constexpr int TJ = 64;        // functions per j-tile
constexpr int TK = 16;        // functions per K chunk
constexpr int A_STRIDE = BP + 4;   // 132: conflict-free fragment loads (stride = 4 mod 16 doubles)
constexpr int B_STRIDE = TK + 4;   // 20
constexpr int STAGES = 3;
constexpr int A_ELEMS = TK * A_STRIDE;   // 2112 doubles
constexpr int B_ELEMS = TJ * B_STRIDE;   // 1280 doubles
constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
constexpr int NJW = 2;
constexpr int CWARPS = 8;                           // DMMA warps of the producer-warp pipeline
constexpr int PTHREADS = (CWARPS + 2) * 32;         // + two producer warps
// producer-warp kernel: the P_s chunk is stored dense and XOR-swizzled (column j holds k at j * 16 + (k ^ 4 (j & 3)): the
// fragment loads stay conflict free without padding columns), which makes room for a fourth stage
constexpr int PB_STRIDE = TK;
constexpr int PB_ELEMS = TJ * PB_STRIDE;            // 1024 doubles
constexpr int PSTAGE_ELEMS = A_ELEMS + PB_ELEMS;    // 3136 doubles = 24.5 KB
constexpr int PSTAGES = 4;
constexpr size_t smem_bytes_pipe(int s_pad_max) {
  return (size_t)PSTAGES * PSTAGE_ELEMS * sizeof(double) + (size_t)NJW * BP * 4 * sizeof(double) +
         (((size_t)(s_pad_max + TJ) * sizeof(int) + 7) & ~(size_t)7) + 2 * PSTAGES * sizeof(uint64_t);
}
constexpr size_t smem_bytes(int s_pad_max) {
  return (size_t)STAGES * STAGE_ELEMS * sizeof(double) + (size_t)NJW * BP * 4 * sizeof(double) +
         (size_t)(s_pad_max + TJ) * sizeof(int);
}
}  // namespace dens

// K loop of one j-tile for a DMMA warp that owns NFRAG (1..4) fragments of 8 functions: compile-time fragment counts keep
// the DMMA stream free of predicates.
template <int NFRAG>
__device__ __forceinline__ void dens_kloop(double (&acc)[4][4][2], const double* __restrict__ stage_base, uint64_t* full,
                                           uint64_t* empty, int& stage, int& pass, int nk, int s8, bool split, int jw,
                                           int cg, int pw, int lane) {
  using namespace dens;
  const int lr = lane >> 2, lc = lane & 3;
  for (int kc = 0; kc < nk; ++kc) {
    mbar_wait(full + stage, pass & 1);
    const double* As = stage_base + stage * PSTAGE_ELEMS;
    const double* Bs = As + A_ELEMS;
    const int kvalid = min(TK, s8 - kc * TK) >> 2;  // k-steps of this chunk that exist
#pragma unroll
    for (int ks = 0; ks < TK / 4; ++ks) {
      if (ks >= kvalid || (split && (ks & 1) != jw)) continue;
      double a[4], bfrag[NFRAG];
#pragma unroll
      for (int m = 0; m < 4; ++m) a[m] = As[(ks * 4 + lc) * A_STRIDE + pw * 32 + m * 8 + lr];
#pragma unroll
      for (int nn = 0; nn < NFRAG; ++nn)
        bfrag[nn] = Bs[(cg * 32 + nn * 8 + lr) * PB_STRIDE + ((ks * 4 + lc) ^ (4 * (lr & 3)))];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nn = 0; nn < NFRAG; ++nn) dmma884(acc[m][nn][0], acc[m][nn][1], a[m], bfrag[nn]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    if (++stage == PSTAGES) {
      stage = 0;
      ++pass;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Warps 8 and 9 are dedicated producers (warp 8: phi rows, 32 x 16 B per lane and chunk; warp 9: the gathered P_s chunk,
// 32 x 8 B per lane) whose copies a "full" mbarrier tracks (cp.async.mbarrier.arrive.noinc); the 8 DMMA warps wait on
// "full", multiply and arrive on "empty" - no CTA-wide barrier and no producer code in their K loop.  The epilogue
// operands ride the same ring: after the K chunks of a j-tile the producers stream its phi / grad phi rows (per
// component, 16 rows x 128 points per chunk) through the A part of the stages, so the DMMA warps read them from shared
// memory instead of keeping global loads in flight in registers.  History (profiles/r01_ncu_source_density.md): a
// barrier-synchronised ring whose 8 warps all issued the copies and whose epilogue read its operands with 128 global
// loads per thread took 2.78 ms (23 % of the warp time at the barrier / in producer code, 0.37 ms in register-limited
// loads); this kernel takes 2.61 ms.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(dens::PTHREADS, 2)
k_density(GridView g, PlanView plan, int nbf, const double* __restrict__ P, const WorkItem* __restrict__ items,
          const double* __restrict__ phi_buf, double* __restrict__ rho, double* __restrict__ gx,
          double* __restrict__ gy, double* __restrict__ gz, int* __restrict__ nonneg, int a_slot, int e_slot0,
          int epi_prefetch) {  // bit 0: L2 prefetch of the epilogue rows; bit 1 (development): epilogue of phi only
  // a_slot / e_slot0: tile slots of the A operand of the product and of the first epilogue component (0 / 0 for the density and
  // its gradient; the second derivatives of the density contract other slots of an 8-slot gradient plan, sxc_density_hessian_on_grid)
  using namespace dens;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  double* red = stage_base + PSTAGES * PSTAGE_ELEMS;        // [NJW][4][128]
  int* sig = reinterpret_cast<int*>(red + NJW * BP * 4);    // [s_pad + TJ]

  const WorkItem item = items[blockIdx.x];
  const int q = item.q;
  const int blk = plan.block_id[q];
  const long first = (long)blk * g.blocksize;
  const int n = (int)min((long)g.blocksize, g.npts - first);
  const int s = plan.s[q];
  const int sp = plan.s_pad[q];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (nonneg && tid == 0) nonneg[blk] = s > 0;
  if (s == 0) {  // MatrixOperatorToGridTransformer.cpp:117-126: outputs stay zero
    if (tid < n) {
      rho[first + tid] = 0.0;
      if (gx) {
        gx[first + tid] = 0.0;
        gy[first + tid] = 0.0;
        gz[first + tid] = 0.0;
      }
    }
    return;
  }
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sig) +
                                               (((size_t)(sp + TJ) * sizeof(int) + 7) & ~(size_t)7));
  uint64_t* empty = full + PSTAGES;
  const double* __restrict__ tile = phi_buf + plan.phi_off[q];
  const size_t comp_stride = (size_t)sp * BP;
  const int* __restrict__ sig_g = plan.sig_bf + (size_t)q * plan.nbf_pad;
  for (int c = tid; c < sp + TJ; c += PTHREADS) sig[c] = c < sp ? sig_g[c] : 0;
  for (int i = tid; i < NJW * BP * 4; i += PTHREADS) red[i] = 0.0;
  if (tid < PSTAGES) {
    mbar_init(full + tid, 64);       // the 64 producer lanes
    mbar_init(empty + tid, CWARPS);  // one arrive per DMMA warp
  }
  __syncthreads();

  // The tile is stored with s padded to 32 (zero rows), but the product only runs over s8 = s rounded up to 8: whole
  // 8-function fragments and 4-function k-steps beyond it are skipped (executed / algorithmic flops 1.10 -> ~1.03).
  const int s8 = (s + 7) & ~7;
  const int nk = (s8 + TK - 1) / TK;          // K chunks per j-tile (the last one may hold 8 functions)
  const int njt_all = (s8 + TJ - 1) / TJ;     // j-tiles of 64 of the block (= ceil(s_pad / 64), the host's count)
  const int jt_begin = item.begin, njt = item.end;  // this CTA's segment of them (normally all)
  const bool partial = jt_begin != 0 || njt != njt_all;
  const int ncomp = (gx && !(epi_prefetch & 2)) ? 4 : 1;
  int stage = 0, pass = 0;  // ring position; every warp walks the same chunk sequence

  if (warp == CWARPS) {
    // ---------------- producer A: phi rows of the K chunks, then the epilogue rows of the j-tile
    const int a_off = lane * 2;  // 16-byte piece t * 32 + lane of a chunk: row t / 2, doubles (t & 1) * 64 + 2 lane ..
    // 16 rows x 128 points -> A part of the stage.  Epilogue chunks are stored with the rows of every group of 8 in the order
    // 0 2 4 6 1 3 5 7: a DMMA lane lc owns rows 2 lc + e, and consecutive slots keep its loads bank-conflict free.
    auto copy_rows = [&](const double* src, bool epi) {
      if (pass > 0) mbar_wait(empty + stage, (pass - 1) & 1);
      double* st = stage_base + stage * PSTAGE_ELEMS + a_off;
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const int row = t >> 1;
        const int slot = epi ? ((row & 8) | ((row & 1) << 2) | ((row >> 1) & 3)) : row;
        cp_async16(st + slot * A_STRIDE + (t & 1) * 64, src + a_off + row * BP + (t & 1) * 64);
      }
      mbar_arrive_cp_async(full + stage);
      if (++stage == PSTAGES) {
        stage = 0;
        ++pass;
      }
    };
    // The ring keeps 3 stages (48 KB) in flight ahead of the DMMA warps: enough for the K chunks, which are multiplied for ~1 us
    // each, but the epilogue stages are consumed in a fraction of that and would run at HBM latency (16 stages per j-tile).  Their
    // rows are therefore pulled into L2 one component ahead with bulk prefetches (64 KB each, one instruction of one lane).
    auto prefetch_epi = [&](int jt, int comp, int nrg) {
      if ((epi_prefetch & 1) && lane == 0 && comp < ncomp)
        bulk_prefetch_l2(tile + (e_slot0 + comp) * comp_stride + (size_t)(jt * TJ) * BP, (unsigned)(nrg * TK * BP * sizeof(double)));
    };
    for (int jt = jt_begin; jt < njt; ++jt) {
      const int nrg = (min(TJ, s8 - jt * TJ) + TK - 1) / TK;
      const int kpf = max(0, nk - 6);  // a few K chunks before the epilogue starts
      for (int kc = 0; kc < nk; ++kc) {
        if (kc == kpf) {
          prefetch_epi(jt, 0, nrg);
          prefetch_epi(jt, 1, nrg);
        }
        copy_rows(tile + a_slot * comp_stride + (size_t)kc * (TK * BP), false);
      }
      for (int comp = 0; comp < ncomp; ++comp) {
        prefetch_epi(jt, comp + 2, nrg);
        for (int rg = 0; rg < nrg; ++rg)
          copy_rows(tile + (e_slot0 + comp) * comp_stride + (size_t)(jt * TJ + rg * TK) * BP, true);
      }
    }
  } else if (warp == CWARPS + 1) {
    // ---------------- producer B: the gathered P_s chunk ("Proj^T P Proj" without materialising it)
    const int bk = lane & (TK - 1), bj0 = lane >> 4;  // gather element t * 32 + lane: k = bk, j = 2 t + bj0
    for (int jt = jt_begin; jt < njt; ++jt) {
      const int ncol = min(TJ, s8 - jt * TJ);
      int colbase[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) colbase[t] = sig[jt * TJ + 2 * t + bj0] * nbf;
      for (int c = 0; c < nk + ncomp * ((ncol + TK - 1) / TK); ++c) {
        if (pass > 0) mbar_wait(empty + stage, (pass - 1) & 1);
        if (c < nk) {
          const double* prow = P + sig[c * TK + bk];
          double* bdst = stage_base + stage * PSTAGE_ELEMS + A_ELEMS + bj0 * PB_STRIDE;
#pragma unroll
          for (int t = 0; t < 32; ++t)  // column j = 2 t + bj0, swizzle 4 (j & 3) = 4 ((2 t & 3) + bj0)
            if (2 * t + bj0 < ncol) cp_async8(bdst + 2 * t * PB_STRIDE + (bk ^ (4 * (((2 * t) & 3) + bj0))), prow + colbase[t]);
        }
        mbar_arrive_cp_async(full + stage);  // (fires at once for an epilogue chunk: nothing of this warp is in flight)
        if (++stage == PSTAGES) {
          stage = 0;
          ++pass;
        }
      }
    }
  } else {
    // ---------------- DMMA warps: 4 point groups x 2 function groups
    const int pw = warp & 3, jw = warp >> 2;
    const int lr = lane >> 2, lc = lane & 3;
    double acc[4][4][2];
    for (int jt = jt_begin; jt < njt; ++jt) {
      // fragments (8 functions) of this j-tile; with <= 4 of them both function-group warps work on the same
      // fragments, on alternating k-steps
      const int nf = min(8, (s8 - jt * TJ) >> 3);
      const bool split = nf <= 4;
      const int cg = split ? 0 : jw;
      const int nfrag = split ? nf : min(4, nf - 4 * jw);
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) acc[m][nn][0] = acc[m][nn][1] = 0.0;
      switch (nfrag) {
        case 1: dens_kloop<1>(acc, stage_base, full, empty, stage, pass, nk, s8, split, jw, cg, pw, lane); break;
        case 2: dens_kloop<2>(acc, stage_base, full, empty, stage, pass, nk, s8, split, jw, cg, pw, lane); break;
        case 3: dens_kloop<3>(acc, stage_base, full, empty, stage, pass, nk, s8, split, jw, cg, pw, lane); break;
        default: dens_kloop<4>(acc, stage_base, full, empty, stage, pass, nk, s8, split, jw, cg, pw, lane); break;
      }
      // epilogue: rho += B o phi, grad rho += B o grad phi   (MatrixOperatorToGridTransformer.cpp:158-163); chunk rg of a
      // component holds rows rg * 16 .. + 15 of the j-tile = the fragments nn = 2 (rg & 1), + 1 of column group rg / 2
      const int nrg = (min(TJ, s8 - jt * TJ) + TK - 1) / TK;
      for (int comp = 0; comp < ncomp; ++comp) {
        double r[4] = {0.0, 0.0, 0.0, 0.0};
        for (int rg = 0; rg < nrg; ++rg) {
          mbar_wait(full + stage, pass & 1);
          if ((rg >> 1) == cg) {
            const double* Es = stage_base + stage * PSTAGE_ELEMS + lc * A_STRIDE + pw * 32 + lr;  // row 2 lc + e + 8 h sits in slot lc + 4 e + 8 h
            if (rg & 1) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                  for (int m = 0; m < 4; ++m) r[m] += acc[m][2 + h][e] * Es[(h * 8 + e * 4) * A_STRIDE + m * 8];
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                  for (int m = 0; m < 4; ++m) r[m] += acc[m][h][e] * Es[(h * 8 + e * 4) * A_STRIDE + m * 8];
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty + stage);
          if (++stage == PSTAGES) {
            stage = 0;
            ++pass;
          }
        }
        // reduce over the 4 lanes of a fragment row.  FP64 vector instructions queue behind the DMMAs of the SM's other warps (a
        // DADD / DFMA waits ~4x as long as a DMMA in the stall samples of profiles/r02_ncu_full_water64.md), so the four sums are
        // reduced as a 4 x 4 transpose - lane lc ends up with the total of m = lc: 3 DADD + 3 shuffles and one update of the CTA
        // scratch by every lane, instead of 8 + 8 and four predicated updates.  Same additions in the same order (pairs lc ^ 1,
        // then lc ^ 2): bit-identical results.
        {
          const bool b0 = lc & 1, b1 = lc & 2;
          const double k0 = b0 ? r[1] : r[0], k1 = b0 ? r[3] : r[2];
          const double s0 = b0 ? r[0] : r[1], s1 = b0 ? r[2] : r[3];
          const double a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 1);  // m = b0,     lanes lc, lc ^ 1
          const double a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 1);  // m = 2 + b0
          const double t = (b1 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, b1 ? a0 : a1, 2);  // m = lc, all four lanes
          red[(jw * 4 + comp) * BP + pw * 32 + lc * 8 + lr] += t;  // [jw][comp][point]: the 32 lanes hit 32 consecutive doubles
        }
      }
    }
  }
  __syncthreads();
  if (tid < n) {
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
    for (int jj = 0; jj < NJW; ++jj) {  // fixed order over the function-group warps
      const double* a = red + (size_t)jj * 4 * BP + tid;
      r0 += a[0];
      r1 += a[BP];
      r2 += a[2 * BP];
      r3 += a[3 * BP];
    }
    if (partial) {  // several CTAs share the block: the (pre-zeroed) outputs are accumulated
      atomicAdd(rho + first + tid, r0);
      if (gx) {
        atomicAdd(gx + first + tid, 2.0 * r1);
        atomicAdd(gy + first + tid, 2.0 * r2);
        atomicAdd(gz + first + tid, 2.0 * r3);
      }
    } else {
      rho[first + tid] = r0;
      if (gx) {
        gx[first + tid] = 2.0 * r1;
        gy[first + tid] = 2.0 * r2;
        gz[first + tid] = 2.0 * r3;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// k_density_tma: the same contraction with the phi operand staged by the TMA unit.
//   * Warp 8: ONE elected thread arms the stage's "full" mbarrier (mbarrier.arrive.expect_tx, 16 KB) and issues eight
//     cp.async.bulk.tensor.2d boxes of 16 points x 16 function rows (128-byte rows, SWIZZLE_128B) per chunk from the 2-D tensor
//     map over the tile workspace - for the K chunks of the product and for the phi / grad phi rows of the epilogue alike.
//     The 16 rows of a chunk are multiplied in the order 0 2 4 6 | 1 3 5 7 | 8 10 12 14 | 9 11 13 15 (one k-step each): with the
//     hardware swizzle (16-byte chunk c of row r at c ^ (r & 7)) the 16 lanes of a half-warp then hit 16 different bank pairs -
//     conflict-free fragment loads without padding columns.
//   * Warp 9 still gathers the P_s chunk ("Proj^T P Proj" without materialising it) with 8-byte cp.async; it writes row k of a
//     column at the position dens_kperm(k), so the DMMA warps read both operands with the same k-step order.
//   * DMMA warps, accumulators, epilogue reduction and the split / partial logic are those of k_density.
// ------------------------------------------------------------------------------------------------------------
namespace dens {
constexpr int TA_ELEMS = TK * BP;                   // 2048 doubles: eight boxes of 16 points x 16 rows
constexpr int TSTAGE_ELEMS = TA_ELEMS + PB_ELEMS;   // 3072 doubles = 24 KB
constexpr int TBOX = 16 * 16;                       // doubles per box
constexpr size_t smem_bytes_tma(int s_pad_max) {
  return (size_t)PSTAGES * TSTAGE_ELEMS * sizeof(double) + (size_t)NJW * BP * 4 * sizeof(double) +
         (((size_t)(s_pad_max + TJ) * sizeof(int) + 7) & ~(size_t)7) + 2 * PSTAGES * sizeof(uint64_t) + 1024;
}
}  // namespace dens

// offset (doubles) of (function row f of the chunk, point p of the block) inside the A part of a stage
__device__ __forceinline__ int dens_tma_off(int f, int p) {
  return (p >> 4) * dens::TBOX + f * 16 + (((((p & 15) >> 1) ^ (f & 7)) << 1) | (p & 1));
}

// Fragment addressing of the swizzled A part.  A 64-bit shared load is served per half-warp (16 lanes = 4 fragment rows lr x 4
// fragment columns lc): it is conflict free iff those 16 lanes hit 16 different 8-byte bank pairs.  With the hardware swizzle
// (16-byte chunk c of row f at c ^ (f & 7)) that needs the four rows of a k-step to differ in bits 1..2 of f, so k-step
// ks = 2 h + e multiplies the rows f = 8 h + 2 lc + e (0 2 4 6 | 1 3 5 7 | 8 10 12 14 | 9 11 13 15) - which is also the row a
// lane's accumulator column 2 lc + e belongs to in the epilogue, so both loops share the constants.  For row f and point
// p = 32 pw + 8 m + lr:   dens_tma_off(f, p) = 512 pw + 32 lc + (lr & 1)             [lane base]
//                                            + 128 h + 16 e + 2 ((lr >> 1) ^ (2 (lc & 1) + e))   [eoff(e) + 128 h]
//                                            + 256 (m >> 1) + 8 ((m & 1) ^ (lc >> 1))            [immediates after the parity of
//                                                                                                lc >> 1 picked a base pointer]
// A last chunk of 8 rows simply stops after k-steps 0 and 1.
struct DensLane {
  int base_even, base_odd;  // lane base for even / odd fragment rows m
  int eoff[2];
};
__device__ __forceinline__ DensLane dens_lane(int pw, int lr, int lc) {
  DensLane d;
  const int base = pw * 512 + lc * 32 + (lr & 1);
  d.base_even = base + 8 * (lc >> 1);
  d.base_odd = base + 8 * (1 - (lc >> 1));
  d.eoff[0] = 2 * ((lr >> 1) ^ (2 * (lc & 1)));
  d.eoff[1] = 16 + 2 * ((lr >> 1) ^ (2 * (lc & 1) + 1));
  return d;
}
// position of row k of a chunk inside a column of the gathered P_s chunk: the DMMA lane (lc, k-step ks) reads position
// 4 ks + lc and must find row 8 (ks >> 1) + 2 lc + (ks & 1) there
__device__ __forceinline__ int dens_kperm(int k) { return 4 * ((k & 1) + 2 * (k >> 3)) + ((k & 7) >> 1); }

template <int NFRAG>
__device__ __forceinline__ void dens_kloop_tma(double (&acc)[4][4][2], const double* __restrict__ stage_base, uint64_t* full,
                                               uint64_t* empty, int& stage, int& pass, int nk, int s8, bool split, int jw,
                                               int cg, const DensLane& dl, int lane) {
  using namespace dens;
  const int lr = lane >> 2, lc = lane & 3;
  const int boff = (cg * 32 + lr) * PB_STRIDE;
  const int bswz = 4 * (lr & 3);
  for (int kc = 0; kc < nk; ++kc) {
    mbar_wait(full + stage, pass & 1);
    const double* As = stage_base + stage * TSTAGE_ELEMS;
    const double* Bs = As + TA_ELEMS + boff;
    const double* Ae = As + dl.base_even;
    const double* Ao = As + dl.base_odd;
    const int nks = (s8 - kc * TK > 8) ? 4 : 2;  // k-steps of this chunk that hold functions
#pragma unroll
    for (int ks = 0; ks < TK / 4; ++ks) {
      if (ks >= nks || (split && (ks & 1) != jw)) continue;
      double a[4], bfrag[NFRAG];
      const int o = dl.eoff[ks & 1] + 128 * (ks >> 1);
      a[0] = Ae[o];
      a[1] = Ao[o];
      a[2] = Ae[o + 256];
      a[3] = Ao[o + 256];
#pragma unroll
      for (int nn = 0; nn < NFRAG; ++nn) bfrag[nn] = Bs[nn * 8 * PB_STRIDE + ((ks * 4 + lc) ^ bswz)];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nn = 0; nn < NFRAG; ++nn) dmma884(acc[m][nn][0], acc[m][nn][1], a[m], bfrag[nn]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    if (++stage == PSTAGES) {
      stage = 0;
      ++pass;
    }
  }
}

__global__ void __launch_bounds__(dens::PTHREADS, 2)
k_density_tma(const __grid_constant__ CUtensorMap tmap, GridView g, PlanView plan, int nbf, const double* __restrict__ P,
              const WorkItem* __restrict__ items, double* __restrict__ rho, double* __restrict__ gx, double* __restrict__ gy,
              double* __restrict__ gz, int* __restrict__ nonneg) {
  using namespace dens;
  extern __shared__ unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(
      smem_raw + ((1024u - (static_cast<unsigned>(__cvta_generic_to_shared(smem_raw)) & 1023u)) & 1023u));
  double* red = stage_base + PSTAGES * TSTAGE_ELEMS;        // [NJW][128][4]
  int* sig = reinterpret_cast<int*>(red + NJW * BP * 4);    // [s_pad + TJ]

  const WorkItem item = items[blockIdx.x];
  const int q = item.q;
  const int blk = plan.block_id[q];
  const long first = (long)blk * g.blocksize;
  const int n = (int)min((long)g.blocksize, g.npts - first);
  const int s = plan.s[q];
  const int sp = plan.s_pad[q];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (nonneg && tid == 0) nonneg[blk] = s > 0;
  if (s == 0) {  // MatrixOperatorToGridTransformer.cpp:117-126: outputs stay zero
    if (tid < n) {
      rho[first + tid] = 0.0;
      if (gx) {
        gx[first + tid] = 0.0;
        gy[first + tid] = 0.0;
        gz[first + tid] = 0.0;
      }
    }
    return;
  }
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sig) +
                                               (((size_t)(sp + TJ) * sizeof(int) + 7) & ~(size_t)7));
  uint64_t* empty = full + PSTAGES;
  const int* __restrict__ sig_g = plan.sig_bf + (size_t)q * plan.nbf_pad;
  for (int c = tid; c < sp + TJ; c += PTHREADS) sig[c] = c < sp ? sig_g[c] : 0;
  for (int i = tid; i < NJW * BP * 4; i += PTHREADS) red[i] = 0.0;
  if (tid == 0) {
    for (int i = 0; i < PSTAGES; ++i) {
      mbar_init(full + i, 33);       // the TMA thread's arrive.expect_tx + the 32 gather lanes
      mbar_init(empty + i, CWARPS);  // one arrive per DMMA warp
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmap);
  }
  __syncthreads();

  const int s8 = (s + 7) & ~7;
  const int nk = (s8 + TK - 1) / TK;          // K chunks per j-tile
  const int njt_all = (s8 + TJ - 1) / TJ;     // j-tiles of 64 of the block
  const int jt_begin = item.begin, njt = item.end;  // this CTA's segment of them (normally all)
  const bool partial = jt_begin != 0 || njt != njt_all;
  const int ncomp = gx ? 4 : 1;
  int stage = 0, pass = 0;  // ring position; every warp walks the same chunk sequence

  if (warp == CWARPS) {
    // ---------------- producer A: one thread, eight TMA boxes per chunk
    if (lane == 0) {
      const int row0 = (int)(plan.phi_off[q] / BP);
      auto copy_rows = [&](int row) {
        if (pass > 0) mbar_wait(empty + stage, (pass - 1) & 1);
        double* st = stage_base + stage * TSTAGE_ELEMS;
        mbar_arrive_expect_tx(full + stage, (unsigned)(TA_ELEMS * sizeof(double)));
#pragma unroll
        for (int bx = 0; bx < 8; ++bx) tma_load_2d(st + bx * TBOX, &tmap, bx * 16, row, full + stage);
        if (++stage == PSTAGES) {
          stage = 0;
          ++pass;
        }
      };
      for (int jt = jt_begin; jt < njt; ++jt) {
        const int nrg = (min(TJ, s8 - jt * TJ) + TK - 1) / TK;
        for (int kc = 0; kc < nk; ++kc) copy_rows(row0 + kc * TK);
        for (int comp = 0; comp < ncomp; ++comp)
          for (int rg = 0; rg < nrg; ++rg) copy_rows(row0 + comp * sp + jt * TJ + rg * TK);
      }
    }
  } else if (warp == CWARPS + 1) {
    // ---------------- producer B: the gathered P_s chunk; row k of a column goes to the position of the permuted index
    const int bk = lane & (TK - 1), bj0 = lane >> 4;  // gather element t * 32 + lane: k = bk, j = 2 t + bj0
    const int kperm = dens_kperm(bk);
    for (int jt = jt_begin; jt < njt; ++jt) {
      const int ncol = min(TJ, s8 - jt * TJ);
      int colbase[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) colbase[t] = sig[jt * TJ + 2 * t + bj0] * nbf;
      for (int c = 0; c < nk + ncomp * ((ncol + TK - 1) / TK); ++c) {
        if (pass > 0) mbar_wait(empty + stage, (pass - 1) & 1);
        if (c < nk) {
          const double* prow = P + sig[c * TK + bk];
          double* bdst = stage_base + stage * TSTAGE_ELEMS + TA_ELEMS + bj0 * PB_STRIDE;
#pragma unroll
          for (int t = 0; t < 32; ++t)  // column j = 2 t + bj0, swizzle 4 (j & 3) = 4 ((2 t & 3) + bj0)
            if (2 * t + bj0 < ncol) cp_async8(bdst + 2 * t * PB_STRIDE + (kperm ^ (4 * (((2 * t) & 3) + bj0))), prow + colbase[t]);
        }
        mbar_arrive_cp_async(full + stage);  // (fires at once for an epilogue chunk: nothing of this warp is in flight)
        if (++stage == PSTAGES) {
          stage = 0;
          ++pass;
        }
      }
    }
  } else {
    // ---------------- DMMA warps: 4 point groups x 2 function groups
    const int pw = warp & 3, jw = warp >> 2;
    const int lr = lane >> 2, lc = lane & 3;
    const DensLane dl = dens_lane(pw, lr, lc);
    double acc[4][4][2];
    for (int jt = jt_begin; jt < njt; ++jt) {
      const int nf = min(8, (s8 - jt * TJ) >> 3);
      const bool split = nf <= 4;
      const int cg = split ? 0 : jw;
      const int nfrag = split ? nf : min(4, nf - 4 * jw);
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) acc[m][nn][0] = acc[m][nn][1] = 0.0;
      switch (nfrag) {
        case 1: dens_kloop_tma<1>(acc, stage_base, full, empty, stage, pass, nk, s8, split, jw, cg, dl, lane); break;
        case 2: dens_kloop_tma<2>(acc, stage_base, full, empty, stage, pass, nk, s8, split, jw, cg, dl, lane); break;
        case 3: dens_kloop_tma<3>(acc, stage_base, full, empty, stage, pass, nk, s8, split, jw, cg, dl, lane); break;
        default: dens_kloop_tma<4>(acc, stage_base, full, empty, stage, pass, nk, s8, split, jw, cg, dl, lane); break;
      }
      // epilogue: rho += B o phi, grad rho += B o grad phi   (MatrixOperatorToGridTransformer.cpp:158-163); chunk rg of a
      // component holds rows rg * 16 .. + 15 of the j-tile = the fragments nn = 2 (rg & 1), + 1 of column group rg / 2;
      // accumulator element (m, nn, e) belongs to function 8 nn + 2 lc + e and point pw * 32 + 8 m + lr
      const int nrg = (min(TJ, s8 - jt * TJ) + TK - 1) / TK;
      for (int comp = 0; comp < ncomp; ++comp) {
        double r[4] = {0.0, 0.0, 0.0, 0.0};
        for (int rg = 0; rg < nrg; ++rg) {
          mbar_wait(full + stage, pass & 1);
          if ((rg >> 1) == cg) {
            // row f = 8 h + 2 lc + e of the chunk, point 32 pw + 8 m + lr:  dens_tma_off(f, p) = 512 pw + 32 lc + (lr & 1)
            //   + 128 h + 16 e + 2 ((lr >> 1) ^ (2 (lc & 1) + e)) + 256 (m >> 1) + 8 ((m & 1) ^ (lc >> 1))
            const double* Es = stage_base + stage * TSTAGE_ELEMS;
            const double* Ee = Es + dl.base_even;
            const double* Eo = Es + dl.base_odd;
            if (rg & 1) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  r[0] += acc[0][2 + h][e] * Ee[dl.eoff[e] + 128 * h];
                  r[1] += acc[1][2 + h][e] * Eo[dl.eoff[e] + 128 * h];
                  r[2] += acc[2][2 + h][e] * Ee[dl.eoff[e] + 128 * h + 256];
                  r[3] += acc[3][2 + h][e] * Eo[dl.eoff[e] + 128 * h + 256];
                }
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  r[0] += acc[0][h][e] * Ee[dl.eoff[e] + 128 * h];
                  r[1] += acc[1][h][e] * Eo[dl.eoff[e] + 128 * h];
                  r[2] += acc[2][h][e] * Ee[dl.eoff[e] + 128 * h + 256];
                  r[3] += acc[3][h][e] * Eo[dl.eoff[e] + 128 * h + 256];
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
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          r[m] += __shfl_xor_sync(0xffffffffu, r[m], 1);
          r[m] += __shfl_xor_sync(0xffffffffu, r[m], 2);
          if (lc == 0) red[((size_t)jw * BP + pw * 32 + m * 8 + lr) * 4 + comp] += r[m];
        }
      }
    }
  }
  __syncthreads();
  if (tid < n) {
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
    for (int jj = 0; jj < NJW; ++jj) {  // fixed order over the function-group warps
      const double* a = red + ((size_t)jj * BP + tid) * 4;
      r0 += a[0];
      r1 += a[1];
      r2 += a[2];
      r3 += a[3];
    }
    if (partial) {  // several CTAs share the block: the (pre-zeroed) outputs are accumulated
      atomicAdd(rho + first + tid, r0);
      if (gx) {
        atomicAdd(gx + first + tid, 2.0 * r1);
        atomicAdd(gy + first + tid, 2.0 * r2);
        atomicAdd(gz + first + tid, 2.0 * r3);
      }
    } else {
      rho[first + tid] = r0;
      if (gx) {
        gx[first + tid] = 2.0 * r1;
        gy[first + tid] = 2.0 * r2;
        gz[first + tid] = 2.0 * r3;
      }
    }
  }
}

}  // namespace sxc
