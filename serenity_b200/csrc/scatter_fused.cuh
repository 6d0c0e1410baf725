// scatter_fused.cuh - k_vmat_fg: the scatter contraction with the formation of G hidden behind the tensor pipe.
//
// ScalarOperatorToMatrixAdder::addBlock (src/data/grid/ScalarOperatorToMatrixAdder.cpp:225-303) first forms the weighted
// operator on the block (a = w o v_rho, b = w o g, :253-260), tests its block average (:262-268), builds
// G = b . grad phi + 1/2 a phi (:276-281) and then contracts (:283-301).  k_form_g + k_vmat_tma do that in two launches: a pure
// HBM phase (reads four tile components, writes one) followed by a pure DMMA phase, each leaving the other unit of the SM idle.
// This kernel runs both phases of DIFFERENT blocks at the same time inside every persistent CTA:
//
//   warpgroup 0     (4 warps)  "G formers": pull the next work item from the device queue, form a, b and the block test of
//                              block n + 1, stream its phi / grad-phi rows through private per-warp TMA rings (4 stages of 2 rows
//                              x 32 points x 4 components, no swizzle: a lane reads consecutive points) and store G into the
//                              tile's fifth slot
//   warpgroups 1, 2 (8 warps)  DMMA warps of k_vmat_tma, unchanged: the rounds of block n
//   warpgroup 3     warp 12 lane 0: the TMA producer of the operand ring (phi and G boxes of block n); the rest exits
//
// Registers are moved between the warpgroups with setmaxnreg: the kernel is launched at 64 registers per thread (2 CTAs of 512
// threads per SM), the DMMA warpgroups grow to 96, the G formers shrink to 40, warpgroup 3 to 24.  The G of a block travels
// through HBM/L2 as before (the rounds of the schedule re-stage it), but the two CTAs of an SM and the two halves of a CTA keep
// the HBM pipe and the DMMA pipe busy together.  Item slots are double-buffered in shared memory: `ready[k]` (G formers ->
// consumers: item index, skip flag and G are valid) and `release[k]` (consumers -> G formers: the slot may be reused).
#pragma once

#include "scatter_tma.cuh"

namespace sxc {

namespace scat3 {
constexpr int WARPS = 8;    // DMMA warps
constexpr int HWARPS = 4;   // G formers
constexpr int THREADS = 512;
constexpr int PRODUCER_WARP = 12;
// every former warp owns 32 of the block's 128 points and a private TMA ring: a stage holds HROWS rows x 32 points of each of
// the four tile components (2 KB); the warps never wait for each other inside a block
constexpr int HROWS = 2;
constexpr int HSTAGES = 4;
constexpr int HPTS = 32;                            // points per former warp
constexpr int H_STAGE_ELEMS = 4 * HROWS * HPTS;     // phi, dx, dy, dz rows of one warp's stage
constexpr int H_WARP_ELEMS = HSTAGES * H_STAGE_ELEMS;
using C = scat2::Cfg<8, 3>;
constexpr int NBAR = 2 * C::STAGES + HWARPS * HSTAGES + 4;
constexpr size_t smem_bytes(int sig_cap) {
  return (size_t)C::STAGES * C::STAGE_ELEMS * sizeof(double) + (size_t)HWARPS * H_WARP_ELEMS * sizeof(double) +
         NBAR * sizeof(uint64_t) + 16 * sizeof(double) + (size_t)sig_cap * sizeof(int) + 1024;
}
}  // namespace scat3

template <int NREG>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(NREG));
}
template <int NREG>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(NREG));
}
// writes of this thread to global memory become visible to later TMA (async-proxy) reads ordered after it
__device__ __forceinline__ void fence_global_to_async_proxy() {
  __threadfence();
  asm volatile("fence.proxy.async;\n" ::: "memory");
}
// Named barrier among `nthreads` threads.  bar.sync is the .aligned form: every thread of a warp must execute it together, and a
// warp that reaches it in two pieces (lane 0 coming late out of an `if (ht == 0)` body with spin loops, which ptxas does not
// always re-join before an opaque asm statement) is counted twice - the barrier then opens before the late lane has written what
// the others are about to read.  So: re-converge the warp explicitly and use the form that counts threads, not warps.
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  __syncwarp();
  asm volatile("barrier.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(scat3::THREADS, 2)
k_vmat_fg(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_rows, GridView g, PlanView plan,
          int nbf, const WorkItem* __restrict__ items, int nitems, int* __restrict__ counter,
          const ScatterRound2* __restrict__ tpl, const int* __restrict__ tpl_off, int sig_cap, double block_ave_thr,
          double a_scale, const double* __restrict__ v_rho, const double* __restrict__ v_gx, const double* __restrict__ v_gy,
          const double* __restrict__ v_gz, int npot, size_t pot_stride, double* __restrict__ phi_buf, double* __restrict__ W,
          int* __restrict__ gflag, int dev_mode) {  // dev_mode (development, SXC_FG_MODE): 1 = the formers skip the row loop (timing experiment: wrong G)
  using namespace scat3;
  extern __shared__ unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(
      smem_raw + ((1024u - (static_cast<unsigned>(__cvta_generic_to_shared(smem_raw)) & 1023u)) & 1023u));
  double* hring = stage_base + C::STAGES * C::STAGE_ELEMS;
  uint64_t* full = reinterpret_cast<uint64_t*>(hring + HWARPS * H_WARP_ELEMS);
  uint64_t* empty = full + C::STAGES;
  uint64_t* hfull = empty + C::STAGES;  // [HWARPS][HSTAGES]
  uint64_t* ready = hfull + HWARPS * HSTAGES;
  uint64_t* release = ready + 2;
  double* scratch = reinterpret_cast<double*>(release + 2);  // [8] partial sums of the block test
  int* s_item = reinterpret_cast<int*>(scratch + 8);         // [2] item index of the slot, [2] skip flag
  int* s_skip = s_item + 2;
  int* s_sig = reinterpret_cast<int*>(scratch + 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, WARPS);
    }
    for (int i = 0; i < HWARPS * HSTAGES; ++i) mbar_init(hfull + i, 1);
    for (int k = 0; k < 2; ++k) {
      mbar_init(ready + k, HWARPS * 32);  // every G former arrives after its own stores
      mbar_init(release + k, WARPS + 1);  // the DMMA warps and the operand producer
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmap);
    tma_prefetch_desc(&tmap_rows);
  }
  __syncthreads();

  if (warp >= PRODUCER_WARP) {
    // ------------------------------------------------------------------------------------ warpgroup 3: operand producer
    setmaxnreg_dec<24>();
    if (warp != PRODUCER_WARP || lane != 0) return;
    int stage = 0, pass = 0;
    for (int it = 0;; ++it) {
      const int k = it & 1;
      mbar_wait(ready + k, (it >> 1) & 1);
      const int qi = s_item[k];
      if (qi >= nitems) return;
      const bool skip = s_skip[k] != 0;
      if (!skip) {
        const WorkItem item = items[qi];
        const int q = item.q;
        const int sp = plan.s_pad[q];
        const ScatterRound2* __restrict__ rounds = tpl + tpl_off[sp >> 5] + item.begin;
        const int row0 = (int)(plan.phi_off[q] / BP);
        if (item.nseg > 1) {
          // the block was cut into several work items (small shard): each item's formers made one piece of its G; wait for all
          // of them (they run in CTAs that pulled their items earlier or at the same time and never wait themselves)
          while (ld_acquire_gpu(gflag + q) < item.nseg) __nanosleep(200);
          asm volatile("fence.proxy.async;\n" ::: "memory");
        }
        vmat2_produce_item<C>(&tmap, rounds, item.end - item.begin, row0, row0 + 4 * sp, stage_base, full, empty, stage, pass);
      }
      mbar_arrive(release + k);
    }
  } else if (warp < HWARPS) {
    // ------------------------------------------------------------------------------------ warpgroup 0: G formers
    // (the lowest warp ids of the CTA: when an FP64-pipe slot frees up they compete with DMMA warps that always have a tensor
    // instruction ready; being the oldest warps of the CTA is the only lever the issue scheduler offers)
    setmaxnreg_dec<40>();
    const int ht = tid;  // 0 .. 127: the point of the block this thread owns
    int hs = 0, hpass = 0;            // ring position, carried from item to item
    for (int it = 0;; ++it) {
      const int k = it & 1;
      if (ht == 0) {
        if (it >= 2) mbar_wait(release + k, ((it >> 1) - 1) & 1);
        s_item[k] = atomicAdd(counter, 1);
      }
      named_bar_sync(2, HWARPS * 32);
      const int qi = s_item[k];
      if (qi >= nitems) {
        mbar_arrive(ready + k);
        return;
      }
      const WorkItem item = items[qi];
      const int q = item.q;
      const long first = (long)plan.block_id[q] * g.blocksize;
      const int n = (int)min((long)g.blocksize, g.npts - first);
      // a = w v_rho, b = w g of the operators that pass their own block-average test (:253-268; npot > 1: summed scatter of
      // several operators, each tested as if it were scattered alone)
      double a = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
      bool any_pass = false;
      for (int kp = 0; kp < npot; ++kp) {
        double pa = 0.0, px = 0.0, py = 0.0, pz = 0.0;
        if (ht < n) {
          const size_t o = (size_t)kp * pot_stride + first + ht;
          const double wp = g.w[first + ht];
          pa = wp * v_rho[o];
          if (v_gx) {
            px = wp * v_gx[o];
            py = wp * v_gy[o];
            pz = wp * v_gz[o];
          }
        }
        const double part = warp_sum(fabs(pa) + fabs(px) + fabs(py) + fabs(pz));
        double* sc = scratch + 4 * (kp & 1);  // alternate halves: one barrier per operator is enough
        if (lane == 0) sc[warp] = part;
        named_bar_sync(2, HWARPS * 32);
        const double total = ((sc[0] + sc[1]) + sc[2]) + sc[3];
        if (!(total / (double)n < block_ave_thr)) {
          any_pass = true;
          a += pa;
          bx += px;
          by += py;
          bz += pz;
        }
      }
      const int s = plan.s[q];
      const bool skip = !any_pass || s == 0;
      if (ht == 0) s_skip[k] = skip ? 1 : 0;
      if (!skip) {
        a *= a_scale;  // 1/2: the T + T^T trick of :281
        const int sp = plan.s_pad[q];
        const int row0 = (int)(plan.phi_off[q] / BP);
        const int ncomp = v_gx ? 4 : 1;
        // rows beyond s rounded up to 8 are never multiplied by the DMMA warps (vmat2_consume_item): G is formed for the others
        const int nst_all = ((dev_mode & 4) ? sp : ((s + 7) & ~7)) / HROWS;
        const int st0 = (int)((long)nst_all * item.seg / item.nseg);  // this item's piece of the rows (all of them: seg 0 of 1)
        const int nst = (dev_mode & 1) ? 0 : (int)((long)nst_all * (item.seg + 1) / item.nseg) - st0;
        double* __restrict__ gout = phi_buf + plan.phi_off[q] + (size_t)4 * sp * BP + (size_t)st0 * HROWS * BP + ht;
        const int hw = warp;  // this warp's quarter of the points and its private ring
        double* wring = hring + hw * H_WARP_ELEMS;
        uint64_t* wfull = hfull + hw * HSTAGES;
        auto issue = [&](int i, int slot) {  // rows [i * HROWS, (i + 1) * HROWS) x 32 points of every component into `slot`
          mbar_arrive_expect_tx(wfull + slot, (unsigned)(ncomp * HROWS * HPTS * sizeof(double)));
          for (int c = 0; c < ncomp; ++c)
            tma_load_2d(wring + slot * H_STAGE_ELEMS + c * HROWS * HPTS, &tmap_rows, hw * HPTS, row0 + c * sp + (st0 + i) * HROWS,
                        wfull + slot);
        };
        // Each warp keeps 3 stages (6 KB) in flight: at HBM latency that is ~3 GB/s per warp, too little to stay ahead of the
        // DMMA warps.  One lane of the CTA therefore pulls the rows 32..47 rows ahead into L2 (one bulk prefetch of 16 rows per
        // component every 8 stages), so that the rings run at L2 latency.
        const double* __restrict__ tile_rows = phi_buf + plan.phi_off[q];
        auto prefetch_rows = [&](int r0) {  // rows [r0, r0 + 16) of this item's piece, every component
          const int r1 = min(r0 + 16, nst * HROWS);
          if (r1 > r0)
            for (int c = 0; c < ncomp; ++c)
              bulk_prefetch_l2(tile_rows + ((size_t)c * sp + (size_t)st0 * HROWS + r0) * BP, (unsigned)((r1 - r0) * BP * sizeof(double)));
        };
        const bool pf = ht == 0 && !(dev_mode & 2);
        if (pf) {
          prefetch_rows(HSTAGES * HROWS);
          prefetch_rows(HSTAGES * HROWS + 16);
        }
        if (lane == 0)
          for (int i = 0; i < min(HSTAGES, nst); ++i) issue(i, (hs + i) % HSTAGES);
        for (int i = 0; i < nst; ++i) {
          if (pf && (i & 7) == 0) prefetch_rows(i * HROWS + HSTAGES * HROWS + 32);
          mbar_wait(wfull + hs, hpass & 1);
          const double* st = wring + hs * H_STAGE_ELEMS + lane;
#pragma unroll
          for (int r = 0; r < HROWS; ++r) {
            double v = a * st[r * HPTS];
            if (ncomp == 4) v += bx * st[(HROWS + r) * HPTS] + by * st[(2 * HROWS + r) * HPTS] + bz * st[(3 * HROWS + r) * HPTS];
            gout[(size_t)(i * HROWS + r) * BP] = v;
          }
          __syncwarp();  // every lane has read the slot
          if (lane == 0 && i + HSTAGES < nst) issue(i + HSTAGES, hs);
          if (++hs == HSTAGES) {
            hs = 0;
            ++hpass;
          }
        }
        fence_global_to_async_proxy();
        if (item.nseg > 1) {  // publish this piece (every former's stores are fenced above; the barrier orders them before the count)
          named_bar_sync(2, HWARPS * 32);
          if (ht == 0) {
            __threadfence();
            atomicAdd(gflag + q, 1);
          }
        }
      }
      mbar_arrive(ready + k);
    }
  } else {
    // ------------------------------------------------------------------------------------ warpgroups 1, 2: DMMA warps
    setmaxnreg_inc<96>();
    const int dwarp = warp - HWARPS, dtid = tid - HWARPS * 32;  // 0 .. 7, 0 .. 255
    int stage = 0, pass = 0;
    for (int it = 0;; ++it) {
      const int k = it & 1;
      mbar_wait(ready + k, (it >> 1) & 1);
      const int qi = s_item[k];
      if (qi >= nitems) return;
      const bool skip = s_skip[k] != 0;
      if (!skip) {
        const WorkItem item = items[qi];
        const int q = item.q;
        const int s = plan.s[q];
        const int sp = plan.s_pad[q];
        const ScatterRound2* __restrict__ rounds = tpl + tpl_off[sp >> 5] + item.begin;
        named_bar_sync(1, WARPS * 32);  // the epilogues of the previous item have read s_sig
        const int* __restrict__ sig_g = plan.sig_bf + (size_t)q * plan.nbf_pad;
        for (int c = dtid; c < sp; c += WARPS * 32) s_sig[c] = sig_g[c];
        named_bar_sync(1, WARPS * 32);
        vmat2_consume_item<C>(rounds, item.end - item.begin, s, sp, nbf, s_sig, stage_base, full, empty, stage, pass, dwarp, lane, W);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(release + k);
    }
  }
}

}  // namespace sxc
