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
// the HBM pipe and the DMMA pipe busy together.  The two sides are decoupled: the formers of all CTAs work through the block
// queue from their own counter and publish a per-block flag in global memory (release), the operand producer of the CTA that
// contracts a block waits for its flag (acquire) - the formers never wait for anybody, so they run as far ahead as their
// throughput allows.  (FP64 FMAs and DMMA share one pipe, so the formers' arithmetic is not free: DESIGN.md section 3.)
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
          int nbf, const WorkItem* __restrict__ items, int nitems, const int* __restrict__ fblocks, int nfblocks,
          int* __restrict__ counter, int* __restrict__ fcounter, const ScatterRound2* __restrict__ tpl,
          const int* __restrict__ tpl_off, int sig_cap,
          double block_ave_thr, double a_scale, const double* __restrict__ v_rho, const double* __restrict__ v_gx,
          const double* __restrict__ v_gy, const double* __restrict__ v_gz, int npot, size_t pot_stride,
          double* __restrict__ phi_buf, double* __restrict__ W, int* __restrict__ gflag, int dev_mode) {
  // items: blocks, or segments of their rounds when a shard is small, in queue order; fblocks: the blocks themselves in the same
  // order.  The formers of all CTAs form fblocks[0 .. nfblocks) from their own counter and publish gflag[q] = 1 (formed) or 2 (the
  // block-average test failed); the contraction side pulls items from `counter` and waits for the flag of its block.
  // dev_mode (development, SXC_FG_MODE): 1 = the formers skip the row loop (timing experiment: wrong G)
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
  int* s_item = reinterpret_cast<int*>(scratch + 8);         // [2] item index of the slot, [2] skip flag, [1] the formers' item
  int* s_skip = s_item + 2;
  int* s_fitem = s_item + 4;
  int* s_sig = reinterpret_cast<int*>(scratch + 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, WARPS);
    }
    for (int i = 0; i < HWARPS * HSTAGES; ++i) mbar_init(hfull + i, 1);
    for (int k = 0; k < 2; ++k) {
      mbar_init(ready + k, 1);        // the operand producer publishes an item slot
      mbar_init(release + k, WARPS);  // the DMMA warps are done with it
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
      if (it >= 2) mbar_wait(release + k, ((it >> 1) - 1) & 1);  // the DMMA warps have left the slot's previous item
      const int qi = atomicAdd(counter, 1);
      s_item[k] = qi;
      if (qi >= nitems) {
        mbar_arrive(ready + k);
        return;
      }
      const WorkItem item = items[qi];
      const int q = item.q;
      int f;
      while ((f = ld_acquire_gpu(gflag + q)) == 0) __nanosleep(100);
      const int skip = f == 2;
      asm volatile("fence.proxy.async;\n" ::: "memory");  // the formers' (generic-proxy) stores before this thread's TMA loads
      s_skip[k] = skip;
      mbar_arrive(ready + k);
      if (!skip) {
        const int sp = plan.s_pad[q];
        const ScatterRound2* __restrict__ rounds = tpl + tpl_off[sp >> 5] + item.begin;
        const int row0 = (int)(plan.phi_off[q] / BP);
        vmat2_produce_item<C>(&tmap, rounds, item.end - item.begin, row0, row0 + 4 * sp, stage_base, full, empty, stage, pass);
      }
    }
  } else if (warp < HWARPS) {
    // ------------------------------------------------------------------------------------ warpgroup 0: G formers
    // (the lowest warp ids of the CTA: when an FP64-pipe slot frees up they compete with DMMA warps that always have a tensor
    // instruction ready; being the oldest warps of the CTA is the only lever the issue scheduler offers)
    setmaxnreg_dec<40>();
    const int ht = tid;  // 0 .. 127: the point of the block this thread owns
    int hs = 0, hpass = 0;  // ring position, carried from item to item
    for (;;) {
      if (ht == 0) *s_fitem = atomicAdd(fcounter, 1);
      named_bar_sync(2, HWARPS * 32);
      const int fi = *s_fitem;
      if (fi >= nfblocks) return;
      const int q = fblocks[fi];
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
      if (!skip) {
        a *= a_scale;  // 1/2: the T + T^T trick of :281
        const int sp = plan.s_pad[q];
        const int row0 = (int)(plan.phi_off[q] / BP);
        const int ncomp = v_gx ? 4 : 1;
        // rows beyond s rounded up to 8 are never multiplied by the DMMA warps (vmat2_consume_item): G is formed for the others
        const int nst = (dev_mode & 1) ? 0 : ((s + 7) & ~7) / HROWS;
        double* __restrict__ gout = phi_buf + plan.phi_off[q] + (size_t)4 * sp * BP + ht;
        const int hw = warp;  // this warp's quarter of the points and its private ring
        double* wring = hring + hw * H_WARP_ELEMS;
        uint64_t* wfull = hfull + hw * HSTAGES;
        auto issue = [&](int i, int slot) {  // rows [i * HROWS, (i + 1) * HROWS) x 32 points of every component into `slot`
          mbar_arrive_expect_tx(wfull + slot, (unsigned)(ncomp * HROWS * HPTS * sizeof(double)));
          for (int c = 0; c < ncomp; ++c)
            tma_load_2d(wring + slot * H_STAGE_ELEMS + c * HROWS * HPTS, &tmap_rows, hw * HPTS, row0 + c * sp + i * HROWS, wfull + slot);
        };
        const bool no_loads = (dev_mode & 16) != 0, no_stores = (dev_mode & 32) != 0;  // development: which part interferes
        if (lane == 0 && !no_loads)
          for (int i = 0; i < min(HSTAGES, nst); ++i) issue(i, (hs + i) % HSTAGES);
        for (int i = 0; i < nst; ++i) {
          if (!no_loads) mbar_wait(wfull + hs, hpass & 1);
          const double* st = wring + hs * H_STAGE_ELEMS + lane;
#pragma unroll
          for (int r = 0; r < HROWS; ++r) {
            double v = a * st[r * HPTS];
            if (ncomp == 4) v += bx * st[(HROWS + r) * HPTS] + by * st[(2 * HROWS + r) * HPTS] + bz * st[(3 * HROWS + r) * HPTS];
            if (!no_stores || v == 1.2345678e300) gout[(size_t)(i * HROWS + r) * BP] = v;
          }
          __syncwarp();  // every lane has read the slot
          if (lane == 0 && i + HSTAGES < nst && !no_loads) issue(i + HSTAGES, hs);
          if (++hs == HSTAGES) {
            hs = 0;
            ++hpass;
          }
        }
        __threadfence();  // this thread's G stores are visible device-wide before the flag below
      }
      named_bar_sync(2, HWARPS * 32);  // all formers have stored and fenced (and read s_fitem)
      if (ht == 0) {
        __threadfence();
        atomicExch(gflag + q, skip ? 2 : 1);
      }
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
