// density_kernel.cuh - rho and grad rho on a block: B = phi_s P_s on the FP64 tensor cores, fused row sums.
//
// Row 8a-2 of SURVEY.md: MatrixOperatorToGridTransformer::transform
// (src/data/grid/MatrixOperatorToGridTransformer.cpp:103-165):
//     phi_s = phi Proj,  P_s = Proj^T P Proj,  B = phi_s P_s,
//     rho_p = sum_mu B_pmu phi_pmu,   grad rho_p = 2 sum_mu B_pmu grad phi_pmu.
// B200 design: one CTA (8 warps = 4 point groups x 2 function groups, warp tile 32 points x 32 functions; two CTAs
// per SM) per 128-point block.  The product runs as DMMA m8n8k4 tiles (mma.sync ... f64, the only FP64 tensor path on
// sm_100a): per j-tile of 64 functions the K loop streams 16-function chunks of the phi tile (cp.async, 16 B) and the
// matching gathered P_s chunk (cp.async, 8 B, straight from the L2-resident nb x nb matrix through the block's
// compact->function map) through a 3-stage shared-memory ring; B never leaves registers - the epilogue multiplies the
// accumulator fragments with phi / grad phi read in fragment layout (full 32-byte sectors) and reduces over functions
// with warp shuffles.  s_pad is a multiple of 32, so the last j-tile may hold a single 32-function group: its K range
// is then split over the two function-group warps (rho is linear in B, so the halves need no extra reduction).
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
constexpr size_t smem_bytes(int s_pad_max) {
  return (size_t)STAGES * STAGE_ELEMS * sizeof(double) + (size_t)NJW * BP * 4 * sizeof(double) +
         (size_t)(s_pad_max + TJ) * sizeof(int);
}
}  // namespace dens

__global__ void __launch_bounds__(dens::THREADS, 2)
k_density(GridView g, PlanView plan, int nbf, const double* __restrict__ P, const WorkItem* __restrict__ items,
          const double* __restrict__ phi_buf, double* __restrict__ rho, double* __restrict__ gx,
          double* __restrict__ gy, double* __restrict__ gz, int* __restrict__ nonneg) {
  using namespace dens;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  double* red = stage_base + STAGES * STAGE_ELEMS;          // [NJW][128][4]
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
  const double* __restrict__ tile = phi_buf + plan.phi_off[q];
  const size_t comp_stride = (size_t)sp * BP;
  const int* __restrict__ sig_g = plan.sig_bf + (size_t)q * plan.nbf_pad;
  for (int c = tid; c < sp + TJ; c += THREADS) sig[c] = c < sp ? sig_g[c] : 0;
  __syncthreads();

  const int nk = sp / TK;                 // K chunks per j-tile
  const int n32 = sp / 32;                // 32-function column groups
  const int njt_all = (n32 + NJW - 1) / NJW;  // j-tiles of 64 of the block
  const int jt_begin = item.begin, njt = item.end;  // this CTA's segment of them (normally all)
  const bool partial = jt_begin != 0 || njt != njt_all;
  const int pw = warp & 3, jw = warp >> 2;
  const int lr = lane >> 2, lc = lane & 3;

  // Producer side of the ring.  Every thread copies 4 x 16 B of the phi chunk and gathers 4 elements of the P_s chunk per
  // stage; all per-thread address parts are loop invariants or advance by constants (no divisions in the loop).
  const double* a_src = tile + (size_t)(tid >> 6) * BP + (tid & 63) * 2;  // + kc * TK * BP
  const int a_dst = (tid >> 6) * A_STRIDE + (tid & 63) * 2;
  const int bk = tid & (TK - 1), bj = tid >> 4;                           // gather: k = bk, j = bj + 16 i
  const int b_dst = A_ELEMS + bj * B_STRIDE + bk;
  int is_jt = jt_begin, is_kc = 0, is_stage = 0;
  int colbase[4] = {0, 0, 0, 0};  // sig[j] * nbf of this thread's four columns of the j-tile being issued
  auto load_colbase = [&]() {
    if (is_jt < njt) {  // (sig holds s_pad + TJ entries: nothing to read past the last j-tile)
#pragma unroll
      for (int i = 0; i < 4; ++i) colbase[i] = sig[is_jt * TJ + bj + 16 * i] * nbf;
    }
  };
  load_colbase();
  auto issue = [&]() {
    if (is_jt < njt) {
      double* st = stage_base + is_stage * STAGE_ELEMS;
      const double* src = a_src + (size_t)is_kc * (TK * BP);
#pragma unroll
      for (int i = 0; i < 4; ++i) cp_async16(st + a_dst + i * 4 * A_STRIDE, src + i * 4 * BP);
      const int ncol = sp - is_jt * TJ;  // columns of this j-tile that exist (>= 64 except for the last tile)
      const double* prow = P + sig[is_kc * TK + bk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (bj + 16 * i < ncol) cp_async8(st + b_dst + i * 16 * B_STRIDE, prow + colbase[i]);
      if (++is_kc == nk) {
        is_kc = 0;
        ++is_jt;
        load_colbase();
      }
      is_stage = (is_stage + 1 == STAGES) ? 0 : is_stage + 1;
    }
    cp_async_commit();
  };

  double acc[4][4][2];
  for (int i = tid; i < NJW * BP * 4; i += THREADS) red[i] = 0.0;

  issue();
  issue();
  int c_stage = 0;
  for (int jt = jt_begin; jt < njt; ++jt) {
    // a last j-tile with one 32-function group: both function-group warps work on it, on alternating k-steps
    const bool split = (n32 - jt * NJW) == 1;
    const int cg = split ? 0 : jw;
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) acc[m][nn][0] = acc[m][nn][1] = 0.0;
    for (int kc = 0; kc < nk; ++kc) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      issue();
      const double* As = stage_base + c_stage * STAGE_ELEMS;
      const double* Bs = As + A_ELEMS;
      c_stage = (c_stage + 1 == STAGES) ? 0 : c_stage + 1;
#pragma unroll
      for (int ks = 0; ks < TK / 4; ++ks) {
        if (split && (ks & 1) != jw) continue;
        double a[4], bfrag[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) a[m] = As[(ks * 4 + lc) * A_STRIDE + pw * 32 + m * 8 + lr];
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) bfrag[nn] = Bs[(cg * 32 + nn * 8 + lr) * B_STRIDE + ks * 4 + lc];
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int nn = 0; nn < 4; ++nn) dmma884(acc[m][nn][0], acc[m][nn][1], a[m], bfrag[nn]);
      }
    }
    {
      // epilogue: rho += B o phi, grad rho += B o grad phi   (MatrixOperatorToGridTransformer.cpp:158-163)
      const int jbase = jt * TJ + cg * 32 + 2 * lc;
      double r_rho[4], r_x[4], r_y[4], r_z[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) r_rho[m] = r_x[m] = r_y[m] = r_z[m] = 0.0;
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const size_t row = (size_t)(jbase + nn * 8 + e) * BP;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int p = pw * 32 + m * 8 + lr;
            const double c = acc[m][nn][e];
            r_rho[m] += c * tile[row + p];
            if (gx) {
              r_x[m] += c * tile[comp_stride + row + p];
              r_y[m] += c * tile[2 * comp_stride + row + p];
              r_z[m] += c * tile[3 * comp_stride + row + p];
            }
          }
        }
      }
      // reduce over the 4 lanes of a fragment row; lane lc == 0 owns the (jw, point) slot of the CTA scratch
#pragma unroll
      for (int m = 0; m < 4; ++m) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          r_rho[m] += __shfl_xor_sync(0xffffffffu, r_rho[m], o);
          r_x[m] += __shfl_xor_sync(0xffffffffu, r_x[m], o);
          r_y[m] += __shfl_xor_sync(0xffffffffu, r_y[m], o);
          r_z[m] += __shfl_xor_sync(0xffffffffu, r_z[m], o);
        }
        if (lc == 0) {
          double* dst = red + ((size_t)jw * BP + pw * 32 + m * 8 + lr) * 4;
          dst[0] += r_rho[m];
          dst[1] += r_x[m];
          dst[2] += r_y[m];
          dst[3] += r_z[m];
        }
      }
    }
  }
  cp_async_wait<0>();
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
