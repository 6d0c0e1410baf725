// scatter_kernel.cuh - v_xc on the grid -> Fock-matrix contribution: V_s = phi_s^T G + G^T phi_s on DMMA tiles.
//
// Row 8a-5 of SURVEY.md: ScalarOperatorToMatrixAdder::addBlock (src/data/grid/ScalarOperatorToMatrixAdder.cpp:225-303):
//     a = w o v_rho,  b_c = w o g_c;   skip the block if (sum|a| + sum|b_x| + sum|b_y| + sum|b_z|)/n < blockAveThreshold;
//     G = diag(b_x) d_x phi_s + diag(b_y) d_y phi_s + diag(b_z) d_z phi_s + 1/2 diag(a) phi_s;
//     T = phi_s^T G;  V_s = T + T^T;  V += Proj V_s Proj^T.
// The LDA variant (:179-223, V_s = phi_s^T diag(a) phi_s) is the same expression with b = 0.
// B200 design: k_form_g builds G in place of the d_x phi tile (bandwidth-bound, fused with the block-average
// test); k_scatter computes only the upper triangle of V_s as stacked-K DMMA tiles
//     U[I,J] = [phi_I | G_I] . [G_J | phi_J]^T      (K = 2 x 128 points)
// with 128 x 128 output tiles (16 warps, 4 x 4 warp tiles of 32 x 32), a 4-stage cp.async ring over 8-point K chunks,
// and accumulates into the GPU-resident
// upper triangle with FP64 red.global (RED.E.ADD.F64); k_mirror copies the strict upper triangle down once per build.
#pragma once

#include "sxc_common.cuh"

namespace sxc {

// ------------------------------------------------------------------------------------------------------------
// K3b: per block: weights x potential, block-average test, G in place of d_x phi.  256 threads.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_form_g(GridView g, PlanView plan, const int* __restrict__ order, double block_ave_thr,
         const double* __restrict__ v_rho, const double* __restrict__ v_gx, const double* __restrict__ v_gy,
         const double* __restrict__ v_gz, double* __restrict__ phi_buf, int* __restrict__ skip_flag) {
  __shared__ double sa[BP], sx[BP], sy[BP], sz[BP];
  __shared__ double scratch[32];
  const int q = order[blockIdx.x];
  const int blk = plan.block_id[q];
  const long first = (long)blk * g.blocksize;
  const int n = (int)min((long)g.blocksize, g.npts - first);
  const int tid = threadIdx.x;
  double mag = 0.0;
  if (tid < BP) {
    double a = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
    if (tid < n) {
      const double wp = g.w[first + tid];
      a = wp * v_rho[first + tid];
      if (v_gx) {
        bx = wp * v_gx[first + tid];
        by = wp * v_gy[first + tid];
        bz = wp * v_gz[first + tid];
      }
    }
    sa[tid] = a;
    sx[tid] = bx;
    sy[tid] = by;
    sz[tid] = bz;
    mag = fabs(a) + fabs(bx) + fabs(by) + fabs(bz);
  }
  const double total = block_sum(mag, scratch);
  const int s = plan.s[q];
  const bool skip = (total / (double)n < block_ave_thr) || s == 0;  // :262-268
  if (tid == 0) skip_flag[q] = skip ? 1 : 0;
  if (skip) return;
  const int sp = plan.s_pad[q];
  const size_t comp_stride = (size_t)sp * BP;
  double* __restrict__ tile = phi_buf + plan.phi_off[q];
  const int p = tid & (BP - 1);
  const double a = 0.5 * sa[p], bx = sx[p], by = sy[p], bz = sz[p];
  for (int c = tid >> 7; c < sp; c += 2) {
    const size_t i = (size_t)c * BP + p;
    tile[comp_stride + i] = bx * tile[comp_stride + i] + by * tile[2 * comp_stride + i] +
                            bz * tile[3 * comp_stride + i] + a * tile[i];  // :276-281
  }
}

// ------------------------------------------------------------------------------------------------------------
// K4: U = phi_I G_J^T + G_I phi_J^T on 128 x 128 tiles, upper triangle only, atomically accumulated into W.
// Work item = (slot q, row tile I); the CTA loops over the column tiles J >= I.
// ------------------------------------------------------------------------------------------------------------
namespace scat {
constexpr int THREADS = 512;
constexpr int TI = 128, TJ = 128;
constexpr int TKP = 8;               // points per K chunk
constexpr int STRIDE = TKP + 4;      // 12 doubles: conflict-free fragment loads
constexpr int ROWS = 2 * TI + 2 * TJ;  // phi_I, G_I, G_J, phi_J
constexpr int STAGE_ELEMS = ROWS * STRIDE;
constexpr int STAGES = 4;
constexpr size_t smem_bytes(int s_pad_max) {
  return (size_t)STAGES * STAGE_ELEMS * sizeof(double) + (size_t)(s_pad_max + TI) * sizeof(int);
}
}  // namespace scat

struct ScatterItem {
  int q;   // plan slot
  int it;  // row tile index (128 rows)
};

__global__ void __launch_bounds__(scat::THREADS, 1)
k_scatter(PlanView plan, int nbf, const ScatterItem* __restrict__ items, const int* __restrict__ skip_flag,
          const double* __restrict__ phi_buf, double* __restrict__ W) {
  using namespace scat;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  int* sig = reinterpret_cast<int*>(stage_base + STAGES * STAGE_ELEMS);

  const ScatterItem item = items[blockIdx.x];
  const int q = item.q;
  if (skip_flag[q]) return;
  const int s = plan.s[q];
  const int sp = plan.s_pad[q];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* __restrict__ phi = phi_buf + plan.phi_off[q];
  const double* __restrict__ G = phi + (size_t)sp * BP;  // G lives in the d_x phi slot
  const int* __restrict__ sig_g = plan.sig_bf + (size_t)q * plan.nbf_pad;
  for (int c = tid; c < sp; c += THREADS) sig[c] = sig_g[c];

  const int i0 = item.it * TI;
  const int nj = sp / TJ + ((sp % TJ) ? 1 : 0);
  const int jt0 = (i0 / TJ);  // first column tile that reaches the diagonal
  const int npairs = nj - jt0;
  constexpr int NKC = BP / TKP;  // 16 chunks per tile pair
  const int total = npairs * NKC;
  const int pw = warp & 3, jw = warp >> 2;
  const int lr = lane >> 2, lc = lane & 3;

  auto issue = [&](int gi) {
    if (gi < total) {
      const int jp = gi / NKC, kc = gi - jp * NKC;
      const int j0 = (jt0 + jp) * TJ;
      double* st = stage_base + (gi % STAGES) * STAGE_ELEMS;
      // 512 rows x 64 B = 2048 x 16 B
#pragma unroll
      for (int i = 0; i < (ROWS * 4) / THREADS; ++i) {
        const int idx = tid + i * THREADS;
        const int row = idx >> 2, c16 = idx & 3;
        int r;
        const double* src;
        if (row < TI) {
          r = i0 + row;
          src = phi;
        } else if (row < 2 * TI) {
          r = i0 + row - TI;
          src = G;
        } else if (row < 2 * TI + TJ) {
          r = j0 + row - 2 * TI;
          src = G;
        } else {
          r = j0 + row - 2 * TI - TJ;
          src = phi;
        }
        r = min(r, sp - 1);  // rows past s_pad are masked at the atomics stage
        cp_async16(st + row * STRIDE + c16 * 2, src + (size_t)r * BP + kc * TKP + c16 * 2);
      }
    }
    cp_async_commit();
  };

  double acc[4][4][2];
  issue(0);
  issue(1);
  issue(2);
  __syncthreads();  // sig[] visible
  int gi = 0;
  for (int jp = 0; jp < npairs; ++jp) {
    const int j0 = (jt0 + jp) * TJ;
    // warp tiles entirely below the diagonal or outside the matrix contribute nothing
    const int wi0 = i0 + pw * 32, wj0 = j0 + jw * 32;
    const bool active = (wi0 < sp) && (wj0 < sp) && (wj0 + 31 >= wi0);
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) acc[m][nn][0] = acc[m][nn][1] = 0.0;
    for (int kc = 0; kc < NKC; ++kc, ++gi) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      issue(gi + 3);
      if (active) {
        const double* st = stage_base + (gi % STAGES) * STAGE_ELEMS;
        const double* sPhiI = st;
        const double* sGI = st + TI * STRIDE;
        const double* sGJ = st + 2 * TI * STRIDE;
        const double* sPhiJ = sGJ + TJ * STRIDE;
#pragma unroll
        for (int ks = 0; ks < TKP / 4; ++ks) {
          double a1[4], a2[4], b1[4], b2[4];
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int o = (pw * 32 + m * 8 + lr) * STRIDE + ks * 4 + lc;
            a1[m] = sPhiI[o];
            a2[m] = sGI[o];
          }
#pragma unroll
          for (int nn = 0; nn < 4; ++nn) {
            const int o = (jw * 32 + nn * 8 + lr) * STRIDE + ks * 4 + lc;
            b1[nn] = sGJ[o];
            b2[nn] = sPhiJ[o];
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
    if (active) {
      // V += Proj V_s Proj^T (:301), upper triangle (compact i <= j <=> global sig[i] <= sig[j])
#pragma unroll
      for (int nn = 0; nn < 4; ++nn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = wj0 + nn * 8 + 2 * lc + e;
          if (j >= s) continue;
          const size_t col = (size_t)sig[j] * nbf;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int i = wi0 + m * 8 + lr;
            if (i <= j) atomicAdd(W + col + sig[i], acc[m][nn][e]);
          }
        }
    }
  }
  cp_async_wait<0>();
}

// V[j,i] = V[i,j] for i < j (column-major, upper triangle holds the sums)
__global__ void k_mirror(int nbf, double* __restrict__ V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < nbf && j < nbf && i < j) V[j + (size_t)i * nbf] = V[i + (size_t)j * nbf];
}

}  // namespace sxc
