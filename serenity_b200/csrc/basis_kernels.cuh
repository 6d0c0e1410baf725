// basis_kernels.cuh - block prescreening and phi / grad phi evaluation on 128-point tiles.
//
// Row 8a-1 of SURVEY.md: BasisFunctionOnGridController::calculateBasisFunctionData
// (src/data/grid/BasisFunctionOnGridController.cpp:150-1105), derivative level 1.
// B200 design: one CTA per block; one thread per grid point (x/y/z/w coalesced SoA loads), warps of a CTA
// walk the block's significant-shell list, whose shell data (centre, l, primitives) is staged in shared memory in
// batches of 128 shells (warp-uniform -> broadcast reads, no divergence in the switch over l); outputs are written function-major [comp][c][128] so every store is a 256-byte coalesced
// row segment and the tiles feed the DMMA kernels without a transpose.
#pragma once

#include "harmonics_gen.cuh"
#include "sxc_common.cuh"

namespace sxc {

// ------------------------------------------------------------------------------------------------------------
// K0: block prescreening (BasisFunctionOnGridController.cpp:211-255) -> significant shell / function lists.
// grid = owned blocks, 128 threads.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_screen(GridView g, ShellView b, PlanView plan) {
  __shared__ double scratch[32];
  __shared__ int wcount[4], wfunc[4];
  __shared__ int base_count, base_func;
  const int q = blockIdx.x;
  const int blk = plan.block_id[q];
  const long first = (long)blk * g.blocksize;
  const int n = (int)min((long)g.blocksize, g.npts - first);
  const int t = threadIdx.x;

  // centre = mean of the block's points, spread = max distance to it (:211-215)
  double px = 0.0, py = 0.0, pz = 0.0;
  if (t < n) {
    px = g.x[first + t];
    py = g.y[first + t];
    pz = g.z[first + t];
  }
  const double cx = block_sum(t < n ? px : 0.0, scratch) / (double)n;
  const double cy = block_sum(t < n ? py : 0.0, scratch) / (double)n;
  const double cz = block_sum(t < n ? pz : 0.0, scratch) / (double)n;
  double d = 0.0;
  if (t < n) d = sqrt((px - cx) * (px - cx) + (py - cy) * (py - cy) + (pz - cz) * (pz - cz));
  const double spread = block_max(d, scratch);

  if (t == 0) {
    base_count = 0;
    base_func = 0;
  }
  __syncthreads();
  int* sig_shell = plan.sig_shell + (size_t)q * b.nshell;
  int* sig_c0 = plan.sig_c0 + (size_t)q * b.nshell;
  int* sig_bf = plan.sig_bf + (size_t)q * plan.nbf_pad;
  const int lane = t & 31, wid = t >> 5;

  for (int base = 0; base < b.nshell; base += 128) {
    const int sh = base + t;
    int sig = 0, nf = 0;
    if (sh < b.nshell) {
      sig = 1;
      nf = b.nfunc[sh];
      const double dx = cx - b.centre[3 * sh], dy = cy - b.centre[3 * sh + 1], dz = cz - b.centre[3 * sh + 2];
      double dist = sqrt(dx * dx + dy * dy + dz * dz) - spread;
      if (!(dist < 1.0)) {  // :235
        dist = dist * dist;
        double radial = 0.0;
        const int o = b.prim_off[sh];
        for (int i = 0; i < b.nprim[sh]; ++i) radial += b.coeff[o + i] * exp(-(b.alpha[o + i] * dist));
        if (fabs(radial) < b.radial_thr) sig = 0;  // :248
      }
    }
    // ordered compaction: exclusive prefix of (sig, sig*nf) over the 128 threads
    const unsigned ball = __ballot_sync(0xffffffffu, sig);
    int incl_f = sig ? nf : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl_f, o);
      if (lane >= o) incl_f += v;
    }
    if (lane == 31) {
      wcount[wid] = __popc(ball);
      wfunc[wid] = incl_f;
    }
    __syncthreads();
    int pre_c = base_count, pre_f = base_func;
    for (int i = 0; i < wid; ++i) {
      pre_c += wcount[i];
      pre_f += wfunc[i];
    }
    if (sig) {
      const int k = pre_c + __popc(ball & ((1u << lane) - 1u));
      const int c0 = pre_f + incl_f - nf;
      sig_shell[k] = sh;
      sig_c0[k] = c0;
      const int f0 = b.first_bf[sh];
      for (int m = 0; m < nf; ++m) sig_bf[c0 + m] = f0 + m;
    }
    __syncthreads();
    if (t == 0) {
      base_count += wcount[0] + wcount[1] + wcount[2] + wcount[3];
      base_func += wfunc[0] + wfunc[1] + wfunc[2] + wfunc[3];
    }
    __syncthreads();
  }
  if (t == 0) {
    plan.nsig_shell[q] = base_count;
    plan.s[q] = base_func;
  }
  // pad the compact->function map (padding rows gather element 0 of P; their phi rows are zero)
  const int s = base_func;
  const int sp = ((s + SPAD - 1) / SPAD) * SPAD;
  for (int c = s + t; c < sp && c < plan.nbf_pad; c += 128) sig_bf[c] = 0;
}

// ------------------------------------------------------------------------------------------------------------
// K1: phi, d/dx phi, d/dy phi, d/dz phi on the significant shells of each block.
// Tile layout per block: [4][s_pad][128] doubles, component-major, then compact function, then point.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ipow(double x, int k) {
  double r = 1.0;
  for (int i = 0; i < k; ++i) r *= x;
  return r;
}

template <int L>
__device__ __forceinline__ void store_spherical(double* __restrict__ out, size_t comp_stride, double radial,
                                                double dradial, double dx, double dy, double dz) {
  double x[L + 1], y[L + 1], z[L + 1];
  x[0] = y[0] = z[0] = 1.0;
#pragma unroll
  for (int e = 1; e <= L; ++e) {
    x[e] = x[e - 1] * dx;
    y[e] = y[e - 1] * dy;
    z[e] = z[e - 1] * dz;
  }
  constexpr int N = 2 * L + 1;
  double Y[N], Yx[N], Yy[N], Yz[N];
  Harmonics<L>::eval(x, y, z, Y, Yx, Yy, Yz);
  // finalisation, BasisFunctionOnGridController.cpp:1068-1080
#pragma unroll
  for (int m = 0; m < N; ++m) {
    out[(size_t)m * BP] = radial * Y[m];
    out[comp_stride + (size_t)m * BP] = radial * Yx[m] + dradial * dx * Y[m];
    out[2 * comp_stride + (size_t)m * BP] = radial * Yy[m] + dradial * dy * Y[m];
    out[3 * comp_stride + (size_t)m * BP] = radial * Yz[m] + dradial * dz * Y[m];
  }
}

// rare path (l >= 4): table driven
__device__ __noinline__ void store_spherical_generic(int l, double* __restrict__ out, size_t comp_stride,
                                                     double radial, double dradial, double dx, double dy,
                                                     double dz) {
  double x[LMAX + 1], y[LMAX + 1], z[LMAX + 1];
  x[0] = y[0] = z[0] = 1.0;
  for (int e = 1; e <= l; ++e) {
    x[e] = x[e - 1] * dx;
    y[e] = y[e - 1] * dy;
    z[e] = z[e - 1] * dz;
  }
  for (int m = 0; m < 2 * l + 1; ++m) {
    double Y = 0.0, Yx = 0.0, Yy = 0.0, Yz = 0.0;
    for (int t = c_harm_off[l][m]; t < c_harm_off[l][m + 1]; ++t) {
      const double c = c_harm_coef[t];
      const int a = c_harm_ex[t], bb = c_harm_ey[t], cc = c_harm_ez[t];
      Y += c * x[a] * y[bb] * z[cc];
      if (a > 0) Yx += c * a * x[a - 1] * y[bb] * z[cc];
      if (bb > 0) Yy += c * bb * x[a] * y[bb - 1] * z[cc];
      if (cc > 0) Yz += c * cc * x[a] * y[bb] * z[cc - 1];
    }
    out[(size_t)m * BP] = radial * Y;
    out[comp_stride + (size_t)m * BP] = radial * Yx + dradial * dx * Y;
    out[2 * comp_stride + (size_t)m * BP] = radial * Yy + dradial * dy * Y;
    out[3 * comp_stride + (size_t)m * BP] = radial * Yz + dradial * dz * Y;
  }
}

// Cartesian shells (BasisFunctionOnGridController.cpp:359-381), order a = l..0, b = l-a..0
__device__ __noinline__ void store_cartesian(int l, const double* __restrict__ normfac, double* __restrict__ out,
                                             size_t comp_stride, double radial, double dradial, double dx,
                                             double dy, double dz) {
  int m = 0;
  for (int a = l; a >= 0; --a) {
    for (int bb = l - a; bb >= 0; --bb, ++m) {
      const int c = l - a - bb;
      const double nrm = normfac[m];
      const double xa = ipow(dx, a), yb = ipow(dy, bb), zc = ipow(dz, c);
      double vx = dradial * xa * dx * yb * zc * nrm;
      if (a > 0) vx += a * ipow(dx, a - 1) * yb * zc * radial * nrm;
      double vy = dradial * xa * yb * dy * zc * nrm;
      if (bb > 0) vy += bb * xa * ipow(dy, bb - 1) * zc * radial * nrm;
      double vz = dradial * xa * yb * zc * dz * nrm;
      if (c > 0) vz += c * xa * yb * ipow(dz, c - 1) * radial * nrm;
      out[(size_t)m * BP] = xa * yb * zc * radial * nrm;
      out[comp_stride + (size_t)m * BP] = vx;
      out[2 * comp_stride + (size_t)m * BP] = vy;
      out[3 * comp_stride + (size_t)m * BP] = vz;
    }
  }
}

constexpr int BASIS_GROUPS = 4;  // 4 shell groups x 128 points = 512 threads
constexpr int SHELL_BATCH = 128;  // significant shells staged in shared memory at a time
constexpr int PRIM_BATCH = 1024;  // primitives (exponent, coefficient) staged with them

// one significant shell of the block, staged in shared memory
struct __align__(8) StagedShell {
  double cx, cy, cz;
  int c0;      // first compact function index (row of the tile)
  int l;
  int nf;
  int pure;
  int np;
  int poff;    // offset into the staged primitives, or -(global offset) - 1 if they did not fit
  int bf0;     // first basis function (Cartesian norm factors)
  int pad;
};

template <int MINB>  // resident CTAs per SM the register allocation aims at (1: 128 registers; 2: 64)
__global__ void __launch_bounds__(BASIS_GROUPS* BP, MINB) k_basis(GridView g, ShellView b, PlanView plan, int slot0,
                                                             const int* __restrict__ order,
                                                             double* __restrict__ phi_buf) {
  __shared__ StagedShell sh_rec[SHELL_BATCH];
  __shared__ double sh_alpha[PRIM_BATCH], sh_coeff[PRIM_BATCH];
  __shared__ int sh_scan[SHELL_BATCH / 32];
  const int q = order ? order[blockIdx.x] : slot0 + blockIdx.x;
  const int blk = plan.block_id[q];
  const long first = (long)blk * g.blocksize;
  const int n = (int)min((long)g.blocksize, g.npts - first);
  const int tid = threadIdx.x;
  const int p = tid & (BP - 1);
  const int grp = tid >> 7;
  const int sp = plan.s_pad[q];
  const int s = plan.s[q];
  const size_t comp_stride = (size_t)sp * BP;
  double* __restrict__ tile = phi_buf + plan.phi_off[q];
  const bool valid = p < n;
  double px = 0.0, py = 0.0, pz = 0.0;
  if (valid) {
    px = g.x[first + p];
    py = g.y[first + p];
    pz = g.z[first + p];
  }
  const int nsig = plan.nsig_shell[q];
  const int* __restrict__ sig_shell = plan.sig_shell + (size_t)q * b.nshell;
  const int* __restrict__ sig_c0 = plan.sig_c0 + (size_t)q * b.nshell;

  for (int base = 0; base < nsig; base += SHELL_BATCH) {
    const int nb = min(SHELL_BATCH, nsig - base);
    // ---- stage the shell data of this batch (centres, angular momenta, primitives) in shared memory
    __syncthreads();  // the previous batch is no longer read
    if (tid < SHELL_BATCH) {
      int np = 0, sh = 0;
      if (tid < nb) {
        sh = sig_shell[base + tid];
        np = b.nprim[sh];
      }
      int incl = np;  // inclusive prefix sum of the primitive counts over the batch
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
      }
      if ((tid & 31) == 31) sh_scan[tid >> 5] = incl;
      asm volatile("bar.sync 1, %0;" ::"n"(SHELL_BATCH));  // the first four warps only
      int off = incl - np;
      for (int w = 0; w < (tid >> 5); ++w) off += sh_scan[w];
      if (tid < nb) {
        StagedShell r;
        r.cx = b.centre[3 * sh];
        r.cy = b.centre[3 * sh + 1];
        r.cz = b.centre[3 * sh + 2];
        r.c0 = sig_c0[base + tid];
        r.l = b.l[sh];
        r.nf = b.nfunc[sh];
        r.pure = b.pure[sh];
        r.np = np;
        r.bf0 = b.first_bf[sh];
        r.pad = 0;
        const int o = b.prim_off[sh];
        if (off + np <= PRIM_BATCH) {
          r.poff = off;
          for (int i = 0; i < np; ++i) {
            sh_alpha[off + i] = b.alpha[o + i];
            sh_coeff[off + i] = b.coeff[o + i];
          }
        } else {
          r.poff = -o - 1;  // rare: read the primitives from global memory
        }
        sh_rec[tid] = r;
      }
    }
    __syncthreads();

    for (int k = grp; k < nb; k += BASIS_GROUPS) {
      const StagedShell& r = sh_rec[k];
      const int l = r.l;
      const int nf = r.nf;
      double* __restrict__ out = tile + (size_t)r.c0 * BP + p;
      const double dx = px - r.cx, dy = py - r.cy, dz = pz - r.cz;
      const double r2 = dx * dx + dy * dy + dz * dz;
      double radial = 0.0, dradial = 0.0;
      const int np = r.np;
      const bool staged = r.poff >= 0;
      const double* __restrict__ al_p = staged ? sh_alpha + r.poff : b.alpha + (-r.poff - 1);
      const double* __restrict__ co_p = staged ? sh_coeff + r.poff : b.coeff + (-r.poff - 1);
      for (int i = 0; i < np; ++i) {
        const double al = al_p[i];
        const double tmp = al * r2;
        if (tmp < b.exp_thr) {  // :300
          const double e = co_p[i] * exp(-tmp);
          radial += e;
          dradial -= 2.0 * al * e;
        }
      }
      if (!valid || fabs(radial) < b.radial_thr) {  // :312-329 (and the padding points of a short block)
        for (int m = 0; m < nf; ++m) {
          out[(size_t)m * BP] = 0.0;
          out[comp_stride + (size_t)m * BP] = 0.0;
          out[2 * comp_stride + (size_t)m * BP] = 0.0;
          out[3 * comp_stride + (size_t)m * BP] = 0.0;
        }
        continue;
      }
      if (r.pure) {
        switch (l) {
          case 0: store_spherical<0>(out, comp_stride, radial, dradial, dx, dy, dz); break;
          case 1: store_spherical<1>(out, comp_stride, radial, dradial, dx, dy, dz); break;
          case 2: store_spherical<2>(out, comp_stride, radial, dradial, dx, dy, dz); break;
          case 3: store_spherical<3>(out, comp_stride, radial, dradial, dx, dy, dz); break;
          default: store_spherical_generic(l, out, comp_stride, radial, dradial, dx, dy, dz); break;
        }
      } else {
        store_cartesian(l, b.normfac + r.bf0, out, comp_stride, radial, dradial, dx, dy, dz);
      }
    }
  }
  // zero the padding rows c in [s, s_pad)
  for (int c = s + grp; c < sp; c += BASIS_GROUPS) {
#pragma unroll
    for (int comp = 0; comp < 4; ++comp) tile[comp * comp_stride + (size_t)c * BP + p] = 0.0;
  }
}

}  // namespace sxc
