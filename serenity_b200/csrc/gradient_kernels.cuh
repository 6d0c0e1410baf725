// gradient_kernels.cuh - XC nuclear gradient: FuncPotential<SCFMode>::getGeomGradients
// (src/potentials/FuncPotential.cpp:114-239).
//
// Reference, per block, over all pairs of significant functions (an O(s^2 n) scalar double loop):
//   g[atom(nu), c] -= 2 P_mu,nu sum_p w [ v phi_mu d_c phi_nu + sum_d vg_d ( phi_mu d_d d_c phi_nu + d_d phi_mu d_c phi_nu ) ]
// B200 formulation - the same sums regrouped into two block GEMMs that reuse the density kernel's machinery:
//   X = phi_s P_s,   M = K P_s   with   K = a phi + b . grad phi   (a = w v, b = w vg; "G" of the scatter with a full a)
//   t[nu, c] = sum_p ( M_p,nu d_c phi_p,nu + X_p,nu q^c_p,nu ),   q^c = sum_d b_d d_d d_c phi   (Hessian contracted on the fly)
//   g[atom(nu), c] = -2 t[nu, c]
// The six Hessian components are never stored: k_hessq evaluates them per (point, function) from the monomial tables
// and writes only the three contractions q^c.  Tile slots of a gradient plan: 0 phi, 1-3 grad phi, 4 K, 5-7 q.
#pragma once

#include "harmonics_gen.cuh"
#include "sxc_common.cuh"

namespace sxc {

constexpr int GRAD_TILE_COMPS = 8;

// ------------------------------------------------------------------------------------------------------------
// q^c = sum_d b_d d_d d_c phi on the significant shells of each block (BasisFunctionOnGridController.cpp:302-304 radial
// second derivative, :1081-1095 / :381-440 finalisation).  For phi = R(r^2) Y(x, y, z):
//   d_d d_c phi = R Y_dc + R1 (x_d Y_c + x_c Y_d + delta_dc Y) + R2 x_d x_c Y,   R1 = -2 sum alpha c e,  R2 = 4 sum alpha^2 c e
//   q^c = R (b . grad) Y_c + R1 [ (b . x) Y_c + x_c (b . grad Y) + b_c Y ] + R2 x_c (b . x) Y
// Table driven for every l (the gradient is evaluated once per geometry step, not per SCF iteration).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double pw_(const double* p, int e) { return e >= 0 ? p[e] : 0.0; }

__global__ void __launch_bounds__(BASIS_GROUPS* BP)
k_hessq(GridView g, ShellView b, PlanView plan, int slot0, const int* __restrict__ order, const double* __restrict__ v_gx,
        const double* __restrict__ v_gy, const double* __restrict__ v_gz, double* __restrict__ phi_buf, int unit_dir = -1) {
  // unit_dir = 0, 1, 2: b is the unit vector of that axis at every point (no weights, v_g* unused), so the three outputs are the
  // second derivatives d_dir d_c phi themselves (sxc_basis_hessian_on_grid)
  const int q = order ? order[blockIdx.x] : slot0 + blockIdx.x;
  const int blk = plan.block_id[q];
  const long first = (long)blk * g.blocksize;
  const int n = (int)min((long)g.blocksize, g.npts - first);
  const int p = threadIdx.x & (BP - 1);
  const int grp = threadIdx.x >> 7;
  const int sp = plan.s_pad[q];
  const int s = plan.s[q];
  const size_t comp_stride = (size_t)sp * BP;
  double* __restrict__ tile = phi_buf + plan.phi_off[q];
  const bool valid = p < n;
  double px = 0.0, py = 0.0, pz = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
  if (valid) {
    px = g.x[first + p];
    py = g.y[first + p];
    pz = g.z[first + p];
    if (unit_dir >= 0) {
      bx = unit_dir == 0 ? 1.0 : 0.0;
      by = unit_dir == 1 ? 1.0 : 0.0;
      bz = unit_dir == 2 ? 1.0 : 0.0;
    } else {
      const double wp = g.w[first + p];
      bx = wp * v_gx[first + p];
      by = wp * v_gy[first + p];
      bz = wp * v_gz[first + p];
    }
  }
  const int nsig = plan.nsig_shell[q];
  const int* __restrict__ sig_shell = plan.sig_shell + (size_t)q * b.nshell;
  const int* __restrict__ sig_c0 = plan.sig_c0 + (size_t)q * b.nshell;

  for (int k = grp; k < nsig; k += BASIS_GROUPS) {
    const int sh = sig_shell[k];
    const int c0 = sig_c0[k];
    const int l = b.l[sh];
    const int nf = b.nfunc[sh];
    const bool pure = b.pure[sh] != 0;
    double* __restrict__ out = tile + 5 * comp_stride + (size_t)c0 * BP + p;
    const double dx = px - b.centre[3 * sh], dy = py - b.centre[3 * sh + 1], dz = pz - b.centre[3 * sh + 2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    double R = 0.0, R1 = 0.0, R2 = 0.0;
    const int o = b.prim_off[sh];
    const int np = b.nprim[sh];
    for (int i = 0; i < np; ++i) {
      const double al = b.alpha[o + i];
      const double tmp = al * r2;
      if (tmp < b.exp_thr) {  // :300
        const double e = b.coeff[o + i] * exp(-tmp);
        R += e;
        R1 -= 2.0 * al * e;
        R2 += 4.0 * al * al * e;
      }
    }
    const bool zero = !valid || fabs(R) < b.radial_thr;  // :312-329
    double x[LMAX + 1], y[LMAX + 1], z[LMAX + 1];
    x[0] = y[0] = z[0] = 1.0;
    for (int e = 1; e <= l; ++e) {
      x[e] = x[e - 1] * dx;
      y[e] = y[e - 1] * dy;
      z[e] = z[e - 1] * dz;
    }
    const double bdotx = bx * dx + by * dy + bz * dz;
    int ca = l, cb = 0;  // Cartesian exponents of component m: a = l..0, b = l-a..0 (:360-362)
    for (int m = 0; m < nf; ++m) {
      double Y = 0.0, Yx = 0.0, Yy = 0.0, Yz = 0.0, Yxx = 0.0, Yxy = 0.0, Yxz = 0.0, Yyy = 0.0, Yyz = 0.0, Yzz = 0.0;
      int t0, t1;
      if (pure) {
        t0 = c_harm_off[l][m];
        t1 = c_harm_off[l][m + 1];
      } else {
        t0 = 0;
        t1 = 1;
      }
      for (int t = t0; t < t1; ++t) {
        double c;
        int ea, eb, ec;
        if (pure) {
          c = c_harm_coef[t];
          ea = c_harm_ex[t];
          eb = c_harm_ey[t];
          ec = c_harm_ez[t];
        } else {
          c = b.normfac[b.first_bf[sh] + m];
          ea = ca;
          eb = cb;
          ec = l - ca - cb;
        }
        const double xa = x[ea], yb = y[eb], zc = z[ec];
        const double xa1 = ea * pw_(x, ea - 1), yb1 = eb * pw_(y, eb - 1), zc1 = ec * pw_(z, ec - 1);
        const double xa2 = ea * (ea - 1) * pw_(x, ea - 2), yb2 = eb * (eb - 1) * pw_(y, eb - 2),
                     zc2 = ec * (ec - 1) * pw_(z, ec - 2);
        Y += c * xa * yb * zc;
        Yx += c * xa1 * yb * zc;
        Yy += c * xa * yb1 * zc;
        Yz += c * xa * yb * zc1;
        Yxx += c * xa2 * yb * zc;
        Yyy += c * xa * yb2 * zc;
        Yzz += c * xa * yb * zc2;
        Yxy += c * xa1 * yb1 * zc;
        Yxz += c * xa1 * yb * zc1;
        Yyz += c * xa * yb1 * zc1;
      }
      if (!pure) {  // next Cartesian component
        if (cb == 0) {
          --ca;
          cb = l - ca;
        } else {
          --cb;
        }
      }
      double qx = 0.0, qy = 0.0, qz = 0.0;
      if (!zero) {
        const double bgY = bx * Yx + by * Yy + bz * Yz;
        qx = R * (bx * Yxx + by * Yxy + bz * Yxz) + R1 * (bdotx * Yx + dx * bgY + bx * Y) + R2 * dx * bdotx * Y;
        qy = R * (bx * Yxy + by * Yyy + bz * Yyz) + R1 * (bdotx * Yy + dy * bgY + by * Y) + R2 * dy * bdotx * Y;
        qz = R * (bx * Yxz + by * Yyz + bz * Yzz) + R1 * (bdotx * Yz + dz * bgY + bz * Y) + R2 * dz * bdotx * Y;
      }
      out[(size_t)m * BP] = qx;
      out[comp_stride + (size_t)m * BP] = qy;
      out[2 * comp_stride + (size_t)m * BP] = qz;
    }
  }
  for (int c = s + grp; c < sp; c += BASIS_GROUPS) {  // padding rows
#pragma unroll
    for (int comp = 5; comp < 8; ++comp) tile[comp * comp_stride + (size_t)c * BP + p] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------------------
// t[nu, c] += sum_p (A_s P_s)_p,nu E^c_p,nu  with A = tile slot a_slot, E^c = slots e_slot .. e_slot + 2:
// the product of k_density (same tiling, ring and DMMA loop) with an epilogue that reduces over the POINTS of the block
// and accumulates per basis function (atomics on gfunc[nbf][3]; the host folds functions into atoms).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(dens::THREADS, 2)
k_grad_contract(PlanView plan, int nbf, const double* __restrict__ P, const int* __restrict__ order,
                const double* __restrict__ phi_buf, int a_slot, int e_slot, double* __restrict__ gfunc) {
  using namespace dens;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  int* sig = reinterpret_cast<int*>(stage_base + STAGES * STAGE_ELEMS + NJW * BP * 4);  // same carve-up as k_density

  const int q = order[blockIdx.x];
  const int s = plan.s[q];
  if (s == 0) return;
  const int sp = plan.s_pad[q];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t comp_stride = (size_t)sp * BP;
  const double* __restrict__ tile = phi_buf + plan.phi_off[q];
  const double* __restrict__ Asrc = tile + (size_t)a_slot * comp_stride;
  const double* __restrict__ Esrc = tile + (size_t)e_slot * comp_stride;
  const int* __restrict__ sig_g = plan.sig_bf + (size_t)q * plan.nbf_pad;
  for (int c = tid; c < sp + TJ; c += THREADS) sig[c] = c < sp ? sig_g[c] : 0;
  __syncthreads();

  const int nk = sp / TK;
  const int n32 = sp / 32;
  const int njt = (n32 + NJW - 1) / NJW;
  const int pw = warp & 3, jw = warp >> 2;
  const int lr = lane >> 2, lc = lane & 3;

  const double* a_src = Asrc + (size_t)(tid >> 6) * BP + (tid & 63) * 2;
  const int a_dst = (tid >> 6) * A_STRIDE + (tid & 63) * 2;
  const int bk = tid & (TK - 1), bj = tid >> 4;
  const int b_dst = A_ELEMS + bj * B_STRIDE + bk;
  int is_jt = 0, is_kc = 0, is_stage = 0;
  int colbase[4] = {0, 0, 0, 0};
  auto load_colbase = [&]() {
    if (is_jt < njt) {
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
      const int ncol = sp - is_jt * TJ;
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
  issue();
  issue();
  int c_stage = 0;
  for (int jt = 0; jt < njt; ++jt) {
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
    // epilogue: per function column, sum over the warp's 32 points of acc * E^c, then over the 8 lanes lr
    const int jbase = jt * TJ + cg * 32 + 2 * lc;
#pragma unroll
    for (int nn = 0; nn < 4; ++nn)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = jbase + nn * 8 + e;
        const size_t row = (size_t)j * BP;
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int p = pw * 32 + m * 8 + lr;
          const double c = acc[m][nn][e];
          t0 += c * Esrc[row + p];
          t1 += c * Esrc[comp_stride + row + p];
          t2 += c * Esrc[2 * comp_stride + row + p];
        }
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
          t0 += __shfl_xor_sync(0xffffffffu, t0, o);
          t1 += __shfl_xor_sync(0xffffffffu, t1, o);
          t2 += __shfl_xor_sync(0xffffffffu, t2, o);
        }
        if (lr == 0 && j < s) {
          double* dst = gfunc + (size_t)sig[j] * 3;
          atomicAdd(dst, t0);
          atomicAdd(dst + 1, t1);
          atomicAdd(dst + 2, t2);
        }
      }
  }
  cp_async_wait<0>();
}

}  // namespace sxc
