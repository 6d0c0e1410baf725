// kernel2.cuh - second functional derivatives on the grid and their contraction with response densities
// (the XC / non-additive kernel of LR-TDDFT and subsystem TDDFT).
//
// Reference: Kernel<SCFMode>::calculateDerivatives / storeDerivatives (src/postHF/LRSCF/Kernel/Kernel.cpp:476-683, :686-747)
// obtain d2F/drho2, d2F/drho d(grad rho), d2F/d(grad rho)2 from FunctionalLibrary::calcData(GRADIENTS, ..., order 2), i.e.
// from xcfun_eval with vars XC_N_NX_NY_NZ / XC_A_B_AX_AY_AZ_BX_BY_BZ (dft/functionals/wrappers/XCFun.cpp:89-112, rows :288-298,
// :486-528).  KernelSigmavector<SCFMode>::contractBlock (postHF/LRSCF/Sigmavectors/KernelSigmavector.cpp:360-497) multiplies
// them with the response density of every trial vector.  Here the energy expressions of functionals.cuh are evaluated on a
// second-order forward-mode jet (value, gradient, packed Hessian) over (rho, sigma) resp. (rho_a, rho_b, s_aa, s_ab, s_bb)
// and the chain rule to the Cartesian gradient variables is applied in closed form.
#pragma once

#include "functionals.cuh"

namespace sxc {

// ---------------------------------------------------------------------------------------- second-order jets
template <int N>
struct Jet2 {
  static constexpr int NH = N * (N + 1) / 2;
  double v;
  double d[N];
  double h[NH];  // upper triangle, row-major: (0,0) (0,1) ... (0,N-1) (1,1) ...
};
template <int N> __host__ __device__ constexpr int hidx(int i, int j) { return i * N - i * (i - 1) / 2 + (j - i); }

// r = f(a) with f'(a) = f1, f''(a) = f2
template <int N> SXC_HD Jet2<N> chain2(const Jet2<N>& a, double f, double f1, double f2) {
  Jet2<N> r;
  r.v = f;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = f1 * a.d[i];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) r.h[hidx<N>(i, j)] = f1 * a.h[hidx<N>(i, j)] + f2 * a.d[i] * a.d[j];
  return r;
}
template <int N> SXC_HD Jet2<N> operator+(const Jet2<N>& a, const Jet2<N>& b) {
  Jet2<N> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
#pragma unroll
  for (int i = 0; i < Jet2<N>::NH; ++i) r.h[i] = a.h[i] + b.h[i];
  return r;
}
template <int N> SXC_HD Jet2<N> operator-(const Jet2<N>& a, const Jet2<N>& b) {
  Jet2<N> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
#pragma unroll
  for (int i = 0; i < Jet2<N>::NH; ++i) r.h[i] = a.h[i] - b.h[i];
  return r;
}
template <int N> SXC_HD Jet2<N> operator-(const Jet2<N>& a) {
  Jet2<N> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
#pragma unroll
  for (int i = 0; i < Jet2<N>::NH; ++i) r.h[i] = -a.h[i];
  return r;
}
template <int N> SXC_HD Jet2<N> operator*(const Jet2<N>& a, const Jet2<N>& b) {
  Jet2<N> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j)
      r.h[hidx<N>(i, j)] = a.h[hidx<N>(i, j)] * b.v + a.v * b.h[hidx<N>(i, j)] + a.d[i] * b.d[j] + a.d[j] * b.d[i];
  return r;
}
template <int N> SXC_HD Jet2<N> jrecip(const Jet2<N>& a) {
  const double inv = 1.0 / a.v;
  return chain2(a, inv, -inv * inv, 2.0 * inv * inv * inv);
}
template <int N> SXC_HD Jet2<N> operator/(const Jet2<N>& a, const Jet2<N>& b) { return a * jrecip(b); }
template <int N> SXC_HD Jet2<N> operator+(const Jet2<N>& a, double b) { Jet2<N> r = a; r.v += b; return r; }
template <int N> SXC_HD Jet2<N> operator+(double b, const Jet2<N>& a) { Jet2<N> r = a; r.v += b; return r; }
template <int N> SXC_HD Jet2<N> operator-(const Jet2<N>& a, double b) { Jet2<N> r = a; r.v -= b; return r; }
template <int N> SXC_HD Jet2<N> operator-(double b, const Jet2<N>& a) { Jet2<N> r = -a; r.v += b; return r; }
template <int N> SXC_HD Jet2<N> operator*(const Jet2<N>& a, double b) { return chain2(a, a.v * b, b, 0.0); }
template <int N> SXC_HD Jet2<N> operator*(double b, const Jet2<N>& a) { return chain2(a, a.v * b, b, 0.0); }
template <int N> SXC_HD Jet2<N> operator/(const Jet2<N>& a, double b) { return chain2(a, a.v / b, 1.0 / b, 0.0); }
template <int N> SXC_HD Jet2<N> operator/(double b, const Jet2<N>& a) {
  const double inv = 1.0 / a.v;
  return chain2(a, b * inv, -b * inv * inv, 2.0 * b * inv * inv * inv);
}
template <int N> SXC_HD Jet2<N> dsqrt(const Jet2<N>& a) {
  const double s = sqrt(a.v);
  return chain2(a, s, 0.5 / s, -0.25 / (s * a.v));
}
template <int N> SXC_HD Jet2<N> dcbrt(const Jet2<N>& a) {
  const double c = cbrt(a.v);
  return chain2(a, c, c / (3.0 * a.v), -2.0 * c / (9.0 * a.v * a.v));
}
template <int N> SXC_HD Jet2<N> dexp(const Jet2<N>& a) {
  const double e = exp(a.v);
  return chain2(a, e, e, e);
}
template <int N> SXC_HD Jet2<N> dexpm1(const Jet2<N>& a) {
  const double e = exp(a.v);
  return chain2(a, expm1(a.v), e, e);
}
template <int N> SXC_HD Jet2<N> dlog(const Jet2<N>& a) {
  const double inv = 1.0 / a.v;
  return chain2(a, log(a.v), inv, -inv * inv);
}
template <int N> SXC_HD Jet2<N> dlog1p(const Jet2<N>& a) {
  const double inv = 1.0 / (1.0 + a.v);
  return chain2(a, log1p(a.v), inv, -inv * inv);
}
template <int N> SXC_HD Jet2<N> datan(const Jet2<N>& a) {
  const double inv = 1.0 / (1.0 + a.v * a.v);
  return chain2(a, atan(a.v), inv, -2.0 * a.v * inv * inv);
}
template <int N> SXC_HD Jet2<N> dasinh(const Jet2<N>& a) {
  const double q = 1.0 / (1.0 + a.v * a.v), r = sqrt(q);
  return chain2(a, asinh(a.v), r, -a.v * q * r);
}
template <int N> SXC_HD Jet2<N> pow43(const Jet2<N>& a) { return a * dcbrt(a); }
template <int N> SXC_HD Jet2<N> pow53(const Jet2<N>& a) { const Jet2<N> c = dcbrt(a); return a * c * c; }

template <int N> SXC_HD Jet2<N> jet_var(double v, int dir) {
  Jet2<N> r;
  r.v = v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (i == dir) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < Jet2<N>::NH; ++i) r.h[i] = 0.0;
  return r;
}

// smallest |grad rho|^2 the order-2 evaluation uses: the energy expressions go through sqrt(sigma), whose derivatives
// are singular at exactly zero although F is analytic in sigma there (xcfun returns NaN at such a point)
constexpr double KERNEL_SIGMA_FLOOR = 1e-40;
constexpr double KERNEL_SCREEN = 1.0e-8;  // Kernel.cpp:496, :606 screeningThreshold (hard-coded in the reference)

// array counts of a kernel store
constexpr int KR_ARRAYS = 10;  // pp, pg x y z, gg xx xy xz yy yz zz
constexpr int KU_ARRAYS = 33;  // pp aa ab bb | pg {x,y,z} x {aa,ab,ba,bb} | gg {xx,xy,xz,yy,yz,zz} x {aa,ab,bb}
constexpr int KU_PG = 3, KU_GG = 15;

// ------------------------------------------------------------------------------------------------------------
// RESTRICTED: store[k][p] += sign * d2F ; then zeroed where rho < 1e-8 (Kernel.cpp:476-520 storeDerivatives).
// XC_N_NX_NY_NZ rows 5..14 (XCFun.cpp:288-298): d2F/dn2, d2F/dn d(grad n), d2F/d(grad n)2.
// The functional is evaluated on literal 128-point blocks with the block skip of XCFun.cpp:133-140.
// store_gga: the store carries the pg/gg arrays (Kernel::_gga); a LDA functional then adds nothing to them but the
// screening still zeroes them (:498-509).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FUNC_BLOCK)
k_kernel2_r(FuncView f, long npts, const int* __restrict__ lit_blocks, const double* __restrict__ dens4, double sign,
            int store_gga, double* __restrict__ store) {
  __shared__ double scratch[32];
  const int lb = lit_blocks ? lit_blocks[blockIdx.x] : blockIdx.x;
  const long first = (long)lb * FUNC_BLOCK;
  const int n = (int)min((long)FUNC_BLOCK, npts - first);
  const int t = threadIdx.x;
  const bool valid = t < n;
  const long p = first + t;
  const double r = valid ? dens4[p] : 0.0;
  const double sum_abs = block_sum(fabs(r), scratch);
  const bool skip = sum_abs < (double)n * 1e-12;
  if (!valid) return;
  const int narr = store_gga ? KR_ARRAYS : 1;
  if (r < KERNEL_SCREEN) {  // covers the block skip and xcfun's tiny-density zero as well
    for (int k = 0; k < narr; ++k) store[(size_t)k * npts + p] = 0.0;
    return;
  }
  if (skip) return;
  double g[3] = {0.0, 0.0, 0.0};
  if (f.gga) {
    g[0] = dens4[npts + p];
    g[1] = dens4[2 * npts + p];
    g[2] = dens4[3 * npts + p];
  }
  const double sigma = fmax(g[0] * g[0] + g[1] * g[1] + g[2] * g[2], KERNEL_SIGMA_FLOOR);
  typedef Jet2<2> T;
  T a = jet_var<2>(0.5 * r, 0), g4 = jet_var<2>(0.25 * sigma, 1);
  a.d[0] = 0.5;
  g4.d[1] = 0.25;
  double Fnn = 0.0, Fns = 0.0, Fss = 0.0, Fs = 0.0;
  for (int c = 0; c < f.ncomp; ++c) {
    const T e = basic_functional<T>(f.id[c], a, a, g4, g4, g4);
    Fnn += f.mix[c] * e.h[0];
    Fns += f.mix[c] * e.h[1];
    Fss += f.mix[c] * e.h[2];
    Fs += f.mix[c] * e.d[1];
  }
  store[p] += sign * Fnn;
  if (f.gga && store_gga) {
    int k = 4;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      store[(size_t)(1 + c) * npts + p] += sign * 2.0 * Fns * g[c];
#pragma unroll
      for (int d = c; d < 3; ++d, ++k)
        store[(size_t)k * npts + p] += sign * (4.0 * Fss * g[c] * g[d] + (c == d ? 2.0 * Fs : 0.0));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// UNRESTRICTED (Kernel.cpp:523-683): dens8 rows rho_a, grad_a, rho_b, grad_b.  xcfun rows (XCFun.cpp:486-528) are the
// derivatives w.r.t. (a, b, ax, ay, az, bx, by, bz); pg.c.st = d2F / d rho_s d(grad_c rho_t); gg.cd.ab = d2F / d(grad_c a)
// d(grad_d b).  Kernel::storeDerivatives copies gg.cd.ab into gg.cd.ba (:580-600), so only aa, ab, bb are stored.
// Screening: rho_a < 1e-8 zeroes every aa / ab / ba entry, rho_b < 1e-8 every ab / ba / bb entry - and gg.xy.aa, which the
// reference's beta branch lists as well (:661).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FUNC_BLOCK)
k_kernel2_u(FuncView f, long npts, const int* __restrict__ lit_blocks, const double* __restrict__ dens8, double sign,
            int store_gga, double* __restrict__ store) {
  __shared__ double scratch[32];
  const int lb = lit_blocks ? lit_blocks[blockIdx.x] : blockIdx.x;
  const long first = (long)lb * FUNC_BLOCK;
  const int n = (int)min((long)FUNC_BLOCK, npts - first);
  const int t = threadIdx.x;
  const bool valid = t < n;
  const long p = first + t;
  const double ra = valid ? dens8[p] : 0.0;
  const double rb = valid ? dens8[4 * npts + p] : 0.0;
  const double sum_a = block_sum(fabs(ra), scratch);
  const double sum_b = block_sum(fabs(rb), scratch);
  const bool skip = sum_a < (double)n * 1e-12 && sum_b < (double)n * 1e-12;
  if (!valid) return;
  const bool za = ra < KERNEL_SCREEN, zb = rb < KERNEL_SCREEN;
  if (!skip && !(ra + rb < 1e-14) && !(za && zb)) {
    double A[3] = {0.0, 0.0, 0.0}, B[3] = {0.0, 0.0, 0.0};
    if (f.gga) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        A[c] = dens8[(size_t)(1 + c) * npts + p];
        B[c] = dens8[(size_t)(5 + c) * npts + p];
      }
    }
    typedef Jet2<5> T;
    const T a = jet_var<5>(fmax(ra, 1e-14), 0), b = jet_var<5>(fmax(rb, 1e-14), 1);
    const T saa = jet_var<5>(fmax(A[0] * A[0] + A[1] * A[1] + A[2] * A[2], KERNEL_SIGMA_FLOOR), 2);
    const T sab = jet_var<5>(A[0] * B[0] + A[1] * B[1] + A[2] * B[2], 3);
    const T sbb = jet_var<5>(fmax(B[0] * B[0] + B[1] * B[1] + B[2] * B[2], KERNEL_SIGMA_FLOOR), 4);
    double D[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, H[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) H[i] = 0.0;
    for (int c = 0; c < f.ncomp; ++c) {
      const T e = basic_functional<T>(f.id[c], a, b, saa, sab, sbb);
#pragma unroll
      for (int i = 0; i < 5; ++i) D[i] += f.mix[c] * e.d[i];
#pragma unroll
      for (int i = 0; i < 15; ++i) H[i] += f.mix[c] * e.h[i];
    }
#define HH(i, j) H[hidx<5>(i, j)]
    store[p] += sign * HH(0, 0);
    store[(size_t)npts + p] += sign * HH(0, 1);
    store[(size_t)2 * npts + p] += sign * HH(1, 1);
    if (f.gga && store_gga) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double* pg = store + (size_t)(KU_PG + 4 * c) * npts + p;
        pg[0] += sign * (2.0 * HH(0, 2) * A[c] + HH(0, 3) * B[c]);                 // aa: d rho_a d grad_c a
        pg[(size_t)npts] += sign * (HH(0, 3) * A[c] + 2.0 * HH(0, 4) * B[c]);       // ab: d rho_a d grad_c b
        pg[(size_t)2 * npts] += sign * (2.0 * HH(1, 2) * A[c] + HH(1, 3) * B[c]);   // ba: d rho_b d grad_c a
        pg[(size_t)3 * npts] += sign * (HH(1, 3) * A[c] + 2.0 * HH(1, 4) * B[c]);   // bb
      }
      int k = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = c; d < 3; ++d, ++k) {
          double* gg = store + (size_t)(KU_GG + 3 * k) * npts + p;
          const double dl = (c == d) ? 1.0 : 0.0;
          gg[0] += sign * (4.0 * HH(2, 2) * A[c] * A[d] + 2.0 * HH(2, 3) * (A[c] * B[d] + B[c] * A[d]) +
                           HH(3, 3) * B[c] * B[d] + 2.0 * D[2] * dl);
          gg[(size_t)npts] += sign * (2.0 * HH(2, 3) * A[c] * A[d] + 4.0 * HH(2, 4) * A[c] * B[d] + HH(3, 3) * B[c] * A[d] +
                                      2.0 * HH(3, 4) * B[c] * B[d] + D[3] * dl);
          gg[(size_t)2 * npts] += sign * (4.0 * HH(4, 4) * B[c] * B[d] + 2.0 * HH(3, 4) * (B[c] * A[d] + A[c] * B[d]) +
                                          HH(3, 3) * A[c] * A[d] + 2.0 * D[4] * dl);
        }
    }
#undef HH
  }
  // screening (after the accumulation, as storeDerivatives does)
  if (za || zb) {
    if (za) store[p] = 0.0;
    store[(size_t)npts + p] = 0.0;
    if (zb) store[(size_t)2 * npts + p] = 0.0;
    if (store_gga) {
      for (int c = 0; c < 3; ++c) {
        double* pg = store + (size_t)(KU_PG + 4 * c) * npts + p;
        if (za) pg[0] = 0.0;
        pg[(size_t)npts] = 0.0;
        pg[(size_t)2 * npts] = 0.0;
        if (zb) pg[(size_t)3 * npts] = 0.0;
      }
      for (int k = 0; k < 6; ++k) {
        double* gg = store + (size_t)(KU_GG + 3 * k) * npts + p;
        if (za || (zb && k == 1)) gg[0] = 0.0;  // k == 1: gg.xy.aa in the beta branch (Kernel.cpp:661)
        gg[(size_t)npts] = 0.0;
        if (zb) gg[(size_t)2 * npts] = 0.0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// KernelSigmavector<SCFMode>::contractBlock (KernelSigmavector.cpp:360-497) on the owned blocks of a chunk.
// dens: response density of one trial vector, rho~ = sum D_ij phi_i phi_j and its gradient, rows [4 * nspin][N]
// (the reference forms p = 1/2 w rho~, g = w sum D_ij grad phi_i phi_j = 1/2 w grad rho~ for symmetric D, :287-301; the
// weights are applied by the scatter that follows).  Up to three stores are summed on the fly (Kernel::getPP/getPG/getGG:
// total-density store + subsystem store when I == J + the exactly-treated-systems store of mixed embedding, Kernel.cpp:170-230).
// mode 0: RESTRICTED singlet (10-array stores); 1: RESTRICTED triplet from UNRESTRICTED stores (aa - ab, :381-404);
// mode 2: UNRESTRICTED.  out rows like dens; accumulate != 0 adds (supersystem contraction, :214-226).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_kernel_apply(long N, int blocksize, const int* __restrict__ block_id, int mode, int gga, const double* __restrict__ st0,
               const double* __restrict__ st1, const double* __restrict__ st2, const double* __restrict__ dens, int accumulate,
               double* __restrict__ out) {
  const long first = (long)block_id[blockIdx.x] * blocksize;
  const int n = (int)min((long)blocksize, N - first);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long p = first + i;
    auto K = [&](int k) {
      return st0[(size_t)k * N + p] + (st1 ? st1[(size_t)k * N + p] : 0.0) + (st2 ? st2[(size_t)k * N + p] : 0.0);
    };
    if (mode == 0 || mode == 1) {
      const double pr = 0.5 * dens[p];
      double g[3] = {0.0, 0.0, 0.0};
      if (gga)
        for (int c = 0; c < 3; ++c) g[c] = 0.5 * dens[(size_t)(1 + c) * N + p];
      double pp, pg[3] = {0.0, 0.0, 0.0}, gg[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      if (mode == 0) {
        pp = K(0);
        if (gga) {
          for (int c = 0; c < 3; ++c) pg[c] = K(1 + c);
          for (int k = 0; k < 6; ++k) gg[k] = K(4 + k);
        }
      } else {
        pp = K(0) - K(1);
        if (gga) {
          for (int c = 0; c < 3; ++c) pg[c] = K(KU_PG + 4 * c) - K(KU_PG + 4 * c + 1);
          for (int k = 0; k < 6; ++k) gg[k] = K(KU_GG + 3 * k) - K(KU_GG + 3 * k + 1);
        }
      }
      double o[4];
      o[0] = pp * pr + pg[0] * g[0] + pg[1] * g[1] + pg[2] * g[2];
      o[1] = pg[0] * pr + gg[0] * g[0] + gg[1] * g[1] + gg[2] * g[2];
      o[2] = pg[1] * pr + gg[1] * g[0] + gg[3] * g[1] + gg[4] * g[2];
      o[3] = pg[2] * pr + gg[2] * g[0] + gg[4] * g[1] + gg[5] * g[2];
      for (int c = 0; c < (gga ? 4 : 1); ++c) {
        double* dst = out + (size_t)c * N + p;
        *dst = (accumulate ? *dst : 0.0) + o[c];
      }
    } else {
      const double pa = 0.5 * dens[p], pb = 0.5 * dens[(size_t)4 * N + p];
      double ga[3] = {0.0, 0.0, 0.0}, gb[3] = {0.0, 0.0, 0.0};
      if (gga)
        for (int c = 0; c < 3; ++c) {
          ga[c] = 0.5 * dens[(size_t)(1 + c) * N + p];
          gb[c] = 0.5 * dens[(size_t)(5 + c) * N + p];
        }
      double oa[4] = {0.0, 0.0, 0.0, 0.0}, ob[4] = {0.0, 0.0, 0.0, 0.0};
      const double ppaa = K(0), ppab = K(1), ppbb = K(2);
      oa[0] = ppaa * pa + ppab * pb;
      ob[0] = ppbb * pb + ppab * pa;
      if (gga) {
        // component index of the symmetric 3x3: (c,d) -> 0 xx 1 xy 2 xz 3 yy 4 yz 5 zz
        const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
        for (int c = 0; c < 3; ++c) {
          const double pgaa = K(KU_PG + 4 * c), pgab = K(KU_PG + 4 * c + 1), pgba = K(KU_PG + 4 * c + 2),
                       pgbb = K(KU_PG + 4 * c + 3);
          oa[0] += pgaa * ga[c] + pgab * gb[c];
          ob[0] += pgba * ga[c] + pgbb * gb[c];
          oa[1 + c] += pgaa * pa + pgba * pb;
          ob[1 + c] += pgab * pa + pgbb * pb;
          for (int d = 0; d < 3; ++d) {
            const int k = sym[c][d];
            const double ggaa = K(KU_GG + 3 * k), ggab = K(KU_GG + 3 * k + 1), ggbb = K(KU_GG + 3 * k + 2);
            oa[1 + c] += ggaa * ga[d] + ggab * gb[d];  // gg.cd.ba := gg.cd.ab in the reference
            ob[1 + c] += ggbb * gb[d] + ggab * ga[d];
          }
        }
      }
      for (int c = 0; c < (gga ? 4 : 1); ++c) {
        double* da = out + (size_t)c * N + p;
        double* db = out + (size_t)(4 + c) * N + p;
        *da = (accumulate ? *da : 0.0) + oa[c];
        *db = (accumulate ? *db : 0.0) + ob[c];
      }
    }
  }
}

}  // namespace sxc
