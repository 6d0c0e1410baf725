/*
 * oracle_kernel2.cpp - second functional derivatives and the LR-TDDFT kernel contraction of the CPU oracle
 * (TEST INFRASTRUCTURE, see oracle.h).  SURVEY.md row f-4.
 *
 * Restates
 *   Kernel<SCFMode>::storeDerivatives            postHF/LRSCF/Kernel/Kernel.cpp:476-520 (RESTRICTED), :523-683 (UNRESTRICTED)
 *   XCFun::calcData(GRADIENTS, order 2) outputs   dft/functionals/wrappers/XCFun.cpp:288-298, :486-528
 *   KernelSigmavector::contractKernel            postHF/LRSCF/Sigmavectors/KernelSigmavector.cpp:254-311
 *   KernelSigmavector::contractBlock             :360-497
 *   KernelSigmavector::numericalIntegration      :313-358, and the F += F^T of calcF (:246-248)
 * The second derivatives come from NESTED first-order forward mode: the energy expressions of functionals_jet.inc are
 * instantiated on a five-direction jet whose scalar type is itself a five-direction jet (a derivative of a derivative) -
 * deliberately a different construction from the device code's packed-Hessian jets (serenity_b200/csrc/kernel2.cuh).
 * PARITY: the reference holds no known-answer value for its kernel sigma vectors (Kernel_test.cpp:51-66 only checks that the
 * calls do not fail); this restatement is pinned to its definition instead - the Hessian to finite differences of the
 * (pinned) first derivatives, the sigma vector to the directional derivative of V_xc[P] (tests/test_kernel_sigma.py).
 */
#include <omp.h>

#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

/* inner jet = the scalar type of the outer one */
struct D1 {
  double v;
  double d[5];
  D1() : v(0.0), d{0.0, 0.0, 0.0, 0.0, 0.0} {}
  D1(double x) : v(x), d{0.0, 0.0, 0.0, 0.0, 0.0} {} /* implicit: constants */
};
inline D1 lift(const D1& a, double f, double fp) {
  D1 r;
  r.v = f;
  for (int i = 0; i < 5; ++i) r.d[i] = fp * a.d[i];
  return r;
}
inline D1 operator+(const D1& a, const D1& b) {
  D1 r;
  r.v = a.v + b.v;
  for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
inline D1 operator-(const D1& a, const D1& b) {
  D1 r;
  r.v = a.v - b.v;
  for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
inline D1 operator-(const D1& a) { return lift(a, -a.v, -1.0); }
inline D1 operator*(const D1& a, const D1& b) {
  D1 r;
  r.v = a.v * b.v;
  for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
inline D1 operator/(const D1& a, const D1& b) {
  D1 r;
  r.v = a.v / b.v;
  for (int i = 0; i < 5; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
inline bool operator>(const D1& a, const D1& b) { return a.v > b.v; }
inline bool operator>=(const D1& a, const D1& b) { return a.v >= b.v; }
inline bool operator<(const D1& a, const D1& b) { return a.v < b.v; }
inline D1 pow(const D1& a, double p) { return lift(a, std::pow(a.v, p), p * std::pow(a.v, p - 1.0)); }
inline D1 sqrt(const D1& a) {
  const double s = std::sqrt(a.v);
  return lift(a, s, 0.5 / s);
}
inline D1 cbrt(const D1& a) {
  const double c = std::cbrt(a.v);
  return lift(a, c, c / (3.0 * a.v));
}
inline D1 exp(const D1& a) { return lift(a, std::exp(a.v), std::exp(a.v)); }
inline D1 log(const D1& a) { return lift(a, std::log(a.v), 1.0 / a.v); }
inline D1 atan(const D1& a) { return lift(a, std::atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
inline D1 asinh(const D1& a) { return lift(a, std::asinh(a.v), 1.0 / std::sqrt(1.0 + a.v * a.v)); }

#define ORC_S D1
#include "functionals_jet.inc"
#undef ORC_S

const double TINY = 1e-14;
const double SIGMA_FLOOR = 1e-40; /* sqrt(sigma) is singular at exactly 0 although F is analytic there (xcfun: NaN) */
double SCREEN = 1.0e-8;           /* Kernel.cpp:496, :606; orc_kernel_set_screen() lets a test quantify its effect */

/* F, first derivatives d5 and Hessian h[5][5] of the composite functional w.r.t. (rho_a, rho_b, s_aa, s_ab, s_bb) */
void composite_d2(const orc_functional* fn, double ra, double rb, double gaa, double gab, double gbb, double* F, double* d5,
                  double (*h)[5]) {
  J v[5];
  const double x[5] = {ra, rb, gaa, gab, gbb};
  for (int i = 0; i < 5; ++i) {
    D1 s(x[i]);
    s.d[i] = 1.0; /* inner seed */
    v[i] = cst(s);
    v[i].d[i] = D1(1.0); /* outer seed */
  }
  *F = 0.0;
  for (int i = 0; i < 5; ++i) {
    d5[i] = 0.0;
    for (int j = 0; j < 5; ++j) h[i][j] = 0.0;
  }
  for (int c = 0; c < fn->ncomp; ++c) {
    J e;
    if (!basic(fn->id[c], v[0], v[1], v[2], v[3], v[4], &e)) continue;
    *F += fn->mix[c] * e.v.v;
    for (int i = 0; i < 5; ++i) {
      d5[i] += fn->mix[c] * e.d[i].v;
      for (int j = 0; j < 5; ++j) h[i][j] += fn->mix[c] * e.d[i].d[j];
    }
  }
}

inline long block_len(long npts, long first) { return npts - first < 128 ? npts - first : 128; }

}  // namespace

extern "C" {

/* test hook: the reference hard-codes 1e-8; with 0 the sigma matrix is the exact directional derivative of V_xc */
void orc_kernel_set_screen(double thr) { SCREEN = thr; }

int orc_basic_functional_d2(int id, double ra, double rb, double gaa, double gab, double gbb, double* F, double* d5,
                            double* h25) {
  const double one = 1.0;
  orc_functional fn = {1, &id, &one};
  double h[5][5];
  composite_d2(&fn, ra, rb, gaa, gab, gbb, F, d5, h);
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) h25[5 * i + j] = h[i][j];
  J probe;
  return basic(id, cst(D1(1.0)), cst(D1(1.0)), cst(D1(1.0)), cst(D1(1.0)), cst(D1(1.0)), &probe) ? 0 : -1;
}

/* Kernel<RESTRICTED>::storeDerivatives (Kernel.cpp:476-520): store [10][npts] = pp, pg x y z, gg xx xy xz yy yz zz
 * ([1][npts] if !store_gga) is ADDED to with factor pm = sign, then zeroed where rho < 1e-8.  The closed-shell values of
 * XC_N_NX_NY_NZ follow from F(n, sigma) = f(n/2, n/2, sigma/4, sigma/4, sigma/4). */
void orc_kernel_store_r(const orc_functional* fn, long npts, const double* rho, const double* gx, const double* gy,
                        const double* gz, double sign, int store_gga, double* store) {
  const int gga = orc_functional_is_gga(fn) && gx != nullptr;
  const long nblocks = (npts + 127) / 128;
#pragma omp parallel for schedule(dynamic)
  for (long blk = 0; blk < nblocks; ++blk) {
    const long first = blk * 128, n = block_len(npts, first);
    double sa = 0.0;
    for (long p = 0; p < n; ++p) sa += std::fabs(rho[first + p]);
    const bool skip = sa < (double)n * 1e-12; /* XCFun.cpp:133-140 */
    for (long p = first; p < first + n; ++p) {
      if (!skip && !(rho[p] < TINY) && fn->ncomp > 0) {
        double g[3] = {0.0, 0.0, 0.0};
        if (gga) {
          g[0] = gx[p];
          g[1] = gy[p];
          g[2] = gz[p];
        }
        const double sigma = std::fmax(g[0] * g[0] + g[1] * g[1] + g[2] * g[2], SIGMA_FLOOR);
        double F, d5[5], h[5][5];
        composite_d2(fn, 0.5 * rho[p], 0.5 * rho[p], 0.25 * sigma, 0.25 * sigma, 0.25 * sigma, &F, d5, h);
        double Fnn = 0.0, Fns = 0.0, Fss = 0.0, Fs = 0.0;
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) Fnn += 0.25 * h[i][j];
        for (int i = 0; i < 2; ++i)
          for (int k = 2; k < 5; ++k) Fns += 0.125 * h[i][k];
        for (int k = 2; k < 5; ++k) {
          Fs += 0.25 * d5[k];
          for (int l = 2; l < 5; ++l) Fss += 0.0625 * h[k][l];
        }
        store[p] += sign * Fnn; /* :484 */
        if (gga && store_gga) {  /* :485-495 */
          int k = 4;
          for (int c = 0; c < 3; ++c) {
            store[(size_t)(1 + c) * npts + p] += sign * 2.0 * Fns * g[c];
            for (int d = c; d < 3; ++d, ++k)
              store[(size_t)k * npts + p] += sign * (4.0 * Fss * g[c] * g[d] + (c == d ? 2.0 * Fs : 0.0));
          }
        }
      }
      if (rho[p] < SCREEN) /* :497-511 */
        for (int k = 0; k < (store_gga ? 10 : 1); ++k) store[(size_t)k * npts + p] = 0.0;
    }
  }
}

/* Kernel<UNRESTRICTED>::storeDerivatives (Kernel.cpp:523-683).  rho [2][npts], grad [2][3][npts] (NULL: LDA).
 * store [33][npts]: pp aa ab bb | pg {x,y,z} x {aa,ab,ba,bb} | gg {xx,xy,xz,yy,yz,zz} x {aa,ab,bb}; the reference fills
 * gg.cd.ba with gg.cd.ab (:580-600), so ba is not stored.  pg.c.st = d2F/d rho_s d(grad_c rho_t), gg.cd.ab =
 * d2F/d(grad_c rho_a) d(grad_d rho_b)  (xcfun rows of XC_A_B_AX_AY_AZ_BX_BY_BZ, XCFun.cpp:486-528). */
void orc_kernel_store_u(const orc_functional* fn, long npts, const double* rho, const double* grad, double sign,
                        int store_gga, double* store) {
  const int gga = orc_functional_is_gga(fn) && grad != nullptr;
  const long nblocks = (npts + 127) / 128;
  const double *ra = rho, *rb = rho + npts;
#pragma omp parallel for schedule(dynamic)
  for (long blk = 0; blk < nblocks; ++blk) {
    const long first = blk * 128, n = block_len(npts, first);
    double sa = 0.0, sb = 0.0;
    for (long p = 0; p < n; ++p) {
      sa += std::fabs(ra[first + p]);
      sb += std::fabs(rb[first + p]);
    }
    const bool skip = sa < (double)n * 1e-12 && sb < (double)n * 1e-12;
    for (long p = first; p < first + n; ++p) {
      if (!skip && !(ra[p] + rb[p] < TINY) && fn->ncomp > 0) {
        double A[3] = {0, 0, 0}, B[3] = {0, 0, 0};
        if (gga)
          for (int c = 0; c < 3; ++c) {
            A[c] = grad[(size_t)c * npts + p];
            B[c] = grad[(size_t)(3 + c) * npts + p];
          }
        double F, D[5], H[5][5];
        composite_d2(fn, std::fmax(ra[p], TINY), std::fmax(rb[p], TINY),
                     std::fmax(A[0] * A[0] + A[1] * A[1] + A[2] * A[2], SIGMA_FLOOR), A[0] * B[0] + A[1] * B[1] + A[2] * B[2],
                     std::fmax(B[0] * B[0] + B[1] * B[1] + B[2] * B[2], SIGMA_FLOOR), &F, D, H);
        store[p] += sign * H[0][0];
        store[(size_t)npts + p] += sign * H[0][1];
        store[(size_t)2 * npts + p] += sign * H[1][1];
        if (gga && store_gga) {
          /* d sigma / d grad_c a = (2 A_c, B_c, 0), d sigma / d grad_c b = (0, A_c, 2 B_c) for (s_aa, s_ab, s_bb) */
          for (int c = 0; c < 3; ++c) {
            const double ua[3] = {2.0 * A[c], B[c], 0.0}, ub[3] = {0.0, A[c], 2.0 * B[c]};
            double pgaa = 0, pgab = 0, pgba = 0, pgbb = 0;
            for (int k = 0; k < 3; ++k) {
              pgaa += H[0][2 + k] * ua[k];
              pgab += H[0][2 + k] * ub[k];
              pgba += H[1][2 + k] * ua[k];
              pgbb += H[1][2 + k] * ub[k];
            }
            double* pg = store + (size_t)(3 + 4 * c) * npts + p;
            pg[0] += sign * pgaa;
            pg[(size_t)npts] += sign * pgab;
            pg[(size_t)2 * npts] += sign * pgba;
            pg[(size_t)3 * npts] += sign * pgbb;
          }
          int kk = 0;
          for (int c = 0; c < 3; ++c)
            for (int d = c; d < 3; ++d, ++kk) {
              const double uac[3] = {2.0 * A[c], B[c], 0.0}, uad[3] = {2.0 * A[d], B[d], 0.0};
              const double ubc[3] = {0.0, A[c], 2.0 * B[c]}, ubd[3] = {0.0, A[d], 2.0 * B[d]};
              double aa = 0, ab = 0, bb = 0;
              for (int k = 0; k < 3; ++k)
                for (int l = 0; l < 3; ++l) {
                  aa += H[2 + k][2 + l] * uac[k] * uad[l];
                  ab += H[2 + k][2 + l] * uac[k] * ubd[l];
                  bb += H[2 + k][2 + l] * ubc[k] * ubd[l];
                }
              if (c == d) { /* second derivatives of the invariants themselves */
                aa += 2.0 * D[2];
                ab += D[3];
                bb += 2.0 * D[4];
              }
              double* gg = store + (size_t)(15 + 3 * kk) * npts + p;
              gg[0] += sign * aa;
              gg[(size_t)npts] += sign * ab;
              gg[(size_t)2 * npts] += sign * bb;
            }
        }
      }
      const bool za = ra[p] < SCREEN, zb = rb[p] < SCREEN; /* :606-680 */
      if (za) store[p] = 0.0;
      if (za || zb) store[(size_t)npts + p] = 0.0;
      if (zb) store[(size_t)2 * npts + p] = 0.0;
      if (store_gga && (za || zb)) {
        for (int c = 0; c < 3; ++c) {
          double* pg = store + (size_t)(3 + 4 * c) * npts + p;
          if (za) pg[0] = 0.0;
          pg[(size_t)npts] = 0.0;
          pg[(size_t)2 * npts] = 0.0;
          if (zb) pg[(size_t)3 * npts] = 0.0;
        }
        for (int k = 0; k < 6; ++k) {
          double* gg = store + (size_t)(15 + 3 * k) * npts + p;
          if (za || (zb && k == 1)) gg[0] = 0.0; /* gg.xy.aa is in the beta list of the reference (:661) */
          gg[(size_t)npts] = 0.0;
          if (zb) gg[(size_t)2 * npts] = 0.0;
        }
      }
    }
  }
}

/* KernelSigmavector::contractKernel + contractBlock for ONE trial vector.
 * D: nspin matrices nbJ x nbJ (column-major, back to back), symmetrised here as calcF does (:201-208: D += D^T).
 * mode 0 RESTRICTED singlet (store [10][N]), 1 RESTRICTED triplet from an UNRESTRICTED store (aa - ab, :381-404),
 * 2 UNRESTRICTED (store [33][N]).  resp [4 * nspin][N] (scalar, gradient x y z per spin) is ADDED to (the reference adds
 * into scalar/gradient, :366-379).  store = Kernel::getPP/getPG/getGG already summed by the caller. */
void orc_kernel_contract(const orc_basis* b, const orc_grid* g, double radial_thr, double block_ave_thr, int mode, int gga,
                         const double* store, const double* D_in, double* resp) {
  const int nbf = b->nbf, nspin = mode == 2 ? 2 : 1;
  const long N = g->npts;
  const int nblocks = orc_nblocks(g);
  std::vector<double> D((size_t)nspin * nbf * nbf);
  for (int sp = 0; sp < nspin; ++sp)
    for (int i = 0; i < nbf; ++i)
      for (int j = 0; j < nbf; ++j)
        D[(size_t)sp * nbf * nbf + i + (size_t)j * nbf] =
            D_in[(size_t)sp * nbf * nbf + i + (size_t)j * nbf] + D_in[(size_t)sp * nbf * nbf + j + (size_t)i * nbf];
#pragma omp parallel
  {
    const size_t sz = (size_t)g->blocksize * nbf;
    std::vector<double> val(sz), dx(sz), dy(sz), dz(sz), ave(nbf), contr(g->blocksize);
    std::vector<int> neg(nbf);
    std::vector<double> sc(2 * (size_t)g->blocksize), gr(6 * (size_t)g->blocksize);
    double centre[3];
#pragma omp for schedule(dynamic)
    for (int blk = 0; blk < nblocks; ++blk) {
      const int n = orc_basis_block(b, g, radial_thr, gga ? 1 : 0, blk, val.data(), dx.data(), dy.data(), dz.data(), nullptr,
                                    nullptr, nullptr, nullptr, nullptr, nullptr, neg.data(), centre);
      const long first = (long)blk * g->blocksize;
      for (int i = 0; i < nbf; ++i) { /* averageFunctionValues, BasisFunctionOnGridController.cpp:1100 */
        double s = 0.0;
        for (int p = 0; p < n; ++p) s += std::fabs(val[(size_t)i * n + p]);
        ave[i] = s / n;
      }
      const double* w = g->w + first;
      for (int sp = 0; sp < nspin; ++sp) {
        double* scal = sc.data() + (size_t)sp * n;
        double* grd[3] = {gr.data() + (size_t)(3 * sp) * n, gr.data() + (size_t)(3 * sp + 1) * n,
                          gr.data() + (size_t)(3 * sp + 2) * n};
        std::fill(scal, scal + n, 0.0);
        for (int c = 0; c < 3; ++c) std::fill(grd[c], grd[c] + n, 0.0);
        const double* Ds = D.data() + (size_t)sp * nbf * nbf;
        for (int i = 0; i < nbf; ++i)
          for (int j = 0; j < nbf; ++j) {
            const double density = Ds[(size_t)i * nbf + j];
            if (std::fabs(density * ave[i] * ave[j]) > block_ave_thr) { /* :283 */
              for (int p = 0; p < n; ++p) contr[p] = val[(size_t)j * n + p] * density;
              for (int p = 0; p < n; ++p) scal[p] += val[(size_t)i * n + p] * contr[p];
              if (gga)
                for (int p = 0; p < n; ++p) {
                  grd[0][p] += dx[(size_t)i * n + p] * contr[p];
                  grd[1][p] += dy[(size_t)i * n + p] * contr[p];
                  grd[2][p] += dz[(size_t)i * n + p] * contr[p];
                }
            }
          }
        for (int p = 0; p < n; ++p) scal[p] = 0.5 * w[p] * scal[p]; /* :294 */
        if (gga)
          for (int c = 0; c < 3; ++c)
            for (int p = 0; p < n; ++p) grd[c][p] = w[p] * grd[c][p];
      }
      /* contractBlock */
      static const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
      for (int p = 0; p < n; ++p) {
        const long q = first + p;
        auto K = [&](int k) { return store[(size_t)k * N + q]; };
        if (mode != 2) {
          const double pr = sc[p];
          const double gv[3] = {gr[p], gr[(size_t)n + p], gr[(size_t)2 * n + p]};
          double pp, pg[3] = {0, 0, 0}, gg[6] = {0, 0, 0, 0, 0, 0};
          if (mode == 0) {
            pp = K(0);
            if (gga) {
              for (int c = 0; c < 3; ++c) pg[c] = K(1 + c);
              for (int k = 0; k < 6; ++k) gg[k] = K(4 + k);
            }
          } else {
            pp = K(0) - K(1);
            if (gga) {
              for (int c = 0; c < 3; ++c) pg[c] = K(3 + 4 * c) - K(3 + 4 * c + 1);
              for (int k = 0; k < 6; ++k) gg[k] = K(15 + 3 * k) - K(15 + 3 * k + 1);
            }
          }
          resp[q] += pp * pr;
          if (gga) {
            for (int c = 0; c < 3; ++c) resp[q] += pg[c] * gv[c];
            for (int c = 0; c < 3; ++c) {
              double* o = resp + (size_t)(1 + c) * N + q;
              *o += pg[c] * pr;
              for (int d = 0; d < 3; ++d) *o += gg[sym[c][d]] * gv[d];
            }
          }
        } else {
          const double pa = sc[p], pb = sc[(size_t)n + p];
          double ga[3], gb[3];
          for (int c = 0; c < 3; ++c) {
            ga[c] = gr[(size_t)c * n + p];
            gb[c] = gr[(size_t)(3 + c) * n + p];
          }
          double* oa = resp + q;
          double* ob = resp + (size_t)4 * N + q;
          oa[0] += K(0) * pa + K(1) * pb;
          ob[0] += K(2) * pb + K(1) * pa;
          if (gga)
            for (int c = 0; c < 3; ++c) {
              const double pgaa = K(3 + 4 * c), pgab = K(3 + 4 * c + 1), pgba = K(3 + 4 * c + 2), pgbb = K(3 + 4 * c + 3);
              oa[0] += pgaa * ga[c] + pgab * gb[c];
              ob[0] += pgba * ga[c] + pgbb * gb[c];
              oa[(size_t)(1 + c) * N] += pgaa * pa + pgba * pb;
              ob[(size_t)(1 + c) * N] += pgab * pa + pgbb * pb;
              for (int d = 0; d < 3; ++d) {
                const int k = sym[c][d];
                oa[(size_t)(1 + c) * N] += K(15 + 3 * k) * ga[d] + K(15 + 3 * k + 1) * gb[d];
                ob[(size_t)(1 + c) * N] += K(15 + 3 * k + 2) * gb[d] + K(15 + 3 * k + 1) * ga[d]; /* ba := ab */
              }
            }
        }
      }
    }
  }
}

/* KernelSigmavector::numericalIntegration (:313-358) + the thread sum and F += F^T of calcF (:236-249) for ONE trial vector.
 * resp [4 * nspin][N]; F nspin matrices nbI x nbI column-major, overwritten. */
void orc_kernel_integrate(const orc_basis* b, const orc_grid* g, double radial_thr, double block_ave_thr, int gga, int nspin,
                          const double* resp, double* F) {
  const int nbf = b->nbf;
  const long N = g->npts;
  const int nblocks = orc_nblocks(g);
  const size_t nb2 = (size_t)nbf * nbf;
  const int nthreads = orc_max_threads();
  std::vector<double> Fxc((size_t)nthreads * nspin * nb2, 0.0);
#pragma omp parallel
  {
    const size_t sz = (size_t)g->blocksize * nbf;
    std::vector<double> val(sz), dx(sz), dy(sz), dz(sz), ave(nbf), contr(g->blocksize);
    std::vector<int> neg(nbf);
    double centre[3];
    double* mine = Fxc.data() + (size_t)omp_get_thread_num() * nspin * nb2;
#pragma omp for schedule(dynamic)
    for (int blk = 0; blk < nblocks; ++blk) {
      const int n = orc_basis_block(b, g, radial_thr, gga ? 1 : 0, blk, val.data(), dx.data(), dy.data(), dz.data(), nullptr,
                                    nullptr, nullptr, nullptr, nullptr, nullptr, neg.data(), centre);
      const long first = (long)blk * g->blocksize;
      for (int i = 0; i < nbf; ++i) {
        double s = 0.0;
        for (int p = 0; p < n; ++p) s += std::fabs(val[(size_t)i * n + p]);
        ave[i] = s / n;
      }
      for (int sp = 0; sp < nspin; ++sp) {
        const double* scal = resp + (size_t)(4 * sp) * N + first;
        const double* gx = resp + (size_t)(4 * sp + 1) * N + first;
        const double* gy = resp + (size_t)(4 * sp + 2) * N + first;
        const double* gz = resp + (size_t)(4 * sp + 3) * N + first;
        double* Fs = mine + (size_t)sp * nb2;
        double scal_sum = 0.0;
        for (int p = 0; p < n; ++p) scal_sum += std::fabs(scal[p]);
        for (int i = 0; i < nbf; ++i)
          for (int j = 0; j < nbf; ++j)
            if (scal_sum * ave[i] * ave[j] > block_ave_thr) { /* :340 */
              double dot = 0.0;
              for (int p = 0; p < n; ++p) {
                double c = 0.5 * scal[p] * val[(size_t)j * n + p];
                if (gga) c += gx[p] * dx[(size_t)j * n + p] + gy[p] * dy[(size_t)j * n + p] + gz[p] * dz[(size_t)j * n + p];
                dot += c * val[(size_t)i * n + p];
              }
              Fs[(size_t)i * nbf + j] += dot;
            }
      }
    }
  }
  for (int sp = 0; sp < nspin; ++sp) {
    double* Fo = F + (size_t)sp * nb2;
    std::fill(Fo, Fo + nb2, 0.0);
    for (int t = 0; t < nthreads; ++t) {
      const double* part = Fxc.data() + ((size_t)t * nspin + sp) * nb2;
      for (size_t k = 0; k < nb2; ++k) Fo[k] += part[k];
    }
    for (int i = 0; i < nbf; ++i) /* F += F^T */
      for (int j = i; j < nbf; ++j) {
        const double s = Fo[(size_t)i * nbf + j] + Fo[(size_t)j * nbf + i];
        Fo[(size_t)i * nbf + j] = Fo[(size_t)j * nbf + i] = s;
      }
  }
}

}  // extern "C"
