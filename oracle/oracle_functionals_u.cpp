/*
 * oracle_functionals_u.cpp - spin-polarised LDA/GGA kernels of the CPU oracle (TEST INFRASTRUCTURE, see oracle.h).
 *
 * UNRESTRICTED branch of XCFun::calcData (dft/functionals/wrappers/XCFun.cpp:100-112 vars XC_A_B_AX_AY_AZ_BX_BY_BZ,
 * :129-153 block loop, :330-380 prepareInput/parseOutput).  XCFun differentiates templated energy expressions
 * f(rho_a, rho_b, s_aa, s_ab, s_bb) by forward-mode AD; this file does the same on a five-direction first-order
 * jet written for the host (std::array based, unrelated to the device's Dual type).  The closed-shell limit is
 * cross-checked against the hand-derived kernels of oracle_functionals.c and everything against torch.autograd
 * (tests/functional_reference.py).  PARITY UNPINNED like the restricted kernels (no KAT in the reference).
 *
 * Tiny densities: output is zero when rho_a + rho_b < 1e-14 (xcfun's XCFUN_TINY_DENSITY on the total density);
 * a single spin channel below 1e-14 is raised to 1e-14 before evaluation (xcfun regularises its densvars the same
 * way), its own derivatives are kept as computed at the clamp.
 */
#include <array>
#include <cmath>
#include <cstring>

#include "oracle.h"

namespace {

#define ORC_S double
#include "functionals_jet.inc"
#undef ORC_S

const double TINY = 1e-14;

}  // namespace

extern "C" {

int orc_basic_functional_u(int id, double ra, double rb, double gaa, double gab, double gbb, double* F, double* d5) {
  J a = cst(ra), b = cst(rb), saa = cst(gaa), sab = cst(gab), sbb = cst(gbb);
  a.d[0] = b.d[1] = saa.d[2] = sab.d[3] = sbb.d[4] = 1.0;
  J e;
  if (!basic(id, a, b, saa, sab, sbb, &e)) return -1;
  *F = e.v;
  for (int i = 0; i < ND; ++i) d5[i] = e.d[i];
  return 0;
}

/* 8a-4 UNRESTRICTED: arrays rho[2][npts] (alpha, beta), grad[2][3][npts]; outputs dFdRho[2][npts], dFdGrad[2][3][npts].
 * dF/d(grad rho_a) = 2 v_saa grad rho_a + v_sab grad rho_b (chain rule of XC_A_B_AX_AY_AZ_BX_BY_BZ; the LibXC route
 * spells it out, LibXC.cpp:280-300). */
double orc_functional_on_grid_u(const orc_functional* fn, long npts, const double* w, const double* rho, const double* grad,
                                double* epuv, double* dFdRho, double* dFdGrad) {
  const bool gga = orc_functional_is_gga(fn) && grad != nullptr;
  const long BS = 128;
  const long nblocks = (npts + BS - 1) / BS;
  std::memset(epuv, 0, sizeof(double) * (size_t)npts);
  std::memset(dFdRho, 0, sizeof(double) * 2 * (size_t)npts);
  if (dFdGrad) std::memset(dFdGrad, 0, sizeof(double) * 6 * (size_t)npts);
  const double *ra = rho, *rb = rho + npts;
#pragma omp parallel for schedule(dynamic)
  for (long blk = 0; blk < nblocks; ++blk) {
    const long first = blk * BS;
    const long n = (blk == nblocks - 1) ? npts - first : BS;
    double sa = 0.0, sb = 0.0;
    for (long p = 0; p < n; ++p) {
      sa += std::fabs(ra[first + p]);
      sb += std::fabs(rb[first + p]);
    }
    if (sa < (double)n * 1e-12 && sb < (double)n * 1e-12) continue; /* XCFun.cpp:135-140: skip only if both spins are empty */
    for (long p = first; p < first + n; ++p) {
      if (ra[p] + rb[p] < TINY) continue;
      double ga[3] = {0, 0, 0}, gb[3] = {0, 0, 0};
      if (gga)
        for (int c = 0; c < 3; ++c) {
          ga[c] = grad[(size_t)c * npts + p];
          gb[c] = grad[(size_t)(3 + c) * npts + p];
        }
      J a = cst(std::fmax(ra[p], TINY)), b = cst(std::fmax(rb[p], TINY));
      J saa = cst(ga[0] * ga[0] + ga[1] * ga[1] + ga[2] * ga[2]);
      J sab = cst(ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2]);
      J sbb = cst(gb[0] * gb[0] + gb[1] * gb[1] + gb[2] * gb[2]);
      a.d[0] = b.d[1] = saa.d[2] = sab.d[3] = sbb.d[4] = 1.0;
      double F = 0.0, d[ND] = {0, 0, 0, 0, 0};
      for (int c = 0; c < fn->ncomp; ++c) {
        J e;
        if (!basic(fn->id[c], a, b, saa, sab, sbb, &e)) continue;
        F += fn->mix[c] * e.v;
        for (int i = 0; i < ND; ++i) d[i] += fn->mix[c] * e.d[i];
      }
      epuv[p] = F;
      dFdRho[p] = d[0];
      dFdRho[npts + p] = d[1];
      if (gga && dFdGrad)
        for (int c = 0; c < 3; ++c) {
          dFdGrad[(size_t)c * npts + p] = 2.0 * d[2] * ga[c] + d[3] * gb[c];
          dFdGrad[(size_t)(3 + c) * npts + p] = 2.0 * d[4] * gb[c] + d[3] * ga[c];
        }
    }
  }
  double energy = 0.0;
#pragma omp parallel for reduction(+ : energy)
  for (long p = 0; p < npts; ++p) energy += epuv[p] * w[p];
  return energy;
}

}  // extern "C"
