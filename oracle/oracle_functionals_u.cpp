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

constexpr int ND = 5;  // d/d rho_a, rho_b, s_aa, s_ab, s_bb
struct J {
  double v;
  std::array<double, ND> d;
};
inline J cst(double v) {
  J r{v, {}};
  return r;
}
inline J un(const J& a, double f, double fp) {
  J r;
  r.v = f;
  for (int i = 0; i < ND; ++i) r.d[i] = fp * a.d[i];
  return r;
}
inline J operator+(const J& a, const J& b) {
  J r;
  r.v = a.v + b.v;
  for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
inline J operator-(const J& a, const J& b) {
  J r;
  r.v = a.v - b.v;
  for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
inline J operator-(const J& a) { return un(a, -a.v, -1.0); }
inline J operator*(const J& a, const J& b) {
  J r;
  r.v = a.v * b.v;
  for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
inline J operator/(const J& a, const J& b) {
  J r;
  r.v = a.v / b.v;
  for (int i = 0; i < ND; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
inline J operator+(const J& a, double b) { return un(a, a.v + b, 1.0); }
inline J operator+(double b, const J& a) { return un(a, a.v + b, 1.0); }
inline J operator-(const J& a, double b) { return un(a, a.v - b, 1.0); }
inline J operator-(double b, const J& a) { return un(a, b - a.v, -1.0); }
inline J operator*(const J& a, double b) { return un(a, a.v * b, b); }
inline J operator*(double b, const J& a) { return un(a, a.v * b, b); }
inline J operator/(const J& a, double b) { return un(a, a.v / b, 1.0 / b); }
inline J operator/(double b, const J& a) { return un(a, b / a.v, -b / (a.v * a.v)); }
inline J jpow(const J& a, double p) { return un(a, std::pow(a.v, p), p * std::pow(a.v, p - 1.0)); }
inline J jsqrt(const J& a) {
  const double s = std::sqrt(a.v);
  return un(a, s, a.v > 0.0 ? 0.5 / s : 0.0);
}
inline J jexp(const J& a) { return un(a, std::exp(a.v), std::exp(a.v)); }
inline J jlog(const J& a) { return un(a, std::log(a.v), 1.0 / a.v); }
inline J jatan(const J& a) { return un(a, std::atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
inline J jasinh(const J& a) { return un(a, std::asinh(a.v), 1.0 / std::sqrt(1.0 + a.v * a.v)); }

const double PI = 3.14159265358979323846;
const double CF = 0.3 * std::pow(3.0 * PI * PI, 2.0 / 3.0);

/* f(zeta) with 1 + zeta = 2 rho_a / n, 1 - zeta = 2 rho_b / n formed from the spin densities (no cancellation at
 * fully polarised points) */
J f_zeta(const J& a, const J& b) {
  const J n = a + b;
  return (jpow(2.0 * a / n, 4.0 / 3.0) + jpow(2.0 * b / n, 4.0 / 3.0) - 2.0) / (std::pow(2.0, 4.0 / 3.0) - 2.0);
}
const double FPP0 = 4.0 / (9.0 * (std::cbrt(2.0) - 1.0));

J slaterx(const J& a, const J& b) { return (-0.75 * std::cbrt(6.0 / PI)) * (jpow(a, 4.0 / 3.0) + jpow(b, 4.0 / 3.0)); }

J vwn_eps(const J& x, double A, double x0, double b, double c) {
  const double Q = std::sqrt(4.0 * c - b * b);
  const J X = x * x + b * x + c;
  const double X0 = x0 * x0 + b * x0 + c;
  const J at = jatan(Q / (2.0 * x + b));
  const J xm = x - x0;
  return A * (jlog(x * x / X) + (2.0 * b / Q) * at - (b * x0 / X0) * (jlog(xm * xm / X) + (2.0 * (b + 2.0 * x0) / Q) * at));
}
J vwn5c(const J& a, const J& b) {
  const J n = a + b, z = (a - b) / n;
  const J x = jpow(3.0 / (4.0 * PI * n), 1.0 / 6.0);
  const J eP = vwn_eps(x, 0.0310907, -0.10498, 3.72744, 12.9352);
  const J eF = vwn_eps(x, 0.01554535, -0.32500, 7.06042, 18.0578);
  const J ac = vwn_eps(x, -1.0 / (6.0 * PI * PI), -0.0047584, 1.13107, 13.0045);
  const J fz = f_zeta(a, b), z4 = z * z * z * z;
  return n * (eP + ac * fz * (1.0 - z4) / FPP0 + (eF - eP) * fz * z4);
}
J tfk(const J& a, const J& b) { return (std::pow(2.0, 2.0 / 3.0) * CF) * (jpow(a, 5.0 / 3.0) + jpow(b, 5.0 / 3.0)); }

J pbex_cs(const J& n, const J& g) {
  const double kappa = 0.804, mu = 0.2195149727645171;
  const J s2 = g / (4.0 * std::pow(3.0 * PI * PI, 2.0 / 3.0) * jpow(n, 8.0 / 3.0));
  const J Fx = (1.0 + kappa) - kappa / (1.0 + (mu / kappa) * s2);
  return (-0.75 * std::cbrt(3.0 / PI)) * jpow(n, 4.0 / 3.0) * Fx;
}
J pbex(const J& a, const J& b, const J& gaa, const J& gbb) { return 0.5 * (pbex_cs(2.0 * a, 4.0 * gaa) + pbex_cs(2.0 * b, 4.0 * gbb)); }

J b88_spin(const J& r, const J& g) {
  const double beta = 0.0042;
  const J r43 = jpow(r, 4.0 / 3.0);
  const J x = jsqrt(g) / r43;
  return (-beta) * r43 * x * x / (1.0 + 6.0 * beta * x * jasinh(x));
}

J lypc(const J& a, const J& b, const J& gaa, const J& gab, const J& gbb) {
  const double A = 0.04918, B = 0.132, C = 0.2533, D = 0.349;
  const J n = a + b, g = gaa + 2.0 * gab + gbb;
  const J q = jpow(n, -1.0 / 3.0);
  const J den = 1.0 + D * q;
  const J omega = jexp(-C * q) * jpow(n, -11.0 / 3.0) / den;
  const J delta = C * q + D * q / den;
  const J t = a * b * (std::pow(2.0, 11.0 / 3.0) * CF * (jpow(a, 8.0 / 3.0) + jpow(b, 8.0 / 3.0)) + (47.0 / 18.0 - 7.0 / 18.0 * delta) * g -
                       (2.5 - delta / 18.0) * (gaa + gbb) - (delta - 11.0) / 9.0 * (a * gaa + b * gbb) / n) -
              (2.0 / 3.0) * n * n * g + ((2.0 / 3.0) * n * n - a * a) * gbb + ((2.0 / 3.0) * n * n - b * b) * gaa;
  return (-A * 4.0) * a * b / (den * n) - (A * B) * omega * t;
}

J pw92_G(const J& rs, double A, double a1, double b1, double b2, double b3, double b4) {
  const J srs = jsqrt(rs);
  const J q1 = (2.0 * A) * (b1 * srs + b2 * rs + b3 * rs * srs + b4 * rs * rs);
  return (-2.0 * A) * (1.0 + a1 * rs) * jlog(1.0 + 1.0 / q1);
}
J pw92_eps(const J& rs, const J& z, const J& a, const J& b) {
  const J e0 = pw92_G(rs, 0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294);
  const J e1 = pw92_G(rs, 0.01554535, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517);
  const J mac = pw92_G(rs, 0.0168869, 0.11125, 10.357, 3.6231, 0.88026, 0.49671);
  const J fz = f_zeta(a, b), z4 = z * z * z * z;
  return e0 - mac * fz * (1.0 - z4) / FPP0 + (e1 - e0) * fz * z4;
}
J pbec(const J& a, const J& b, const J& gaa, const J& gab, const J& gbb) {
  const double beta = 0.06672455060314922, gamma = (1.0 - std::log(2.0)) / (PI * PI);
  const J n = a + b, g = gaa + 2.0 * gab + gbb, z = (a - b) / n;
  const J rs = jpow(3.0 / (4.0 * PI * n), 1.0 / 3.0);
  const J eps = pw92_eps(rs, z, a, b);
  const J phi = 0.5 * (jpow(2.0 * a / n, 2.0 / 3.0) + jpow(2.0 * b / n, 2.0 / 3.0));
  const J phi3 = phi * phi * phi;
  const J kF = jpow(3.0 * PI * PI * n, 1.0 / 3.0);
  const J t2 = g * (PI / 16.0) / (phi * phi * kF * n * n);
  const J Aa = (beta / gamma) / (jexp(-eps / (gamma * phi3)) - 1.0);
  const J At2 = Aa * t2;
  const J H = gamma * phi3 * jlog(1.0 + (beta / gamma) * t2 * (1.0 + At2) / (1.0 + At2 + At2 * At2));
  return n * (eps + H);
}

J pz81(const J& rs, double g, double b1, double b2, double A, double B, double C, double D) {
  if (rs.v >= 1.0) return g / (1.0 + b1 * jsqrt(rs) + b2 * rs);
  const J lr = jlog(rs);
  return A * lr + B + C * rs * lr + D * rs;
}
J p86c(const J& a, const J& b, const J& gaa, const J& gab, const J& gbb) {
  const J n = a + b, g = gaa + 2.0 * gab + gbb;
  const J rs = jpow(3.0 / (4.0 * PI * n), 1.0 / 3.0);
  const J eU = pz81(rs, -0.1423, 1.0529, 0.3334, 0.0311, -0.048, 0.0020, -0.0116);
  const J eP = pz81(rs, -0.0843, 1.3981, 0.2611, 0.01555, -0.0269, 0.0007, -0.0048);
  const J eps = eU + f_zeta(a, b) * (eP - eU);
  const J rs2 = rs * rs;
  const J Cn = 0.001667 + (0.002568 + 0.023266 * rs + 7.389e-6 * rs2) / (1.0 + 8.723 * rs + 0.472 * rs2 + 0.07389 * rs2 * rs);
  const J Phi = (1.7454151061251240 /* (9 pi)^(1/6) */ * 0.11 * 0.004235) * jsqrt(g) / (Cn * jpow(n, 7.0 / 6.0));
  const J d = std::cbrt(2.0) * jsqrt(jpow(a / n, 5.0 / 3.0) + jpow(b / n, 5.0 / 3.0));
  return n * eps + jexp(-Phi) * Cn * g / (d * jpow(n, 4.0 / 3.0));
}

J lc94_spin(const J& r, const J& g) {
  const double a1 = 0.093907, a2 = 76.320, a3 = 0.26608, a4 = 0.0809615, aa = 100.0, bb = 0.57767e-4;
  const J s = jsqrt(g) / ((2.0 * std::cbrt(6.0 * PI * PI)) * jpow(r, 4.0 / 3.0));
  const J s2 = s * s;
  const J L = a1 * s * jasinh(a2 * s);
  const J F = (1.0 + L + (a3 - a4 * jexp(-aa * s2)) * s2) / (1.0 + L + bb * s2 * s2);
  return (std::pow(2.0, 2.0 / 3.0) * CF) * jpow(r, 5.0 / 3.0) * F;
}
J llp_spin(const J& r, const J& g) {
  const J x = jsqrt(g) / jpow(r, 4.0 / 3.0);
  return (std::pow(2.0, 2.0 / 3.0) * CF) * jpow(r, 5.0 / 3.0) * (1.0 + 0.0044188 * x * x / (1.0 + 0.0253 * x * jasinh(x)));
}

bool basic(int id, const J& a, const J& b, const J& gaa, const J& gab, const J& gbb, J* out) {
  switch (id) {
    case ORC_NONE: *out = cst(0.0); return true;
    case ORC_X_SLATER: *out = slaterx(a, b); return true;
    case ORC_C_VWN: *out = vwn5c(a, b); return true;
    case ORC_K_TF: *out = tfk(a, b); return true;
    case ORC_X_B88: *out = slaterx(a, b) + b88_spin(a, gaa) + b88_spin(b, gbb); return true;
    case ORC_X_B88_CORR: *out = b88_spin(a, gaa) + b88_spin(b, gbb); return true;
    case ORC_X_PBE: *out = pbex(a, b, gaa, gbb); return true;
    case ORC_C_LYP: *out = lypc(a, b, gaa, gab, gbb); return true;
    case ORC_C_P86: *out = p86c(a, b, gaa, gab, gbb); return true;
    case ORC_C_PBE: *out = pbec(a, b, gaa, gab, gbb); return true;
    case ORC_K_PW91: *out = lc94_spin(a, gaa) + lc94_spin(b, gbb); return true;
    case ORC_K_LLP: *out = llp_spin(a, gaa) + llp_spin(b, gbb); return true;
    default: return false;
  }
}

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
