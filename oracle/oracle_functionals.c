/*
 * oracle_functionals.c - closed-shell LDA/GGA kernels with hand-derived first derivatives
 * (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Row 8a-4 of SURVEY.md: the arithmetic the reference obtains from the un-vendored third-party library
 * XCFun (qcserenity/xcfun, default branch, no tag pinned - cmake/ImportXCFun.cmake:13-17) through
 * xcfun_eval (dft/functionals/wrappers/XCFun.cpp:148-150).  The published formulas are restated in XCFun's
 * parametrisation.  PARITY UNPINNED at the 1e-9 Eh level: the reference holds no
 * per-functional known-answer test; indirect anchors are listed in DESIGN.md.
 *
 * Conventions (dft/functionals/wrappers/XCFun.cpp:89-90, :280-288; LibXC.cpp:176-182):
 *   F        energy density per volume ("epuv")
 *   dF/drho  derivative w.r.t. the total density
 *   dF/dgrad = 2 dF/dsigma * grad rho,  sigma = |grad rho|^2     (XC_N_NX_NY_NZ output rows 2..4)
 * The device code (serenity_b200/csrc/functionals.cuh) derives the same quantities by forward-mode
 * automatic differentiation of the spin-resolved energy expressions - an independent derivation.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define TINY_DENSITY 1e-14 /* XCFUN_TINY_DENSITY; LibXC.cpp:85-86 sets the same "to match xcfun" */

/* ---------------------------------------------------------------- constants (one place, re-pinnable) */
static const double PBE_KAPPA = 0.804;
static const double PBE_MU = 0.2195149727645171;     /* beta*pi^2/3, beta = 0.06672455060314922 */
static const double PBE_BETA = 0.06672455060314922;
static const double B88_BETA = 0.0042;
static const double LYP_A = 0.04918, LYP_B = 0.132, LYP_C = 0.2533, LYP_D = 0.349;
/* VWN5 paramagnetic (Hartree) */
static const double VWN_A = 0.0310907, VWN_X0 = -0.10498, VWN_B = 3.72744, VWN_C = 12.9352;
/* PW92 eps_c(rs, zeta=0): A, alpha1, beta1..beta4 */
static const double PW92_A = 0.0310907, PW92_A1 = 0.21370, PW92_B1 = 7.5957, PW92_B2 = 3.5876, PW92_B3 = 1.6382,
                    PW92_B4 = 0.49294;
/* PZ81 (unpolarised) */
static const double PZ_GAMMA = -0.1423, PZ_BETA1 = 1.0529, PZ_BETA2 = 0.3334, PZ_A = 0.0311, PZ_B = -0.048,
                    PZ_C = 0.0020, PZ_D = -0.0116;
/* PW91k = Lembarki-Chermette LC94 */
static const double LC_A1 = 0.093907, LC_A2 = 76.320, LC_A3 = 0.26608, LC_A4 = 0.0809615, LC_A = 100.0,
                    LC_B = 0.57767e-4;
/* LLP91k */
static const double LLP_B = 0.0044188, LLP_C = 0.0253;

static double cx_lda(void) { return -0.75 * cbrt(3.0 / M_PI); }                      /* -(3/4)(3/pi)^(1/3) */
static double cf_tf(void) { return 0.3 * pow(3.0 * M_PI * M_PI, 2.0 / 3.0); }         /* (3/10)(3 pi^2)^(2/3) */
static double s2_pref(void) { return 1.0 / (4.0 * pow(3.0 * M_PI * M_PI, 2.0 / 3.0)); } /* s^2 = c sigma rho^(-8/3) */

/* ---------------------------------------------------------------- LDA pieces */
static void slater(double rho, double* F, double* vr) {
  const double r13 = cbrt(rho);
  *F = cx_lda() * rho * r13;
  *vr = (4.0 / 3.0) * cx_lda() * r13;
}

/* eps_c^VWN5(x), x = sqrt(rs), and d eps/dx */
static void vwn5_eps(double x, double* eps, double* deps_dx) {
  const double A = VWN_A, x0 = VWN_X0, b = VWN_B, c = VWN_C;
  const double Q = sqrt(4.0 * c - b * b);
  const double X = x * x + b * x + c, X0 = x0 * x0 + b * x0 + c;
  const double at = atan(Q / (2.0 * x + b));
  *eps = A * (log(x * x / X) + (2.0 * b / Q) * at -
              (b * x0 / X0) * (log((x - x0) * (x - x0) / X) + (2.0 * (b + 2.0 * x0) / Q) * at));
  const double den = Q * Q + (2.0 * x + b) * (2.0 * x + b);
  *deps_dx = A * (2.0 / x - (2.0 * x + b) / X - 4.0 * b / den -
                  (b * x0 / X0) * (2.0 / (x - x0) - (2.0 * x + b) / X - 4.0 * (2.0 * x0 + b) / den));
}

static void vwn5(double rho, double* F, double* vr) {
  const double rs = cbrt(3.0 / (4.0 * M_PI * rho));
  const double x = sqrt(rs);
  double e, de;
  vwn5_eps(x, &e, &de);
  *F = rho * e;
  *vr = e - (x / 6.0) * de; /* dx/drho = -x/(6 rho) */
}

static void pw92_eps(double rs, double* eps, double* deps_drs) {
  const double A = PW92_A, a1 = PW92_A1;
  const double sr = sqrt(rs);
  const double Q0 = -2.0 * A * (1.0 + a1 * rs);
  const double Q1 = 2.0 * A * (PW92_B1 * sr + PW92_B2 * rs + PW92_B3 * rs * sr + PW92_B4 * rs * rs);
  const double Q1p = A * (PW92_B1 / sr + 2.0 * PW92_B2 + 3.0 * PW92_B3 * sr + 4.0 * PW92_B4 * rs);
  const double lg = log1p(1.0 / Q1);
  *eps = Q0 * lg;
  *deps_drs = -2.0 * A * a1 * lg - Q0 * Q1p / (Q1 * Q1 + Q1);
}

static void pz81_eps(double rs, double* eps, double* deps_drs) {
  if (rs >= 1.0) {
    const double sr = sqrt(rs);
    const double den = 1.0 + PZ_BETA1 * sr + PZ_BETA2 * rs;
    *eps = PZ_GAMMA / den;
    *deps_drs = -PZ_GAMMA * (0.5 * PZ_BETA1 / sr + PZ_BETA2) / (den * den);
  } else {
    const double lr = log(rs);
    *eps = PZ_A * lr + PZ_B + PZ_C * rs * lr + PZ_D * rs;
    *deps_drs = PZ_A / rs + PZ_C * (lr + 1.0) + PZ_D;
  }
}

static void tfk(double rho, double* F, double* vr) {
  const double r23 = cbrt(rho) * cbrt(rho);
  *F = cf_tf() * rho * r23;
  *vr = (5.0 / 3.0) * cf_tf() * r23;
}

/* ---------------------------------------------------------------- GGA pieces */
static void pbex(double rho, double sigma, double* F, double* vr, double* vs) {
  const double r13 = cbrt(rho), r43 = rho * r13;
  const double c2 = s2_pref();
  const double u = c2 * sigma / (r43 * r43); /* s^2 */
  const double D = 1.0 + PBE_MU * u / PBE_KAPPA;
  const double Fx = 1.0 + PBE_KAPPA - PBE_KAPPA / D;
  const double dFdu = PBE_MU / (D * D);
  const double cx = cx_lda();
  *F = cx * r43 * Fx;
  *vr = cx * ((4.0 / 3.0) * r13 * Fx - (8.0 / 3.0) * r13 * u * dFdu);
  *vs = cx * dFdu * c2 / r43;
}

/* B88 gradient correction for one spin channel: f(ra, saa) = -beta ra^(4/3) x^2/(1 + 6 beta x asinh x) */
static void b88_spin(double ra, double saa, double* f, double* fra, double* fsaa) {
  const double beta = B88_BETA;
  const double r13 = cbrt(ra), r43 = ra * r13;
  const double x = sqrt(saa) / r43;
  const double as = asinh(x);
  const double D = 1.0 + 6.0 * beta * x * as;
  const double Dp = 6.0 * beta * (as + x / sqrt(1.0 + x * x));
  const double g = x * x / D;
  const double h = (2.0 * D - x * Dp) / (D * D); /* g'(x)/x */
  *f = -beta * r43 * g;
  *fra = -(4.0 / 3.0) * beta * r13 * (g - x * x * h);
  *fsaa = -beta * h / (2.0 * r43);
}

static void b88corr(double rho, double sigma, double* F, double* vr, double* vs) {
  double f, fra, fsaa;
  b88_spin(0.5 * rho, 0.25 * sigma, &f, &fra, &fsaa);
  *F = 2.0 * f;
  *vr = fra;
  *vs = 0.5 * fsaa;
}

static void lyp(double rho, double sigma, double* F, double* vr, double* vs) {
  const double a = LYP_A, b = LYP_B, c = LYP_C, d = LYP_D;
  const double CF = cf_tf();
  const double q = 1.0 / cbrt(rho);      /* rho^(-1/3) */
  const double qp = -q / (3.0 * rho);    /* dq/drho */
  const double D = 1.0 + d * q;
  const double E = exp(-c * q);
  const double delta = c * q + d * q / D;
  const double deltap = (c + d / (D * D)) * qp;
  const double T1 = -a * rho / D;
  const double T1p = -a / D - a * d * q / (3.0 * D * D);
  const double W = E / D;
  const double Wp = W * (q / (3.0 * rho)) * (c + d / D);
  const double r53 = q * q * q * q * q; /* rho^(-5/3) */
  const double G = CF * rho - r53 * sigma * (3.0 + 7.0 * delta) / 72.0;
  const double Gr = CF + (5.0 / 3.0) * (r53 / rho) * sigma * (3.0 + 7.0 * delta) / 72.0 - r53 * sigma * 7.0 * deltap / 72.0;
  const double Gs = -r53 * (3.0 + 7.0 * delta) / 72.0;
  *F = T1 - a * b * W * G;
  *vr = T1p - a * b * (Wp * G + W * Gr);
  *vs = -a * b * W * Gs;
}

static void pbec(double rho, double sigma, double* F, double* vr, double* vs) {
  const double gamma = (1.0 - log(2.0)) / (M_PI * M_PI);
  const double beta = PBE_BETA;
  const double rs = cbrt(3.0 / (4.0 * M_PI * rho));
  double eps, deps_drs;
  pw92_eps(rs, &eps, &deps_drs);
  const double epsp = deps_drs * (-rs / (3.0 * rho)); /* d eps/d rho */
  /* t^2 = sigma pi / (16 kF rho^2), kF = (3 pi^2 rho)^(1/3) */
  const double ct = M_PI / (16.0 * cbrt(3.0 * M_PI * M_PI));
  const double r13 = cbrt(rho);
  const double r73 = rho * rho * r13;
  const double y = ct * sigma / r73;
  const double em1 = expm1(-eps / gamma);
  const double A = (beta / gamma) / em1;
  const double dA_deps = A * A * (em1 + 1.0) / beta;
  const double N = 1.0 + A * y;
  const double Dn = 1.0 + A * y + A * A * y * y;
  const double Pq = (beta / gamma) * y * N / Dn;
  const double H = gamma * log1p(Pq);
  const double dH_dy = beta * (1.0 + 2.0 * A * y) / ((1.0 + Pq) * Dn * Dn);
  const double dH_dA = -beta * A * y * y * y * (2.0 + A * y) / ((1.0 + Pq) * Dn * Dn);
  *F = rho * (eps + H);
  *vr = eps + H + rho * (epsp + dH_dy * (-(7.0 / 3.0) * y / rho) + dH_dA * dA_deps * epsp);
  *vs = rho * dH_dy * ct / r73;
}

static void p86c(double rho, double sigma, double* F, double* vr, double* vs) {
  const double rs = cbrt(3.0 / (4.0 * M_PI * rho));
  double eps, deps_drs;
  pz81_eps(rs, &eps, &deps_drs);
  const double drs = -rs / (3.0 * rho);
  /* C(rho) */
  const double al = 0.023266, be = 7.389e-6, ga = 8.723, de = 0.472;
  const double num = 0.002568 + al * rs + be * rs * rs;
  const double den = 1.0 + ga * rs + de * rs * rs + 1.0e4 * be * rs * rs * rs;
  const double C = 0.001667 + num / den;
  const double dC_drs = ((al + 2.0 * be * rs) * den - num * (ga + 2.0 * de * rs + 3.0e4 * be * rs * rs)) / (den * den);
  const double Cp = dC_drs * drs;
  const double k = 1.7454151061251240 * 0.11 * 0.004235; /* (9 pi)^(1/6): the paper's rounded 1.745 misses FuncPotential_test.cpp:148-187 by 8e-6 */
  const double r16 = pow(rho, 1.0 / 6.0);
  const double r76 = rho * r16, r43 = rho * cbrt(rho);
  const double Phi = k * sqrt(sigma) / (C * r76);
  const double ex = exp(-Phi);
  const double T = ex * C * sigma / r43;
  const double dPhi_drho = Phi * (-Cp / C - 7.0 / (6.0 * rho));
  *F = rho * eps + T;
  *vr = eps + rho * deps_drs * drs + T * (-dPhi_drho + Cp / C - 4.0 / (3.0 * rho));
  *vs = ex * C / r43 * (1.0 - 0.5 * Phi);
}

/* PW91-like kinetic enhancement (Lembarki-Chermette) F(s) and F'(s)/s */
static void lc94_enh(double s, double* Fk, double* dFk_over_s) {
  const double s2 = s * s;
  const double as = asinh(LC_A2 * s);
  const double ex = exp(-LC_A * s2);
  const double L = LC_A1 * s * as;
  const double N = 1.0 + L + (LC_A3 - LC_A4 * ex) * s2;
  const double Dn = 1.0 + L + LC_B * s2 * s2;
  /* L'/s, finite for s -> 0 */
  const double Lp_s = (s > 1e-8 ? LC_A1 * as / s : LC_A1 * LC_A2) + LC_A1 * LC_A2 / sqrt(1.0 + LC_A2 * LC_A2 * s2);
  const double Np_s = Lp_s + 2.0 * (LC_A3 - LC_A4 * ex) + 2.0 * LC_A * LC_A4 * s2 * ex;
  const double Dp_s = Lp_s + 4.0 * LC_B * s2;
  *Fk = N / Dn;
  *dFk_over_s = (Np_s * Dn - N * Dp_s) / (Dn * Dn);
}

static void pw91k(double rho, double sigma, double* F, double* vr, double* vs) {
  const double CF = cf_tf();
  const double r13 = cbrt(rho), r23 = r13 * r13, r53 = rho * r23, r43 = rho * r13;
  const double c2 = s2_pref();
  const double s2 = c2 * sigma / (r43 * r43);
  const double s = sqrt(s2);
  double Fk, dFs;
  lc94_enh(s, &Fk, &dFs);
  *F = CF * r53 * Fk;
  /* dF/drho = CF[5/3 rho^(2/3) Fk + rho^(5/3) F' (-4/3 s/rho)], F' = dFs*s */
  *vr = CF * r23 * ((5.0 / 3.0) * Fk - (4.0 / 3.0) * dFs * s2);
  /* dF/dsigma = CF rho^(5/3) F' s/(2 sigma) = CF rho^(5/3) dFs c2 rho^(-8/3)/2 */
  *vs = 0.5 * CF * dFs * c2 / rho;
}

/* LLP91 kinetic: 2^(2/3) CF sum_s rho_s^(5/3) [1 + b x^2/(1 + c x asinh x)], x = |grad rho_s|/rho_s^(4/3) */
static void llp91k(double rho, double sigma, double* F, double* vr, double* vs) {
  const double CF = cf_tf() * cbrt(4.0); /* 2^(2/3) CF */
  const double ra = 0.5 * rho, saa = 0.25 * sigma;
  const double r13 = cbrt(ra), r23 = r13 * r13, r43 = ra * r13, r53 = ra * r23;
  const double x = sqrt(saa) / r43;
  const double as = asinh(x);
  const double D = 1.0 + LLP_C * x * as;
  const double Dp = LLP_C * (as + x / sqrt(1.0 + x * x));
  const double g = x * x / D;
  const double h = (2.0 * D - x * Dp) / (D * D); /* g'/x */
  const double f = CF * r53 * (1.0 + LLP_B * g);
  /* d/dra: 5/3 ra^(2/3)(1+b g) + ra^(5/3) b g' (-4/3 x/ra) */
  const double fra = CF * r23 * ((5.0 / 3.0) * (1.0 + LLP_B * g) - (4.0 / 3.0) * LLP_B * x * x * h);
  /* d/dsaa: ra^(5/3) b g' x/(2 saa) = ra^(5/3) b h /(2 ra^(8/3)) */
  const double fsaa = CF * LLP_B * h / (2.0 * ra);
  *F = 2.0 * f;
  *vr = fra;
  *vs = 0.5 * fsaa;
}

/* ---------------------------------------------------------------- dispatch */
static int is_gga_id(int id) {
  switch (id) {
    case ORC_X_B88:
    case ORC_X_B88_CORR:
    case ORC_X_PBE:
    case ORC_C_LYP:
    case ORC_C_P86:
    case ORC_C_PBE:
    case ORC_K_PW91:
    case ORC_K_LLP:
      return 1;
    default:
      return 0;
  }
}

int orc_functional_is_gga(const orc_functional* f) {
  for (int i = 0; i < f->ncomp; ++i)
    if (is_gga_id(f->id[i])) return 1;
  return 0;
}

int orc_basic_functional(int id, double rho, double sigma, double* F, double* vr, double* vs) {
  double f = 0, a = 0, s = 0, f2, a2;
  switch (id) {
    case ORC_NONE:
      break;
    case ORC_X_SLATER:
      slater(rho, &f, &a);
      break;
    case ORC_C_VWN:
      vwn5(rho, &f, &a);
      break;
    case ORC_K_TF:
      tfk(rho, &f, &a);
      break;
    case ORC_X_B88_CORR:
      b88corr(rho, sigma, &f, &a, &s);
      break;
    case ORC_X_B88: /* beckex = slaterx + beckecorrx */
      b88corr(rho, sigma, &f, &a, &s);
      slater(rho, &f2, &a2);
      f += f2;
      a += a2;
      break;
    case ORC_X_PBE:
      pbex(rho, sigma, &f, &a, &s);
      break;
    case ORC_C_LYP:
      lyp(rho, sigma, &f, &a, &s);
      break;
    case ORC_C_P86:
      p86c(rho, sigma, &f, &a, &s);
      break;
    case ORC_C_PBE:
      pbec(rho, sigma, &f, &a, &s);
      break;
    case ORC_K_PW91:
      pw91k(rho, sigma, &f, &a, &s);
      break;
    case ORC_K_LLP:
      llp91k(rho, sigma, &f, &a, &s);
      break;
    default:
      return -1;
  }
  *F = f;
  *vr = a;
  *vs = s;
  return 0;
}

/* 8a-4: XCFun::calcData, dft/functionals/wrappers/XCFun.cpp:39-159 (RESTRICTED, GRADIENTS) */
double orc_functional_on_grid(const orc_functional* fn, long npts, const double* w, const double* rho,
                              const double* gx, const double* gy, const double* gz, double* epuv, double* dFdRho,
                              double* dFdGx, double* dFdGy, double* dFdGz) {
  const int gga = orc_functional_is_gga(fn) && gx != NULL;
  const int BS = 128; /* the literal of FuncPotential.cpp:85 / NAddFuncPotential.cpp:196 */
  const long nblocks = (npts + BS - 1) / BS;
  memset(epuv, 0, sizeof(double) * (size_t)npts);
  memset(dFdRho, 0, sizeof(double) * (size_t)npts);
  if (dFdGx) {
    memset(dFdGx, 0, sizeof(double) * (size_t)npts);
    memset(dFdGy, 0, sizeof(double) * (size_t)npts);
    memset(dFdGz, 0, sizeof(double) * (size_t)npts);
  }
#pragma omp parallel for schedule(dynamic)
  for (long blk = 0; blk < nblocks; ++blk) { /* XCFun.cpp:129 */
    const long first = blk * BS;
    const long n = (blk == nblocks - 1) ? npts - first : BS;
    double sum = 0.0;
    for (long p = 0; p < n; ++p) sum += fabs(rho[first + p]);
    if (sum < (double)n * 1e-12) continue; /* XCFun.cpp:135-140 */
    for (long p = first; p < first + n; ++p) {
      const double r = rho[p];
      if (r < TINY_DENSITY) continue; /* xcfun_eval returns zeros below the tiny density */
      const double sigma = gga ? gx[p] * gx[p] + gy[p] * gy[p] + gz[p] * gz[p] : 0.0;
      double F = 0.0, vr = 0.0, vs = 0.0;
      for (int c = 0; c < fn->ncomp; ++c) {
        double f, a, s;
        orc_basic_functional(fn->id[c], r, sigma, &f, &a, &s);
        F += fn->mix[c] * f;
        vr += fn->mix[c] * a;
        vs += fn->mix[c] * s;
      }
      epuv[p] = F;
      dFdRho[p] = vr;
      if (gga && dFdGx) {
        dFdGx[p] = 2.0 * vs * gx[p];
        dFdGy[p] = 2.0 * vs * gy[p];
        dFdGz[p] = 2.0 * vs * gz[p];
      }
    }
  }
  /* calcEnergy, XCFun.cpp:752-765 */
  double energy = 0.0;
#pragma omp parallel for reduction(+ : energy)
  for (long p = 0; p < npts; ++p) energy += epuv[p] * w[p];
  return energy;
}
