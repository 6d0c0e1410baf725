// functionals.cuh - LDA/GGA exchange-correlation and kinetic-energy kernels in place of xcfun_eval.
//
// Row 8a-4 of SURVEY.md: XCFun::calcData (src/dft/functionals/wrappers/XCFun.cpp:39-159) evaluates, per grid
// point, the composite F = sum_i c_i f_i and its first derivatives through the third-party library XCFun, which
// differentiates templated energy expressions by forward-mode AD.  The device code does the same in registers:
// the published spin-resolved energy densities f(rho_a, rho_b, s_aa, s_ab, s_bb) (XCFun parametrisation,
// SURVEY.md Appendix A) are written once on a small dual-number type; RESTRICTED seeds two directions
// (d/d rho, d/d sigma at rho_a = rho_b = rho/2, s_xx = sigma/4).  Output convention of XC_N_NX_NY_NZ
// (XCFun.cpp:280-288): F, dF/drho, dF/d(grad rho) = 2 dF/dsigma grad rho.
// Bandwidth-class kernel: 32 B in, 32 B out per point, coalesced; block sums by warp shuffles.
#pragma once

#include "sxc_common.cuh"

namespace sxc {

// ---------------------------------------------------------------------------------------- dual numbers
template <int N>
struct Dual {
  double v;
  double d[N];
};

#define SXC_HD __host__ __device__ __forceinline__

template <int N> SXC_HD Dual<N> mk(double v) {
  Dual<N> r;
  r.v = v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = 0.0;
  return r;
}
// r = f(a) with f'(a) = fp
template <int N> SXC_HD Dual<N> chain(const Dual<N>& a, double f, double fp) {
  Dual<N> r;
  r.v = f;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = fp * a.d[i];
  return r;
}
template <int N> SXC_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N> SXC_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N> SXC_HD Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N> SXC_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N> SXC_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  const double inv = 1.0 / b.v;
  r.v = a.v * inv;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int N> SXC_HD Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> SXC_HD Dual<N> operator+(double b, const Dual<N>& a) { Dual<N> r = a; r.v += b; return r; }
template <int N> SXC_HD Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> SXC_HD Dual<N> operator-(double b, const Dual<N>& a) { Dual<N> r = -a; r.v += b; return r; }
template <int N> SXC_HD Dual<N> operator*(const Dual<N>& a, double b) { return chain(a, a.v * b, b); }
template <int N> SXC_HD Dual<N> operator*(double b, const Dual<N>& a) { return chain(a, a.v * b, b); }
template <int N> SXC_HD Dual<N> operator/(const Dual<N>& a, double b) { return chain(a, a.v / b, 1.0 / b); }
template <int N> SXC_HD Dual<N> operator/(double b, const Dual<N>& a) {
  const double inv = 1.0 / a.v;
  return chain(a, b * inv, -b * inv * inv);
}
template <int N> SXC_HD Dual<N> dsqrt(const Dual<N>& a) {
  const double s = sqrt(a.v);
  return chain(a, s, a.v > 0.0 ? 0.5 / s : 0.0);  // d sqrt at 0: the factor it multiplies (grad rho) vanishes too
}
template <int N> SXC_HD Dual<N> dcbrt(const Dual<N>& a) {
  const double c = cbrt(a.v);
  return chain(a, c, c / (3.0 * a.v));
}
template <int N> SXC_HD Dual<N> dexp(const Dual<N>& a) {
  const double e = exp(a.v);
  return chain(a, e, e);
}
template <int N> SXC_HD Dual<N> dexpm1(const Dual<N>& a) { return chain(a, expm1(a.v), exp(a.v)); }
template <int N> SXC_HD Dual<N> dlog(const Dual<N>& a) { return chain(a, log(a.v), 1.0 / a.v); }
template <int N> SXC_HD Dual<N> dlog1p(const Dual<N>& a) { return chain(a, log1p(a.v), 1.0 / (1.0 + a.v)); }
template <int N> SXC_HD Dual<N> datan(const Dual<N>& a) { return chain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
template <int N> SXC_HD Dual<N> dasinh(const Dual<N>& a) {
  return chain(a, asinh(a.v), 1.0 / sqrt(1.0 + a.v * a.v));
}
// a^(4/3), a^(5/3), a^(8/3) through one cbrt
template <int N> SXC_HD Dual<N> pow43(const Dual<N>& a) { return a * dcbrt(a); }
template <int N> SXC_HD Dual<N> pow53(const Dual<N>& a) { const Dual<N> c = dcbrt(a); return a * c * c; }

// ---------------------------------------------------------------------------------------- constants
namespace fc {
constexpr double PI = 3.14159265358979323846;
constexpr double CBRT_3_OVER_PI = 0.98474502184269654115;      // (3/pi)^(1/3)
constexpr double CBRT_6_OVER_PI = 1.2407009817988000333;       // (6/pi)^(1/3)
constexpr double CF = 2.8712340001881918160;                   // (3/10)(3 pi^2)^(2/3)
constexpr double CBRT_3PI2 = 3.0936677262801359310;            // (3 pi^2)^(1/3)
constexpr double CBRT_6PI2 = 3.8977770897207539590;            // (6 pi^2)^(1/3)
constexpr double TWO_23 = 1.5874010519681994748;               // 2^(2/3)
constexpr double TWO_13 = 1.2599210498948731648;               // 2^(1/3)
constexpr double TWO_43 = 2.5198420997897463295;               // 2^(4/3)
constexpr double TWO_113 = 12.699208415745595798;              // 2^(11/3)
constexpr double FPP0 = 1.7099209341613656176;                 // f''(0) = 4/(9 (2^(1/3) - 1))
constexpr double PBE_KAPPA = 0.804;
constexpr double PBE_MU = 0.2195149727645171;
constexpr double PBE_BETA = 0.06672455060314922;
constexpr double PBE_GAMMA = 0.031090690869654895035;          // (1 - ln 2)/pi^2
constexpr double B88_BETA = 0.0042;
}  // namespace fc

// ---------------------------------------------------------------------------------------- energy expressions
// f(zeta) = [(1+z)^(4/3) + (1-z)^(4/3) - 2]/(2^(4/3) - 2) from zp = 1 + zeta = 2 rho_a / n and zm = 1 - zeta = 2 rho_b / n:
// forming 1 -+ zeta from the spin densities keeps full precision at (nearly) fully polarised points
template <class T> SXC_HD T f_zeta(const T& zp, const T& zm) {
  return (pow43(zp) + pow43(zm) - 2.0) / (fc::TWO_43 - 2.0);
}

template <class T> SXC_HD T e_slaterx(const T& a, const T& b) {
  return (pow43(a) + pow43(b)) * (-0.75 * fc::CBRT_6_OVER_PI);
}

template <class T> SXC_HD T vwn_eps(const T& x, double A, double x0, double b, double c) {
  const double Q = sqrt(4.0 * c - b * b);
  const T X = x * x + b * x + c;
  const double X0 = x0 * x0 + b * x0 + c;
  const T at = datan(Q / (2.0 * x + b));
  const T xm = x - x0;
  return A * (dlog(x * x / X) + (2.0 * b / Q) * at - (b * x0 / X0) * (dlog(xm * xm / X) + (2.0 * (b + 2.0 * x0) / Q) * at));
}

template <class T> SXC_HD T e_vwn5c(const T& a, const T& b) {
  const T n = a + b;
  const T z = (a - b) / n;
  const T rs = dcbrt(3.0 / (4.0 * fc::PI * n));
  const T x = dsqrt(rs);
  const T eP = vwn_eps(x, 0.0310907, -0.10498, 3.72744, 12.9352);
  const T eF = vwn_eps(x, 0.01554535, -0.32500, 7.06042, 18.0578);
  const T ac = vwn_eps(x, -1.0 / (6.0 * fc::PI * fc::PI), -0.0047584, 1.13107, 13.0045);
  const T fz = f_zeta(2.0 * a / n, 2.0 * b / n);
  const T z2 = z * z, z4 = z2 * z2;
  return n * (eP + ac * fz * (1.0 - z4) / fc::FPP0 + (eF - eP) * fz * z4);
}

template <class T> SXC_HD T e_tfk(const T& a, const T& b) { return (pow53(a) + pow53(b)) * (fc::TWO_23 * fc::CF); }

template <class T> SXC_HD T pbex_cs(const T& n, const T& g) {  // E_x of a closed-shell density n, |grad n|^2 = g
  const T n43 = pow43(n);
  const T s2 = g / (4.0 * fc::CBRT_3PI2 * fc::CBRT_3PI2 * n43 * n43);
  const T Fx = (1.0 + fc::PBE_KAPPA) - fc::PBE_KAPPA / (1.0 + (fc::PBE_MU / fc::PBE_KAPPA) * s2);
  return n43 * Fx * (-0.75 * fc::CBRT_3_OVER_PI);
}
template <class T> SXC_HD T e_pbex(const T& a, const T& b, const T& gaa, const T& gbb) {
  return 0.5 * (pbex_cs(2.0 * a, 4.0 * gaa) + pbex_cs(2.0 * b, 4.0 * gbb));
}

template <class T> SXC_HD T b88_spin(const T& r, const T& g) {
  const T r43 = pow43(r);
  const T x = dsqrt(g) / r43;
  return -fc::B88_BETA * r43 * x * x / (1.0 + 6.0 * fc::B88_BETA * x * dasinh(x));
}
template <class T> SXC_HD T e_beckecorrx(const T& a, const T& b, const T& gaa, const T& gbb) {
  return b88_spin(a, gaa) + b88_spin(b, gbb);
}

template <class T> SXC_HD T e_lypc(const T& a, const T& b, const T& gaa, const T& gab, const T& gbb) {
  constexpr double A = 0.04918, B = 0.132, C = 0.2533, D = 0.349;
  const T n = a + b;
  const T g = gaa + 2.0 * gab + gbb;
  const T q = 1.0 / dcbrt(n);  // n^(-1/3)
  const T n2 = n * n;
  const T q2 = q * q;
  const T nm113 = q2 / (n2 * n);  // n^(-11/3)
  const T den = 1.0 + D * q;
  const T omega = dexp(-C * q) * nm113 / den;
  const T delta = C * q + D * q / den;
  const T ca = dcbrt(a), cb = dcbrt(b);
  const T a83 = a * a * ca * ca, b83 = b * b * cb * cb;
  const T t = a * b * (fc::TWO_113 * fc::CF * (a83 + b83) + (47.0 / 18.0 - 7.0 / 18.0 * delta) * g -
                       (2.5 - delta / 18.0) * (gaa + gbb) - (delta - 11.0) / 9.0 * (a * gaa + b * gbb) / n) -
              (2.0 / 3.0) * n2 * g + ((2.0 / 3.0) * n2 - a * a) * gbb + ((2.0 / 3.0) * n2 - b * b) * gaa;
  return -A * 4.0 * a * b / (den * n) - (A * B) * omega * t;
}

template <class T> SXC_HD T pw92_G(const T& rs, const T& srs, double A, double a1, double b1, double b2, double b3,
                                    double b4) {
  const T q1 = (2.0 * A) * (b1 * srs + b2 * rs + b3 * rs * srs + b4 * rs * rs);
  return (-2.0 * A) * (1.0 + a1 * rs) * dlog1p(1.0 / q1);
}
template <class T> SXC_HD T pw92_eps(const T& rs, const T& z, const T& zp, const T& zm) {
  const T srs = dsqrt(rs);
  const T e0 = pw92_G(rs, srs, 0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294);
  const T e1 = pw92_G(rs, srs, 0.01554535, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517);
  const T mac = pw92_G(rs, srs, 0.0168869, 0.11125, 10.357, 3.6231, 0.88026, 0.49671);  // -alpha_c
  const T fz = f_zeta(zp, zm);
  const T z2 = z * z, z4 = z2 * z2;
  return e0 - mac * fz * (1.0 - z4) / fc::FPP0 + (e1 - e0) * fz * z4;
}

template <class T> SXC_HD T e_pbec(const T& a, const T& b, const T& gaa, const T& gab, const T& gbb) {
  const T n = a + b;
  const T g = gaa + 2.0 * gab + gbb;
  const T z = (a - b) / n;
  const T rs = dcbrt(3.0 / (4.0 * fc::PI * n));
  const T zp = 2.0 * a / n, zm = 2.0 * b / n;
  const T eps = pw92_eps(rs, z, zp, zm);
  const T cp = dcbrt(zp), cm = dcbrt(zm);
  const T phi = 0.5 * (cp * cp + cm * cm);
  const T phi3 = phi * phi * phi;
  const T kF = fc::CBRT_3PI2 * dcbrt(n);
  const T t2 = g * (fc::PI / 16.0) / (phi * phi * kF * n * n);  // g / (4 phi^2 ks^2 n^2), ks^2 = 4 kF/pi
  const T Aa = (fc::PBE_BETA / fc::PBE_GAMMA) / dexpm1(-eps / (fc::PBE_GAMMA * phi3));
  const T At2 = Aa * t2;
  const T H = fc::PBE_GAMMA * phi3 *
              dlog1p((fc::PBE_BETA / fc::PBE_GAMMA) * t2 * (1.0 + At2) / (1.0 + At2 + At2 * At2));
  return n * (eps + H);
}

template <class T> SXC_HD T pz81_branch(const T& rs, double g, double b1, double b2, double A, double B, double C,
                                         double D) {
  if (rs.v >= 1.0) return g / (1.0 + b1 * dsqrt(rs) + b2 * rs);
  const T lr = dlog(rs);
  return A * lr + B + C * rs * lr + D * rs;
}
template <class T> SXC_HD T e_p86c(const T& a, const T& b, const T& gaa, const T& gab, const T& gbb) {
  const T n = a + b;
  const T g = gaa + 2.0 * gab + gbb;
  const T z = (a - b) / n;
  const T rs = dcbrt(3.0 / (4.0 * fc::PI * n));
  const T eU = pz81_branch(rs, -0.1423, 1.0529, 0.3334, 0.0311, -0.048, 0.0020, -0.0116);
  const T eP = pz81_branch(rs, -0.0843, 1.3981, 0.2611, 0.01555, -0.0269, 0.0007, -0.0048);
  const T eps = eU + f_zeta(2.0 * a / n, 2.0 * b / n) * (eP - eU);
  const T rs2 = rs * rs;
  const T Cn = 0.001667 + (0.002568 + 0.023266 * rs + 7.389e-6 * rs2) / (1.0 + 8.723 * rs + 0.472 * rs2 + 0.07389 * rs2 * rs);
  const T c = dcbrt(n);
  const T n43 = n * c;
  const T n76 = n * dsqrt(c);  // n^(7/6)
  const T Phi = (1.7454151061251240 /* (9 pi)^(1/6), pinned by FuncPotential_test.cpp:148-187 */ * 0.11 * 0.004235) * dsqrt(g) / (Cn * n76);
  const T hp = a / n, hm = b / n;
  const T d = fc::TWO_13 * dsqrt(pow53(hp) + pow53(hm));
  return n * eps + dexp(-Phi) * Cn * g / (d * n43);
}

template <class T> SXC_HD T lc94_spin(const T& r, const T& g) {
  constexpr double a1 = 0.093907, a2 = 76.320, a3 = 0.26608, a4 = 0.0809615, aa = 100.0, bb = 0.57767e-4;
  const T r43 = pow43(r);
  const T s = dsqrt(g) / ((2.0 * fc::CBRT_6PI2) * r43);
  const T s2 = s * s;
  const T L = a1 * s * dasinh(a2 * s);
  const T F = (1.0 + L + (a3 - a4 * dexp(-aa * s2)) * s2) / (1.0 + L + bb * s2 * s2);
  return (fc::TWO_23 * fc::CF) * pow53(r) * F;
}
template <class T> SXC_HD T llp_spin(const T& r, const T& g) {
  const T x = dsqrt(g) / pow43(r);
  return (fc::TWO_23 * fc::CF) * pow53(r) * (1.0 + 0.0044188 * x * x / (1.0 + 0.0253 * x * dasinh(x)));
}

// one basic functional by BASIC_FUNCTIONALS id (src/dft/functionals/BasicFunctionals.h:39-...)
template <class T>
__host__ __device__ T basic_functional(int id, const T& a, const T& b, const T& gaa, const T& gab, const T& gbb) {
  switch (id) {
    case 2: return e_slaterx(a, b);
    case 45: return e_vwn5c(a, b);
    case 66: return e_tfk(a, b);
    case 80: return e_slaterx(a, b) + e_beckecorrx(a, b, gaa, gbb);
    case 81: return e_beckecorrx(a, b, gaa, gbb);
    case 135: return e_pbex(a, b, gaa, gbb);
    case 184: return e_lypc(a, b, gaa, gab, gbb);
    case 193: return e_p86c(a, b, gaa, gab, gbb);
    case 197: return e_pbec(a, b, gaa, gab, gbb);
    case 283: return lc94_spin(a, gaa) + lc94_spin(b, gbb);
    case 286: return llp_spin(a, gaa) + llp_spin(b, gbb);
    default: {
      T zero = a;
      zero = zero - zero;
      return zero;
    }
  }
}

constexpr int MAX_COMP = 8;
struct FuncView {
  int ncomp;
  int gga;
  int id[MAX_COMP];
  double mix[MAX_COMP];
};

__host__ __device__ inline bool functional_id_is_gga(int id) {
  return id == 80 || id == 81 || id == 135 || id == 184 || id == 193 || id == 197 || id == 283 || id == 286;
}
__host__ inline bool functional_id_supported(int id) {
  return id == 0 || id == 2 || id == 45 || id == 66 || functional_id_is_gga(id);
}

// ------------------------------------------------------------------------------------------------------------
// K3: functional on literal 128-point blocks (XCFun.cpp:129-153).  One CTA (128 threads) per block.
//   out_v[0..3] (+)= sign * {dF/drho, dF/dgx, dF/dgy, dF/dgz};  e_part[lb] = sum_p w F;  n_part[lb] = sum_p w rho
// lit_blocks == nullptr: block index = blockIdx.x.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void functional_block(const FuncView& f, long npts, const int* __restrict__ lit_blocks,
                                                 const double* __restrict__ w, const double* __restrict__ rho,
                                                 const double* __restrict__ gx, const double* __restrict__ gy,
                                                 const double* __restrict__ gz, double sign, int accumulate,
                                                 double* __restrict__ epuv, double* __restrict__ v_rho, double* __restrict__ v_gx,
                                                 double* __restrict__ v_gy, double* __restrict__ v_gz, double* __restrict__ e_part,
                                                 double* __restrict__ n_part) {
  __shared__ double scratch[32];
  const int lb = lit_blocks ? lit_blocks[blockIdx.x] : blockIdx.x;
  const long first = (long)lb * FUNC_BLOCK;
  const int n = (int)min((long)FUNC_BLOCK, npts - first);
  const int t = threadIdx.x;
  const bool valid = t < n;
  const long p = first + t;
  const double r = valid ? rho[p] : 0.0;
  const double wp = valid ? w[p] : 0.0;
  const double sum_abs = block_sum(fabs(r), scratch);
  const bool skip = sum_abs < (double)n * 1e-12;  // XCFun.cpp:135-140
  double F = 0.0, vr = 0.0, vgx = 0.0, vgy = 0.0, vgz = 0.0;
  if (!skip && valid && !(r < 1e-14)) {  // xcfun returns zeros below its tiny density (LibXC.cpp:85-86)
    double x = 0.0, y = 0.0, z = 0.0;
    if (f.gga) {
      x = gx[p];
      y = gy[p];
      z = gz[p];
    }
    const double sigma = x * x + y * y + z * z;
    typedef Dual<2> T;
    T a, g4;
    a.v = 0.5 * r;
    a.d[0] = 0.5;
    a.d[1] = 0.0;
    g4.v = 0.25 * sigma;
    g4.d[0] = 0.0;
    g4.d[1] = 0.25;
    double vs = 0.0;
    for (int c = 0; c < f.ncomp; ++c) {
      const T e = basic_functional<T>(f.id[c], a, a, g4, g4, g4);
      F += f.mix[c] * e.v;
      vr += f.mix[c] * e.d[0];
      vs += f.mix[c] * e.d[1];
    }
    vgx = 2.0 * vs * x;
    vgy = 2.0 * vs * y;
    vgz = 2.0 * vs * z;
  }
  if (valid) {
    if (epuv) epuv[p] = F;
    if (accumulate) {
      v_rho[p] += sign * vr;
      if (v_gx) {
        v_gx[p] += sign * vgx;
        v_gy[p] += sign * vgy;
        v_gz[p] += sign * vgz;
      }
    } else {
      v_rho[p] = sign * vr;
      if (v_gx) {
        v_gx[p] = sign * vgx;
        v_gy[p] = sign * vgy;
        v_gz[p] = sign * vgz;
      }
    }
  }
  const double e = block_sum(wp * F, scratch);
  const double ne = block_sum(wp * r, scratch);
  if (t == 0) {
    if (e_part) e_part[lb] = e;
    if (n_part) n_part[lb] = ne;
  }
}

__global__ void __launch_bounds__(FUNC_BLOCK)
k_functional(FuncView f, long npts, const int* __restrict__ lit_blocks, const double* __restrict__ w,
             const double* __restrict__ rho, const double* __restrict__ gx, const double* __restrict__ gy,
             const double* __restrict__ gz, double sign, int accumulate, double* __restrict__ epuv,
             double* __restrict__ v_rho, double* __restrict__ v_gx, double* __restrict__ v_gy,
             double* __restrict__ v_gz, double* __restrict__ e_part, double* __restrict__ n_part) {
  functional_block(f, npts, lit_blocks, w, rho, gx, gy, gz, sign, accumulate, epuv, v_rho, v_gx, v_gy, v_gz, e_part, n_part);
}

// The same kernel compiled for MINB resident CTAs per SM (128 / 96 registers instead of 168, a few spilled doubles): the dependent
// FP64 chains of a functional leave the pipe idle with 12 warps per SM; SXC_FUNC selects it (measured in profiles/README.md).
template <int MINB>
__global__ void __launch_bounds__(FUNC_BLOCK, MINB)
k_functional_occ(FuncView f, long npts, const int* __restrict__ lit_blocks, const double* __restrict__ w,
                 const double* __restrict__ rho, const double* __restrict__ gx, const double* __restrict__ gy,
                 const double* __restrict__ gz, double sign, int accumulate, double* __restrict__ epuv,
                 double* __restrict__ v_rho, double* __restrict__ v_gx, double* __restrict__ v_gy,
                 double* __restrict__ v_gz, double* __restrict__ e_part, double* __restrict__ n_part) {
  functional_block(f, npts, lit_blocks, w, rho, gx, gy, gz, sign, accumulate, epuv, v_rho, v_gx, v_gy, v_gz, e_part, n_part);
}

// ------------------------------------------------------------------------------------------------------------
// K3u: UNRESTRICTED functional on literal 128-point blocks (XCFun.cpp:100-112 vars XC_A_B_AX_AY_AZ_BX_BY_BZ, :129-153).
// dens8 / out8 rows: rho_a, gax, gay, gaz, rho_b, gbx, gby, gbz (each [npts]).  Five seeded directions
// (rho_a, rho_b, s_aa, s_ab, s_bb); dF/d(grad rho_a) = 2 v_aa grad rho_a + v_ab grad rho_b.  The block is skipped only
// if BOTH spin densities are below the block threshold (:133-137); output is zero when rho_a + rho_b < 1e-14 and a
// single channel below 1e-14 is raised to it (xcfun's density regularisation; "parity unpinned").
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FUNC_BLOCK)
k_functional_u(FuncView f, long npts, const int* __restrict__ lit_blocks, const double* __restrict__ w,
               const double* __restrict__ dens8, double sign, int accumulate, double* __restrict__ epuv,
               double* __restrict__ out8, double* __restrict__ e_part, double* __restrict__ n_part) {
  __shared__ double scratch[32];
  const int lb = lit_blocks ? lit_blocks[blockIdx.x] : blockIdx.x;
  const long first = (long)lb * FUNC_BLOCK;
  const int n = (int)min((long)FUNC_BLOCK, npts - first);
  const int t = threadIdx.x;
  const bool valid = t < n;
  const long p = first + t;
  const double ra = valid ? dens8[p] : 0.0;
  const double rb = valid ? dens8[4 * npts + p] : 0.0;
  const double wp = valid ? w[p] : 0.0;
  const double sum_a = block_sum(fabs(ra), scratch);
  const double sum_b = block_sum(fabs(rb), scratch);
  const bool skip = sum_a < (double)n * 1e-12 && sum_b < (double)n * 1e-12;
  double F = 0.0, o[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (!skip && valid && !(ra + rb < 1e-14)) {
    double ga[3] = {0.0, 0.0, 0.0}, gb[3] = {0.0, 0.0, 0.0};
    if (f.gga) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ga[c] = dens8[(size_t)(1 + c) * npts + p];
        gb[c] = dens8[(size_t)(5 + c) * npts + p];
      }
    }
    typedef Dual<5> T;
    T a = mk<5>(fmax(ra, 1e-14)), b = mk<5>(fmax(rb, 1e-14));
    T saa = mk<5>(ga[0] * ga[0] + ga[1] * ga[1] + ga[2] * ga[2]);
    T sab = mk<5>(ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2]);
    T sbb = mk<5>(gb[0] * gb[0] + gb[1] * gb[1] + gb[2] * gb[2]);
    a.d[0] = 1.0;
    b.d[1] = 1.0;
    saa.d[2] = 1.0;
    sab.d[3] = 1.0;
    sbb.d[4] = 1.0;
    double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < f.ncomp; ++c) {
      const T e = basic_functional<T>(f.id[c], a, b, saa, sab, sbb);
      F += f.mix[c] * e.v;
#pragma unroll
      for (int i = 0; i < 5; ++i) d[i] += f.mix[c] * e.d[i];
    }
    o[0] = d[0];
    o[4] = d[1];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[1 + c] = 2.0 * d[2] * ga[c] + d[3] * gb[c];
      o[5 + c] = 2.0 * d[4] * gb[c] + d[3] * ga[c];
    }
  }
  if (valid) {
    if (epuv) epuv[p] = F;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (!f.gga && (r & 3)) continue;
      double* dst = out8 + (size_t)r * npts + p;
      *dst = (accumulate ? *dst : 0.0) + sign * o[r];
    }
  }
  const double e = block_sum(wp * F, scratch);
  const double ne = block_sum(wp * (ra + rb), scratch);
  if (t == 0) {
    if (e_part) e_part[lb] = e;
    if (n_part) n_part[lb] = ne;
  }
}

// fixed-order tree reduction of per-block partial sums (deterministic E_xc and N_el); one CTA of 256 threads
__global__ void __launch_bounds__(256) k_reduce_partials(const double* __restrict__ part, int n, double scale,
                                                          int accumulate, double* __restrict__ out) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += part[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) *out = (accumulate ? *out : 0.0) + scale * s;
}

}  // namespace sxc
