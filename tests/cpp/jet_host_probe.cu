// jet_host_probe.cu - host-side probe of the DEVICE functional source (serenity_b200/csrc/functionals.cuh, kernel2.cuh):
// the energy expressions and the first/second-order jets are __host__ __device__, so their arithmetic can be checked against
// the oracle on a machine without a GPU (tests/test_kernel_sigma.py).  Test infrastructure, not part of the product library.
#include "../../serenity_b200/csrc/kernel2.cuh"

using namespace sxc;

extern "C" {

// F, d5, packed upper-triangle Hessian h15 w.r.t. (rho_a, rho_b, s_aa, s_ab, s_bb) from Jet2<5>
int jet_probe_u(int id, double ra, double rb, double gaa, double gab, double gbb, double* F, double* d5, double* h15) {
  typedef Jet2<5> T;
  const T a = jet_var<5>(ra, 0), b = jet_var<5>(rb, 1), saa = jet_var<5>(gaa, 2), sab = jet_var<5>(gab, 3), sbb = jet_var<5>(gbb, 4);
  const T e = basic_functional<T>(id, a, b, saa, sab, sbb);
  *F = e.v;
  for (int i = 0; i < 5; ++i) d5[i] = e.d[i];
  for (int i = 0; i < 15; ++i) h15[i] = e.h[i];
  return functional_id_supported(id) ? 0 : -1;
}

// closed shell, the seeding of k_kernel2_r: out = F, F_n, F_sigma, F_nn, F_nsigma, F_sigmasigma
int jet_probe_r(int id, double rho, double sigma, double* out6) {
  typedef Jet2<2> T;
  T a = jet_var<2>(0.5 * rho, 0), g4 = jet_var<2>(0.25 * sigma, 1);
  a.d[0] = 0.5;
  g4.d[1] = 0.25;
  const T e = basic_functional<T>(id, a, a, g4, g4, g4);
  out6[0] = e.v;
  out6[1] = e.d[0];
  out6[2] = e.d[1];
  out6[3] = e.h[0];
  out6[4] = e.h[1];
  out6[5] = e.h[2];
  return 0;
}

// first-order device jets (k_functional_u seeding): F, d5
int dual_probe_u(int id, double ra, double rb, double gaa, double gab, double gbb, double* F, double* d5) {
  typedef Dual<5> T;
  T a = mk<5>(ra), b = mk<5>(rb), saa = mk<5>(gaa), sab = mk<5>(gab), sbb = mk<5>(gbb);
  a.d[0] = b.d[1] = saa.d[2] = sab.d[3] = sbb.d[4] = 1.0;
  const T e = basic_functional<T>(id, a, b, saa, sab, sbb);
  *F = e.v;
  for (int i = 0; i < 5; ++i) d5[i] = e.d[i];
  return 0;
}
}
