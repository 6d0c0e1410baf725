// b200_bridge_test.cpp - compiles serenity_b200/host/B200Bridge.h (the reference-side glue of INTEGRATION.md section 3) against
// stand-ins that expose exactly the accessors the bridge uses from Serenity's GridController / BasisController / Shell /
// Functional, and drives FuncPotential::getMatrix through it:
//   b200_bridge_test --compile-only          (CPU suite: the header compiles and links against the C ABI)
//   b200_bridge_test <in.bin> <out.bin> <ngpu>   (GPU suite: V_xc, E_xc of one build, through one context or an sxc_group)
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <vector>

#include "../../serenity_b200/host/B200Bridge.h"

namespace mock {
struct Matrix3Xd {  // Eigen::Matrix3Xd: 3 x N column-major
  std::vector<double> v;
  const double* data() const { return v.data(); }
};
struct VectorXd {
  std::vector<double> v;
  const double* data() const { return v.data(); }
  size_t size() const { return v.size(); }
  double operator[](size_t i) const { return v[i]; }
};
struct GridController {
  Matrix3Xd pts;
  VectorXd w;
  const Matrix3Xd& getGridPoints() { return pts; }
  const VectorXd& getWeights() { return w; }
  unsigned int getNGridPoints() { return (unsigned)w.v.size(); }
};
struct Shell {
  unsigned l;
  bool spherical;
  double x, y, z;
  std::vector<double> ex, co;
  VectorXd norm;
  unsigned int getAngularMomentum() const { return l; }
  bool isSpherical() const { return spherical; }
  unsigned int getNPrimitives() const { return (unsigned)ex.size(); }
  const double& getX() const { return x; }
  const double& getY() const { return y; }
  const double& getZ() const { return z; }
  std::vector<double> getExponents() const { return ex; }
  std::vector<double> getContractions() const { return co; }
  const VectorXd& getNormFactors() const { return norm; }
};
struct BasisController {
  std::vector<std::shared_ptr<Shell>> shells;
  std::vector<unsigned> first;
  unsigned nbf = 0;
  const std::vector<std::shared_ptr<Shell>>& getBasis() { return shells; }
  unsigned int getNBasisFunctions() { return nbf; }
  unsigned int extendedIndex(unsigned i) { return first[i]; }
};
struct Functional {
  std::vector<int> ids;
  std::vector<double> mix;
  const std::vector<int>& getBasicFunctionals() const { return ids; }
  std::vector<double> getMixingFactors() const { return mix; }
};
}  // namespace mock

template <class T>
static std::vector<T> rd(std::istream& in) {
  int64_t n = 0;
  in.read(reinterpret_cast<char*>(&n), 8);
  std::vector<T> v((size_t)n);
  in.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
  return v;
}
static void wr(std::ostream& out, const std::vector<double>& v) {
  const int64_t n = (int64_t)v.size();
  out.write(reinterpret_cast<const char*>(&n), 8);
  out.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(n * sizeof(double)));
}

int main(int argc, char** argv) {
  if (argc == 2 && std::strcmp(argv[1], "--compile-only") == 0) {
    std::cout << "B200Bridge.h compiled; C ABI version " << sxc_abi_version() << "\n";
    return 0;
  }
  if (argc != 4) return 2;
  try {
    std::ifstream in(argv[1], std::ios::binary);
    std::ofstream out(argv[2], std::ios::binary);
    Serenity::B200Bridge::configure(0, std::atoi(argv[3]));
    auto& b200 = Serenity::B200Bridge::instance();
    mock::GridController grid;
    grid.pts.v = rd<double>(in);
    grid.w.v = rd<double>(in);
    mock::Functional func{rd<int>(in), rd<double>(in)};
    // shell table as written by tests/test_cpp_host.py::_basis
    auto l = rd<int>(in), pure = rd<int>(in), nprim = rd<int>(in), first = rd<int>(in);
    auto centre = rd<double>(in), alpha = rd<double>(in), coeff = rd<double>(in), normfac = rd<double>(in);
    rd<int>(in);  // atom indices (unused here)
    mock::BasisController basis;
    size_t po = 0;
    int nbf = 0;
    for (size_t i = 0; i < l.size(); ++i) {
      auto sh = std::make_shared<mock::Shell>();
      sh->l = (unsigned)l[i];
      sh->spherical = pure[i] != 0;
      sh->x = centre[3 * i];
      sh->y = centre[3 * i + 1];
      sh->z = centre[3 * i + 2];
      sh->ex.assign(alpha.begin() + po, alpha.begin() + po + nprim[i]);
      sh->co.assign(coeff.begin() + po, coeff.begin() + po + nprim[i]);
      po += (size_t)nprim[i];
      const int nf = sh->spherical ? 2 * l[i] + 1 : (l[i] + 1) * (l[i] + 2) / 2;
      if (!sh->spherical) sh->norm.v.assign(normfac.begin() + first[i], normfac.begin() + first[i] + nf);
      basis.shells.push_back(sh);
      basis.first.push_back((unsigned)first[i]);
      nbf = std::max(nbf, first[i] + nf);
    }
    basis.nbf = (unsigned)nbf;
    std::vector<double> P = rd<double>(in);
    const int g = b200.grid(grid, 128);
    const int b = b200.basis(basis, 1e-9);
    const int f = b200.functional(func);
    if (b200.grid(grid, 128) != g || b200.basis(basis, 1e-9) != b || b200.functional(func) != f) return 4;  // cached handles
    std::vector<double> V((size_t)nbf * nbf), en(2);
    b200.buildXC(g, b, f, 1, P.data(), 1e-11, V.data(), &en[0], &en[1]);
    wr(out, V);
    wr(out, en);
    b200.forgetGrid(grid);  // a Grid notify(): the next use uploads again
    const int g2 = b200.grid(grid, 128);
    std::vector<double> V2((size_t)nbf * nbf), en2(2);
    b200.buildXC(g2, b, f, 1, P.data(), 1e-11, V2.data(), &en2[0], &en2[1]);
    wr(out, V2);
    wr(out, std::vector<double>{(double)b200.nGpus()});
    return 0;
  } catch (const std::exception& e) {
    std::cout << "error: " << e.what() << "\n";
    return 3;
  }
}
