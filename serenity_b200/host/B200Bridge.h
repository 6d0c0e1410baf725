// B200Bridge.h - the reference-side glue INTEGRATION.md section 3 binds through (to be dropped into src/potentials/ of Serenity).
//
// A process-wide owner of the device context(s) of libserenity_xc_b200 and of the handle caches that map Serenity's controller
// objects to the library's integer handles - the role RememberingFactory (src/misc/RememberingFactory.h:76) plays for the
// reference's own on-grid controllers.  Header only; it touches Serenity's classes through the accessors listed below and nothing
// else, so it is written against those NAMES (templates) and compiles both inside Serenity and against the stand-ins of
// tests/cpp/b200_bridge_test.cpp:
//
//   GridController   getGridPoints() -> Eigen::Matrix3Xd (3 x N column-major = xyz interleaved), getWeights() -> Eigen::VectorXd,
//                    getNGridPoints()                                              (src/grid/GridController.h:58-69)
//   BasisController  getBasis() -> std::vector<std::shared_ptr<Shell>>, getNBasisFunctions(), extendedIndex(i)
//                                                                                  (src/basis/BasisController.h:114-175)
//   Shell            getAngularMomentum(), isSpherical(), getNPrimitives(), getX/Y/Z(), getExponents(), getContractions(),
//                    getNormFactors()                                              (src/basis/Shell.h:88-185)
//   Functional       getBasicFunctionals(), getMixingFactors()                     (src/dft/Functional.h:76, :134)
//
// One GPU: every call goes to one sxc_ctx.  SERENITY_B200_GPUS = n > 1 (or B200Bridge::configure) creates an sxc_group instead:
// one context and one worker thread per GPU inside the library, the grid blocks sharded over them, one ncclAllReduce of
// [V | E | N] per build - the caller still makes ONE call from its SCF driver thread and receives the full matrix.
#pragma once

#include <serenity_xc_b200.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace Serenity {

class B200Bridge {
 public:
  // device / number of GPUs must be chosen before the first use (defaults: SERENITY_B200_DEVICE or 0, SERENITY_B200_GPUS or 1)
  static void configure(int firstDevice, int nGpus) {
    settings().first = firstDevice;
    settings().second = nGpus;
  }
  static B200Bridge& instance() {
    static B200Bridge bridge;
    return bridge;
  }
  B200Bridge(const B200Bridge&) = delete;
  B200Bridge& operator=(const B200Bridge&) = delete;
  ~B200Bridge() {
    if (_group) sxc_group_destroy(_group);
    if (_ctx) sxc_destroy(_ctx);
  }

  int nGpus() const { return _ngpu; }
  sxc_ctx* ctx() const { return _group ? sxc_group_ctx(_group, 0) : _ctx; }
  sxc_group* group() const { return _group; }

  // the reference throws SerenityError (src/misc/SerenityError.h:36); inside Serenity replace std::runtime_error by it
  void check(int rc) const {
    if (rc == SXC_OK) return;
    throw std::runtime_error(std::string("B200 XC library: ") + (_group ? sxc_group_last_error(_group) : sxc_last_error(_ctx)));
  }

  // ---- handle caches: keyed on the controller object; forget*() is what the potentials' notify() calls ---------------------
  template <class GridControllerT>
  int grid(GridControllerT& gc, unsigned blocksize) {
    std::lock_guard<std::mutex> lock(_mutex);
    auto it = _grids.find(&gc);
    if (it != _grids.end()) return it->second;
    const auto& pts = gc.getGridPoints();  // 3 x N, column-major
    const auto& w = gc.getWeights();
    int h = -1;
    check(_group ? sxc_group_set_grid(_group, (int64_t)gc.getNGridPoints(), pts.data(), w.data(), (int)blocksize, &h)
                 : sxc_set_grid(_ctx, (int64_t)gc.getNGridPoints(), pts.data(), w.data(), (int)blocksize, &h));
    return _grids[&gc] = h;
  }
  template <class GridControllerT>
  void forgetGrid(GridControllerT& gc) {
    std::lock_guard<std::mutex> lock(_mutex);
    auto it = _grids.find(&gc);
    if (it == _grids.end()) return;
    check(_group ? sxc_group_release_grid(_group, it->second) : sxc_release_grid(_ctx, it->second));
    _grids.erase(it);
  }

  template <class BasisControllerT>
  int basis(BasisControllerT& bc, double radialThreshold) {
    std::lock_guard<std::mutex> lock(_mutex);
    auto it = _bases.find(&bc);
    if (it != _bases.end()) return it->second;
    const auto& shells = bc.getBasis();
    const int ns = (int)shells.size();
    std::vector<int> l(ns), pure(ns), nprim(ns), first(ns);
    std::vector<double> centre(3 * (size_t)ns), alpha, coeff, normfac((size_t)bc.getNBasisFunctions(), 1.0);
    for (int i = 0; i < ns; ++i) {
      const auto& sh = *shells[i];
      l[i] = (int)sh.getAngularMomentum();
      pure[i] = sh.isSpherical() ? 1 : 0;
      nprim[i] = (int)sh.getNPrimitives();
      first[i] = (int)bc.extendedIndex(i);
      centre[3 * i] = sh.getX();
      centre[3 * i + 1] = sh.getY();
      centre[3 * i + 2] = sh.getZ();
      const auto ex = sh.getExponents();
      const auto co = sh.getContractions();  // libint-renormalised (src/basis/Shell.h:173-181)
      alpha.insert(alpha.end(), ex.begin(), ex.end());
      coeff.insert(coeff.end(), co.begin(), co.end());
      const auto& nf = sh.getNormFactors();  // Cartesian shells only (src/basis/Shell.cpp:37-47)
      for (int m = 0; m < (int)nf.size(); ++m) normfac[first[i] + m] = nf[m];
    }
    int h = -1;
    check(_group ? sxc_group_add_basis(_group, ns, l.data(), pure.data(), nprim.data(), first.data(), centre.data(), alpha.data(),
                                       coeff.data(), normfac.data(), radialThreshold, &h)
                 : sxc_add_basis(_ctx, ns, l.data(), pure.data(), nprim.data(), first.data(), centre.data(), alpha.data(),
                                 coeff.data(), normfac.data(), radialThreshold, &h));
    return _bases[&bc] = h;
  }
  template <class BasisControllerT>
  void forgetBasis(BasisControllerT& bc) {
    std::lock_guard<std::mutex> lock(_mutex);
    auto it = _bases.find(&bc);
    if (it == _bases.end()) return;
    check(_group ? sxc_group_release_basis(_group, it->second) : sxc_release_basis(_ctx, it->second));
    _bases.erase(it);
  }

  // a functional is identified by its components (src/dft/Functional.h:203-209 compares exactly these)
  template <class FunctionalT>
  int functional(const FunctionalT& f) {
    std::lock_guard<std::mutex> lock(_mutex);
    std::vector<int> ids;
    for (auto b : f.getBasicFunctionals()) ids.push_back((int)b);
    const std::vector<double> mix = f.getMixingFactors();
    auto key = std::make_pair(ids, mix);
    auto it = _funcs.find(key);
    if (it != _funcs.end()) return it->second;
    int h = -1;
    check(_group ? sxc_group_set_functional(_group, (int)ids.size(), ids.data(), mix.data(), &h)
                 : sxc_set_functional(_ctx, (int)ids.size(), ids.data(), mix.data(), &h));
    return _funcs[key] = h;
  }

  // ---- the calls of INTEGRATION.md section 3, routed to one context or to the group ---------------------------------------
  void buildXC(int grid, int basis, int func, int nspin, const double* P, double blockAveThreshold, double* V, double* E,
               double* nElectrons) {
    check(_group ? sxc_group_build_xc(_group, grid, basis, func, nspin, P, blockAveThreshold, V, E, nElectrons)
                 : sxc_build_xc(_ctx, grid, basis, func, nspin, P, blockAveThreshold, V, E, nElectrons));
  }
  // nfunc functionals on the same densities in one device pass (FDEPotentials::getFockMatrix adds naddXC + naddKin,
  // src/potentials/bundles/FDEPotentials.cpp:43-61); E: nfunc x (2 + nenv) energies, V: nfunc matrices or their sum
  void buildNAdd(int grid, int nfunc, const int* funcs, int nspin, int basisAct, const double* Pact, int nenv, const int* basisEnv,
                 const double* const* Penv, int envFrozen, double blockAveThreshold, int sumMatrices, double* V, double* E) {
    check(_group ? sxc_group_build_nadd_multi(_group, grid, nfunc, funcs, nspin, basisAct, Pact, nenv, basisEnv, Penv, envFrozen,
                                              blockAveThreshold, sumMatrices, V, E)
                 : sxc_build_nadd_multi(_ctx, grid, nfunc, funcs, nspin, basisAct, Pact, nenv, basisEnv, Penv, envFrozen,
                                        blockAveThreshold, sumMatrices, V, E));
  }
  void xcGradient(int grid, int basis, int func, int nspin, const double* P, int natoms, const int* atomOfBf, double* grad) {
    check(_group ? sxc_group_xc_gradient(_group, grid, basis, func, nspin, P, natoms, atomOfBf, grad)
                 : sxc_xc_gradient(_ctx, grid, basis, func, nspin, P, natoms, atomOfBf, grad));
  }

 private:
  static std::pair<int, int>& settings() {
    static std::pair<int, int> s = [] {
      const char* d = std::getenv("SERENITY_B200_DEVICE");
      const char* n = std::getenv("SERENITY_B200_GPUS");
      return std::make_pair(d ? std::atoi(d) : 0, n ? std::atoi(n) : 1);
    }();
    return s;
  }
  B200Bridge() {
    const int first = settings().first;
    _ngpu = settings().second < 1 ? 1 : settings().second;
    if (_ngpu > 1) {
      std::vector<int> dev(_ngpu);
      for (int i = 0; i < _ngpu; ++i) dev[i] = first + i;
      const int rc = sxc_group_create(&_group, _ngpu, dev.data());
      if (rc != SXC_OK) throw std::runtime_error("B200 XC library: sxc_group_create failed (no CPU fallback)");
    } else {
      const int rc = sxc_create(&_ctx, first);
      if (rc != SXC_OK) throw std::runtime_error("B200 XC library: sxc_create failed (no CPU fallback)");
    }
  }

  sxc_ctx* _ctx = nullptr;
  sxc_group* _group = nullptr;
  int _ngpu = 1;
  std::mutex _mutex;
  std::map<const void*, int> _grids, _bases;
  std::map<std::pair<std::vector<int>, std::vector<double>>, int> _funcs;
};

}  // namespace Serenity
