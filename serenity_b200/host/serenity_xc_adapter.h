// serenity_xc_adapter.h - C++ host side above the C ABI: drop-in bodies for Serenity's FuncPotential and
// NAddFuncPotential with the reference's class names, constructor arguments, lazy-evaluation and error behaviour.
//
// Reference interfaces mirrored (paths relative to /root/reference/src):
//   Potential<SCFMode>            potentials/Potential.h:43-86        getMatrix / getEnergy / getGeomGradients
//   FuncPotential<SCFMode>        potentials/FuncPotential.h:47-147   (getMatrix: FuncPotential.cpp:74-111)
//   NAddFuncPotential<SCFMode>    potentials/NAddFuncPotential.h:118-155 (getMatrix: NAddFuncPotential.cpp:192-300)
//   NotifyingClass / ObjectSensitiveClass   notification/ (lazy invalidation: notify() drops the cached potential)
//   SerenityError                 misc/SerenityError.h:36
// Inside Serenity the light stand-ins below (Matrix, GridController, BasisController, DensityMatrixController,
// Functional) are the program's own classes; INTEGRATION.md shows the five call sites that change.  Header-only,
// C++14, no dependency besides include/serenity_xc_b200.h and the shared library.
#ifndef SERENITY_XC_ADAPTER_H
#define SERENITY_XC_ADAPTER_H

#include <algorithm>
#include <cctype>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/serenity_xc_b200.h"

namespace Serenity {

class SerenityError : public std::runtime_error {
 public:
  explicit SerenityError(const std::string& msg) : std::runtime_error(msg) {}
};

namespace Options {
enum class SCF_MODES { RESTRICTED = 0, UNRESTRICTED = 1 };
// settings/EmbeddingOptions.h:42-52
enum class KIN_EMBEDDING_MODES {
  NONE = 0, NADD_FUNC = 1, LEVELSHIFT = 2, HUZINAGA = 3, HOFFMANN = 4, RECONSTRUCTION = 5, FERMI_SHIFTED_HUZINAGA = 6, LOEWDIN = 7, ALMO = 8
};
}

// Column-major dense matrix with Eigen::MatrixXd's memory layout (data()[row + rows * col]).
struct Matrix {
  int nrows = 0, ncols = 0;
  std::vector<double> values;
  Matrix() = default;
  Matrix(int r, int c) : nrows(r), ncols(c), values((size_t)r * c, 0.0) {}
  int rows() const { return nrows; }
  int cols() const { return ncols; }
  double* data() { return values.data(); }
  const double* data() const { return values.data(); }
  double& operator()(int i, int j) { return values[(size_t)i + (size_t)nrows * j]; }
  double operator()(int i, int j) const { return values[(size_t)i + (size_t)nrows * j]; }
};
using FockMatrix = Matrix;     // data/matrices/FockMatrix.h
using DensityMatrix = Matrix;  // data/matrices/DensityMatrix.h (RESTRICTED: total density, occupation 2)

// notification/ObjectSensitiveClass.h, NotifyingClass.h
class ObjectSensitive {
 public:
  virtual ~ObjectSensitive() = default;
  virtual void notify() = 0;
};
class Notifying {
 public:
  void addSensitiveObject(std::weak_ptr<ObjectSensitive> o) { _sensitive.push_back(std::move(o)); }

 protected:
  void notifyObjects() {
    for (auto& w : _sensitive)
      if (auto s = w.lock()) s->notify();
  }

 private:
  std::vector<std::weak_ptr<ObjectSensitive>> _sensitive;
};

namespace B200 {

// The device side of the process, shared by all potentials: ONE GPU (sxc_create) or, for a single-process host like Serenity,
// N GPUs of the node as a group (sxc_group_create: one context + one host worker thread per GPU, grid blocks sharded, one
// ncclAllReduce of [V | E | N] inside the library per build).  The potentials only use the dispatchers below, so the same
// FuncPotential / NAddFuncPotential code runs on either.
class XCDevice {
 public:
  explicit XCDevice(int device = 0) {
    const int rc = sxc_create(&_ctx, device);
    if (rc != SXC_OK)
      throw SerenityError("serenity_xc_b200: no usable CUDA device " + std::to_string(device) + " (status " +
                          std::to_string(rc) + "); the XC build has no CPU fallback");
  }
  explicit XCDevice(const std::vector<int>& devices) {
    if (devices.size() == 1) {
      const int rc = sxc_create(&_ctx, devices[0]);
      if (rc != SXC_OK) throw SerenityError("serenity_xc_b200: no usable CUDA device (status " + std::to_string(rc) + ")");
      return;
    }
    const int rc = sxc_group_create(&_group, (int)devices.size(), devices.data());
    if (rc != SXC_OK)
      throw SerenityError("serenity_xc_b200: could not create a group of " + std::to_string(devices.size()) +
                          " GPUs (status " + std::to_string(rc) + "); the XC build has no CPU fallback");
  }
  ~XCDevice() {
    if (_group) sxc_group_destroy(_group);
    if (_ctx) sxc_destroy(_ctx);
  }
  XCDevice(const XCDevice&) = delete;
  XCDevice& operator=(const XCDevice&) = delete;
  int nGPUs() const { return _group ? sxc_group_size(_group) : 1; }
  // the single context, for the classes that are not sharded through a group (stage-level classes, AB potentials, LR-TDDFT kernel)
  sxc_ctx* get() const {
    if (_group) throw SerenityError("serenity_xc_b200: this class runs on a single-GPU XCDevice (the group shards FuncPotential and "
                                    "NAddFuncPotential builds only)");
    return _ctx;
  }
  void check(int rc) const {
    if (rc != SXC_OK)
      throw SerenityError(std::string("serenity_xc_b200: ") + (_group ? sxc_group_last_error(_group) : sxc_last_error(_ctx)));
  }
  // ---- dispatchers (one GPU or the group)
  int setGrid(int64_t n, const double* xyz, const double* w, int blocksize) const {
    int h = -1;
    check(_group ? sxc_group_set_grid(_group, n, xyz, w, blocksize, &h) : sxc_set_grid(_ctx, n, xyz, w, blocksize, &h));
    return h;
  }
  int addBasis(int nshell, const int* l, const int* pure, const int* nprim, const int* firstBf, const double* centre,
               const double* alpha, const double* coeff, const double* normfac, double thr) const {
    int h = -1;
    check(_group ? sxc_group_add_basis(_group, nshell, l, pure, nprim, firstBf, centre, alpha, coeff, normfac, thr, &h)
                 : sxc_add_basis(_ctx, nshell, l, pure, nprim, firstBf, centre, alpha, coeff, normfac, thr, &h));
    return h;
  }
  int setFunctional(const std::vector<int>& ids, const std::vector<double>& mix) const {
    int h = -1;
    check(_group ? sxc_group_set_functional(_group, (int)ids.size(), ids.data(), mix.data(), &h)
                 : sxc_set_functional(_ctx, (int)ids.size(), ids.data(), mix.data(), &h));
    return h;
  }
  void releaseGrid(int h) const noexcept { _group ? sxc_group_release_grid(_group, h) : sxc_release_grid(_ctx, h); }
  void releaseBasis(int h) const noexcept { _group ? sxc_group_release_basis(_group, h) : sxc_release_basis(_ctx, h); }
  void buildXC(int grid, int basis, int func, int nspin, const double* P, double thr, double* V, double* E, double* nel) const {
    check(_group ? sxc_group_build_xc(_group, grid, basis, func, nspin, P, thr, V, E, nel)
                 : sxc_build_xc(_ctx, grid, basis, func, nspin, P, thr, V, E, nel));
  }
  void buildNAddMulti(int grid, int nfunc, const int* funcs, int nspin, int basisAct, const double* Pact, int nenv,
                      const int* basisEnv, const double* const* Penv, int envTag, double thr, bool sumMatrices, double* V,
                      double* E) const {
    check(_group ? sxc_group_build_nadd_multi(_group, grid, nfunc, funcs, nspin, basisAct, Pact, nenv, basisEnv, Penv, envTag, thr,
                                              sumMatrices ? 1 : 0, V, E)
                 : sxc_build_nadd_multi(_ctx, grid, nfunc, funcs, nspin, basisAct, Pact, nenv, basisEnv, Penv, envTag, thr,
                                        sumMatrices ? 1 : 0, V, E));
  }
  void xcGradient(int grid, int basis, int func, int nspin, const double* P, int natoms, const int* atomOfBf, double* grad) const {
    check(_group ? sxc_group_xc_gradient(_group, grid, basis, func, nspin, P, natoms, atomOfBf, grad)
                 : sxc_xc_gradient(_ctx, grid, basis, func, nspin, P, natoms, atomOfBf, grad));
  }

 private:
  sxc_ctx* _ctx = nullptr;
  sxc_group* _group = nullptr;
};

}  // namespace B200

// grid/GridController.h: 3 x N points (Eigen::Matrix3Xd, xyz interleaved) + N weights; uploads once per device.
class GridController : public Notifying {
 public:
  GridController(std::vector<double> xyz, std::vector<double> weights, int blocksize = 128)
    : _xyz(std::move(xyz)), _w(std::move(weights)), _blocksize(blocksize) {}
  GridController(const GridController&) = delete;  // (owns a device handle)
  GridController& operator=(const GridController&) = delete;
  const std::vector<double>& getGridPoints() const { return _xyz; }
  const std::vector<double>& getWeights() const { return _w; }
  unsigned int getNGridPoints() const { return (unsigned int)_w.size(); }
  ~GridController() {  // the device copy (points, per-point work arrays, screening plans) goes with the host object
    if (_handle >= 0)
      if (auto d = _dev.lock()) d->releaseGrid(_handle);
  }
  int handle(const std::shared_ptr<B200::XCDevice>& dev) {  // lazily uploaded (RememberingFactory-style key: this object)
    if (_handle < 0) {
      _handle = dev->setGrid((int64_t)_w.size(), _xyz.data(), _w.data(), _blocksize);
      _dev = dev;
    }
    return _handle;
  }

 private:
  std::vector<double> _xyz, _w;
  int _blocksize;
  int _handle = -1;
  std::weak_ptr<B200::XCDevice> _dev;
};

// basis/BasisController.h + basis/Shell.h: what BasisFunctionOnGridController reads from the shells.
struct ShellTable {
  std::vector<int> l, pure, nprim, firstBf;          // per shell (firstBf = extendedIndex)
  std::vector<double> centre, alpha, coeff, normfac;  // 3 per shell | per primitive (renormalised) | per function
  std::vector<int> atomOfBf;                          // BasisController::getAtomIndicesOfBasis() (gradients only)
  int nAtoms = 0;
};
class BasisController {
 public:
  BasisController(ShellTable t, int nBasisFunctions, double radialThreshold = 1e-9)
    : _t(std::move(t)), _nbf(nBasisFunctions), _thr(radialThreshold) {}
  BasisController(const BasisController& o) : _t(o._t), _nbf(o._nbf), _thr(o._thr) {}  // (a copy owns no device handle yet)
  BasisController& operator=(const BasisController&) = delete;
  // AtomCenteredBasisController + BasisFunctionProvider (basis/BasisFunctionProvider.cpp:32-140): geometry + basis-set file
  // of the reference's data/basis directory -> shells (row f-2, sxc_shell_table_from_file)
  static std::shared_ptr<BasisController> fromFile(const std::string& path, const std::string& label,
                                                   const std::vector<std::string>& elements, const std::vector<double>& coordsBohr,
                                                   bool spherical = true, double radialThreshold = 1e-9) {
    std::vector<const char*> names;
    for (auto& e : elements) names.push_back(e.c_str());
    sxc_shell_table* h = nullptr;
    if (sxc_shell_table_from_file(path.c_str(), label.c_str(), (int)elements.size(), names.data(), coordsBohr.data(), spherical ? 1 : 0,
                                  &h) != SXC_OK)
      throw SerenityError(sxc_host_last_error());
    int ns = 0, np = 0, nbf = 0;
    sxc_shell_table_sizes(h, &ns, &np, &nbf);
    ShellTable t;
    t.l.resize(ns), t.pure.resize(ns), t.nprim.resize(ns), t.firstBf.resize(ns);
    t.centre.resize(3 * (size_t)ns), t.alpha.resize(np), t.coeff.resize(np), t.normfac.resize(nbf), t.atomOfBf.resize(nbf);
    sxc_shell_table_copy(h, t.l.data(), t.pure.data(), t.nprim.data(), t.firstBf.data(), t.centre.data(), t.alpha.data(),
                         t.coeff.data(), t.normfac.data(), t.atomOfBf.data());
    sxc_shell_table_free(h);
    t.nAtoms = (int)elements.size();
    return std::make_shared<BasisController>(std::move(t), nbf, radialThreshold);
  }
  unsigned int getNBasisFunctions() const { return (unsigned int)_nbf; }
  const std::vector<int>& getAtomIndicesOfBasis() const { return _t.atomOfBf; }
  int getNAtoms() const { return _t.nAtoms; }
  ~BasisController() {  // a new geometry makes a new basis: the old shell table and its plans leave the device
    if (_handle >= 0)
      if (auto d = _dev.lock()) d->releaseBasis(_handle);
  }
  int handle(const std::shared_ptr<B200::XCDevice>& dev) {
    if (_handle < 0) {
      _handle = dev->addBasis((int)_t.l.size(), _t.l.data(), _t.pure.data(), _t.nprim.data(), _t.firstBf.data(), _t.centre.data(),
                              _t.alpha.data(), _t.coeff.data(), _t.normfac.data(), _thr);
      _dev = dev;
    }
    return _handle;
  }

 private:
  ShellTable _t;
  int _nbf;
  double _thr;
  int _handle = -1;
  std::weak_ptr<B200::XCDevice> _dev;
};

// data/matrices/DensityMatrixController.h: owns P, notifies dependants when it changes.
template<Options::SCF_MODES SCFMode>
class DensityMatrixController : public Notifying {
 public:
  DensityMatrixController(std::shared_ptr<BasisController> basis, DensityMatrix P) : _basis(std::move(basis)), _P(std::move(P)) {}
  const DensityMatrix& getDensityMatrix() const { return _P; }
  void setDensityMatrix(DensityMatrix P) {
    _P = std::move(P);
    notifyObjects();
  }
  std::shared_ptr<BasisController> getBasisController() const { return _basis; }

 private:
  std::shared_ptr<BasisController> _basis;
  DensityMatrix _P;
};

// dft/Functional.h: basic functionals (BASIC_FUNCTIONALS enum values) and mixing factors
struct Functional {
  std::vector<int> basicFunctionals;
  std::vector<double> mixingFactors;
};

// dft/functionals/CompositeFunctionals.h:165-169 resolveFunctional, for the composites whose basic functionals this library
// implements (rows of dft/functionals/functional_definitions.dat, XCFun route).  The exact-exchange fraction of the hybrids is
// returned separately: it is ERI work and stays with the reference's exchange potentials.
namespace CompositeFunctionals {
inline Functional resolveFunctional(std::string name, double* exactExchange = nullptr) {
  std::transform(name.begin(), name.end(), name.begin(), [](unsigned char c) { return (char)std::toupper(c); });
  struct Row {
    const char* name;
    std::vector<int> ids;
    std::vector<double> mix;
    double hfx;
  };
  static const std::vector<Row> table = {
      {"NONE", {}, {}, 0.0},
      {"LDA", {SXC_X_SLATER, SXC_C_VWN}, {1.0, 1.0}, 0.0},
      {"SLATER", {SXC_X_SLATER}, {1.0}, 0.0},
      {"BLYP", {SXC_X_B88, SXC_C_LYP}, {1.0, 1.0}, 0.0},
      {"PBE", {SXC_X_PBE, SXC_C_PBE}, {1.0, 1.0}, 0.0},
      {"BP86", {SXC_X_B88, SXC_C_P86}, {1.0, 1.0}, 0.0},
      {"BHLYP", {SXC_X_B88, SXC_C_LYP}, {0.50, 1.0}, 0.50},
      {"PBE0", {SXC_X_PBE, SXC_C_PBE}, {0.75, 1.0}, 0.25},
      {"B3LYP", {SXC_X_SLATER, SXC_X_B88_CORR, SXC_C_LYP, SXC_C_VWN}, {0.80, 0.72, 0.81, 0.19}, 0.20},
      {"TF", {SXC_K_TF}, {1.0}, 0.0},
      {"PW91K", {SXC_K_PW91}, {1.0}, 0.0},
      {"LLP91K", {SXC_K_LLP}, {1.0}, 0.0},
  };
  for (const Row& r : table)
    if (name == r.name) {
      if (exactExchange) *exactExchange = r.hfx;
      return Functional{r.ids, r.mix};
    }
  throw SerenityError("CompositeFunctionals::resolveFunctional: functional " + name + " is not available on the B200 path");
}
}  // namespace CompositeFunctionals

// potentials/Potential.h:43-86
template<Options::SCF_MODES SCFMode>
class Potential {
 public:
  virtual ~Potential() = default;
  virtual FockMatrix& getMatrix() = 0;
  virtual double getEnergy(const DensityMatrix& P) = 0;
  virtual Matrix getGeomGradients() = 0;
};

namespace detail {
inline int functionalHandle(const B200::XCDevice& dev, const Functional& f) {
  return dev.setFunctional(f.basicFunctionals, f.mixingFactors);  // (equal definitions share one handle inside the library)
}
// Tag of an environment STATE for the frozen-environment cache of the NAdd builds (sxc_build_nadd: env_frozen).  All NAdd
// objects that watch the same environment density controllers on the same grid share one state - the XC and the kinetic
// object of an FDE iteration are served by one cached sum of environment densities - and a changed environment density
// draws a new tag for all of them.
struct EnvState {
  int tag;
};
inline int nextEnvTag() {
  static int counter = 0;
  return ++counter;
}
inline std::shared_ptr<EnvState> sharedEnvState(const std::vector<const void*>& key) {
  static std::vector<std::pair<std::vector<const void*>, std::weak_ptr<EnvState>>> registry;
  for (auto it = registry.begin(); it != registry.end();) {
    if (it->second.expired()) {
      it = registry.erase(it);
      continue;
    }
    if (it->first == key) return it->second.lock();
    ++it;
  }
  auto st = std::make_shared<EnvState>(EnvState{nextEnvTag()});
  registry.emplace_back(key, st);
  return st;
}
template<Options::SCF_MODES SCFMode>
constexpr int nspin() {
  return SCFMode == Options::SCF_MODES::RESTRICTED ? 1 : 2;
}
}  // namespace detail

// potentials/FuncPotential.h:47-147.  The `system` argument of the reference only supplies settings
// (grid.blockAveThreshold); it is passed here as that number.
template<Options::SCF_MODES SCFMode>
class FuncPotential : public Potential<SCFMode>, public ObjectSensitive {
 public:
  FuncPotential(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<DensityMatrixController<SCFMode>> dMat,
                std::shared_ptr<GridController> grid, Functional functional, double blockAveThreshold = 1e-11)
    : _dev(std::move(device)), _dMatController(std::move(dMat)), _grid(std::move(grid)), _functional(std::move(functional)),
      _thr(blockAveThreshold), _func(detail::functionalHandle(*_dev, _functional)) {}

  // call once after make_shared: registers for lazy invalidation (FuncPotential.cpp:56-63)
  void registerSensitivity(const std::shared_ptr<FuncPotential>& self) {
    _dMatController->addSensitiveObject(self);
    _grid->addSensitiveObject(self);
  }

  FockMatrix& getMatrix() override final {  // FuncPotential.cpp:74-111
    if (!_potential) {
      const int nb = (int)_dMatController->getBasisController()->getNBasisFunctions();
      auto V = std::make_unique<FockMatrix>(nb, nb * detail::nspin<SCFMode>());
      double nel = 0.0;
      _dev->buildXC(_grid->handle(_dev), _dMatController->getBasisController()->handle(_dev), _func, detail::nspin<SCFMode>(),
                    _dMatController->getDensityMatrix().data(), _thr, V->data(), &_energy, &nel);
      _nElectronsOnGrid = nel;
      _potential = std::move(V);
    }
    return *_potential;
  }
  double getEnergy(const DensityMatrix& /*P*/) override final {  // FuncPotential.cpp:67-71: cached, P is ignored
    if (!_potential) getMatrix();
    return _energy;
  }
  Matrix getGeomGradients() override final {  // FuncPotential.cpp:114-239
    auto basis = _dMatController->getBasisController();
    if (basis->getNAtoms() <= 0 || basis->getAtomIndicesOfBasis().size() != basis->getNBasisFunctions())
      throw SerenityError("FuncPotential::getGeomGradients: the basis carries no atom indices");
    Matrix grad(basis->getNAtoms(), 3);
    _dev->xcGradient(_grid->handle(_dev), basis->handle(_dev), _func, detail::nspin<SCFMode>(),
                     _dMatController->getDensityMatrix().data(), basis->getNAtoms(), basis->getAtomIndicesOfBasis().data(),
                     grad.data());
    return grad;
  }
  void notify() override final { _potential = nullptr; }  // FuncPotential.h:107-109
  Functional getFunctional() { return _functional; }
  std::shared_ptr<GridController> getGridController() { return _grid; }
  double getNElectronsOnGrid() const { return _nElectronsOnGrid; }

 private:
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<DensityMatrixController<SCFMode>> _dMatController;
  std::shared_ptr<GridController> _grid;
  Functional _functional;
  double _thr;
  int _func;
  std::unique_ptr<FockMatrix> _potential;
  double _energy = 0.0;
  double _nElectronsOnGrid = 0.0;
};

// ---- stage-level classes (one level below the potentials; RESTRICTED) --------------------------------------------------
// data/grid/DensityOnGridCalculator.h:45-103: rho (and grad rho) of a density matrix on the grid
struct DensityOnGrid {  // data/grid/DensityOnGrid.h + math/Derivatives.h Gradient<>
  std::vector<double> rho, x, y, z;
};
class DensityOnGridCalculator {
 public:
  DensityOnGridCalculator(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<BasisController> basis,
                          std::shared_ptr<GridController> grid)
    : _dev(std::move(device)), _basis(std::move(basis)), _grid(std::move(grid)) {}
  // calcDensityAndGradientOnGrid (DensityOnGridCalculator.cpp:55-65)
  DensityOnGrid calcDensityAndGradientOnGrid(const DensityMatrix& P) {
    const size_t n = _grid->getNGridPoints();
    DensityOnGrid d{std::vector<double>(n), std::vector<double>(n), std::vector<double>(n), std::vector<double>(n)};
    _dev->check(sxc_density_on_grid(_dev->get(), _grid->handle(_dev), _basis->handle(_dev), P.data(), d.rho.data(), d.x.data(),
                                    d.y.data(), d.z.data()));
    return d;
  }

 private:
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<BasisController> _basis;
  std::shared_ptr<GridController> _grid;
};

// dft/functionals/wrappers/XCFun.h / FunctionalLibrary.h: calcData(GRADIENTS, functional, density, order 1)
struct FunctionalData {  // dft/functionals/FunctionalData: epuv, dFdRho, dFdGradRho, energy
  std::vector<double> epuv, dFdRho, dFdGradRhoX, dFdGradRhoY, dFdGradRhoZ;
  double energy = 0.0;
};
class FunctionalLibrary {
 public:
  FunctionalLibrary(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<GridController> grid)
    : _dev(std::move(device)), _grid(std::move(grid)) {}
  FunctionalData calcData(const Functional& functional, const DensityOnGrid& d) {
    const size_t n = _grid->getNGridPoints();
    if (d.rho.size() != n) throw SerenityError("FunctionalLibrary: density and grid do not match");
    FunctionalData f;
    f.epuv.resize(n), f.dFdRho.resize(n), f.dFdGradRhoX.resize(n), f.dFdGradRhoY.resize(n), f.dFdGradRhoZ.resize(n);
    _dev->check(sxc_functional_on_grid(_dev->get(), detail::functionalHandle(*_dev, functional), (int64_t)n,
                                       _grid->getWeights().data(), d.rho.data(), d.x.data(), d.y.data(), d.z.data(), f.epuv.data(),
                                       f.dFdRho.data(), f.dFdGradRhoX.data(), f.dFdGradRhoY.data(), f.dFdGradRhoZ.data(), &f.energy));
    return f;
  }

 private:
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<GridController> _grid;
};

// data/grid/ScalarOperatorToMatrixAdder.h:44-135: adds <mu| v + g . nabla |nu> (symmetrised, :262-303) to a matrix
class ScalarOperatorToMatrixAdder {
 public:
  ScalarOperatorToMatrixAdder(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<BasisController> basis,
                              std::shared_ptr<GridController> grid, double blockAveThreshold = 1e-11)
    : _dev(std::move(device)), _basis(std::move(basis)), _grid(std::move(grid)), _thr(blockAveThreshold) {}
  // addScalarOperatorToMatrix(matrix, scalarPart, gradientPart) (ScalarOperatorToMatrixAdder.cpp:97-116); LDA: pass empty vectors
  void addScalarOperatorToMatrix(Matrix& m, const std::vector<double>& scalar, const std::vector<double>& gx = {},
                                 const std::vector<double>& gy = {}, const std::vector<double>& gz = {}) {
    const bool gga = !gx.empty();
    _dev->check(sxc_scalar_to_matrix(_dev->get(), _grid->handle(_dev), _basis->handle(_dev), _thr, scalar.data(),
                                     gga ? gx.data() : nullptr, gga ? gy.data() : nullptr, gga ? gz.data() : nullptr, m.data()));
  }

 private:
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<BasisController> _basis;
  std::shared_ptr<GridController> _grid;
  double _thr;
};

// potentials/NAddFuncPotential.h:118-155 (exact-exchange and solvation parts are ERI work and stay with the reference's
// ExchangeInteractionPotential).  Both constructors of the reference are mirrored; getGridPotentialDerivative()
// (NAddFuncPotential.h:207) is declared there but has no definition and no caller, so there is nothing to mirror.
template<Options::SCF_MODES SCFMode>
class FDEPotentials;
template<Options::SCF_MODES SCFMode>
class NAddFuncPotential : public Potential<SCFMode>, public ObjectSensitive {
 public:
  using DMC = DensityMatrixController<SCFMode>;
  NAddFuncPotential(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<DMC> activeDMat, std::vector<std::shared_ptr<DMC>> envDMats,
                    std::shared_ptr<GridController> grid, Functional functional, double blockAveThreshold = 1e-11)
    : _dev(std::move(device)), _act(std::move(activeDMat)), _env(std::move(envDMats)), _grid(std::move(grid)),
      _functional(std::move(functional)), _thr(blockAveThreshold), _func(detail::functionalHandle(*_dev, _functional)) {
    initEnvState();
  }
  // Second constructor (NAddFuncPotential.cpp:105-176): further exactly treated subsystems are folded into the active density
  // through their basis-set projections, P_comb = P_act + sum_I BtoA_I^T P_I BtoA_I (BtoA_I: nbf_I x nbf_act; the reference
  // passes Eigen::SparseMatrix, here a dense Matrix).  As in the reference the combination is formed ONCE, here, from the
  // densities of that moment (:142-151 build a fresh controller around the combined matrix).
  NAddFuncPotential(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<DMC> activeDMat,
                    std::vector<std::shared_ptr<DMC>> otherExactDmats, const std::vector<std::shared_ptr<Matrix>>& BtoAProjections,
                    std::vector<std::shared_ptr<DMC>> envDMats, std::shared_ptr<GridController> grid, Functional functional,
                    double blockAveThreshold = 1e-11)
    : _dev(std::move(device)), _act(std::move(activeDMat)), _otherExact(std::move(otherExactDmats)), _env(std::move(envDMats)),
      _grid(std::move(grid)), _functional(std::move(functional)), _thr(blockAveThreshold),
      _func(detail::functionalHandle(*_dev, _functional)) {
    if (BtoAProjections.size() != _otherExact.size())
      throw SerenityError("NAddFuncPotential: one BtoA projection per exactly treated subsystem is needed");
    const int nA = (int)_act->getBasisController()->getNBasisFunctions();
    const int ns = detail::nspin<SCFMode>();
    DensityMatrix comb = _act->getDensityMatrix();
    for (size_t I = 0; I < _otherExact.size(); ++I) {
      const Matrix& B = *BtoAProjections[I];
      const DensityMatrix& D = _otherExact[I]->getDensityMatrix();
      const int nI = B.rows();
      if (B.cols() != nA || D.rows() != nI) throw SerenityError("NAddFuncPotential: BtoA projection has the wrong shape");
      Matrix T(nI, nA);  // T = D_spin * BtoA, then comb_spin += BtoA^T * T
      for (int sp = 0; sp < ns; ++sp) {
        const double* Dp = D.data() + (size_t)sp * nI * nI;
        for (int a = 0; a < nA; ++a)
          for (int i = 0; i < nI; ++i) {
            double t = 0.0;
            for (int j = 0; j < nI; ++j) t += Dp[i + (size_t)nI * j] * B(j, a);
            T(i, a) = t;
          }
        double* Cp = comb.data() + (size_t)sp * nA * nA;
        for (int b = 0; b < nA; ++b)
          for (int a = 0; a < nA; ++a) {
            double t = 0.0;
            for (int i = 0; i < nI; ++i) t += B(i, a) * T(i, b);
            Cp[a + (size_t)nA * b] += t;
          }
      }
    }
    _combined = std::make_shared<DMC>(_act->getBasisController(), std::move(comb));
    initEnvState();
  }

  struct EnvWatcher : ObjectSensitive {  // a changed environment density: new tag -> the cached rho_env on the grid is not used
    explicit EnvWatcher(NAddFuncPotential* o) : owner(o) {}
    void notify() override {
      owner->_potential = nullptr;
      owner->_envState->tag = detail::nextEnvTag();
    }
    NAddFuncPotential* owner;
  };
  void registerSensitivity(const std::shared_ptr<NAddFuncPotential>& self) {
    _act->addSensitiveObject(self);
    _grid->addSensitiveObject(self);
    for (auto& o : _otherExact) o->addSensitiveObject(self);
    _watcher = std::make_shared<EnvWatcher>(this);
    for (auto& e : _env) e->addSensitiveObject(_watcher);
  }

  FockMatrix& getMatrix() override final {  // NAddFuncPotential.cpp:192-300
    if (!_potential) {
      const int nb = (int)_act->getBasisController()->getNBasisFunctions();
      auto V = std::make_unique<FockMatrix>(nb, nb * detail::nspin<SCFMode>());
      std::vector<int> be;
      std::vector<const double*> pe;
      envArguments(be, pe);
      std::vector<double> parts(2 + _env.size(), 0.0);
      _dev->buildNAddMulti(_grid->handle(_dev), 1, &_func, detail::nspin<SCFMode>(), _act->getBasisController()->handle(_dev),
                           activeDensity().data(), (int)_env.size(), be.data(), pe.data(), _envState->tag, _thr, false, V->data(),
                           parts.data());
      setResult(std::move(V), parts);
    }
    return *_potential;
  }
  double getEnergy(const DensityMatrix& /*P*/) override final {
    if (!_potential) getMatrix();
    return _energy;
  }
  // NAddFuncPotential.cpp:180-189: scaling * sum_spin sum_ij V_ij P_ij
  double getLinearizedEnergy(const DensityMatrix& P, double scaling) {
    if (!_potential) getMatrix();
    const FockMatrix& pot = *_potential;
    if ((size_t)P.rows() * P.cols() != (size_t)pot.rows() * pot.cols())
      throw SerenityError("NAddFuncPotential::getLinearizedEnergy: density matrix and potential do not match");
    double e = 0.0;
    for (size_t k = 0; k < pot.values.size(); ++k) e += pot.values[k] * P.values[k];
    return scaling * e;
  }
  Matrix getGeomGradients() override final {  // NAddFuncPotential.cpp:329-493
    auto basis = _act->getBasisController();
    if (basis->getNAtoms() <= 0 || basis->getAtomIndicesOfBasis().size() != basis->getNBasisFunctions())
      throw SerenityError("NAddFuncPotential: Missed gradient element in gradient evaluation.");  // :367-369
    std::vector<int> be;
    std::vector<const double*> pe;
    envArguments(be, pe);
    Matrix grad(basis->getNAtoms(), 3);
    _dev->check(sxc_nadd_gradient(_dev->get(), _grid->handle(_dev), _func, detail::nspin<SCFMode>(), basis->handle(_dev),
                                  activeDensity().data(), (int)_env.size(), be.data(), pe.data(), basis->getNAtoms(),
                                  basis->getAtomIndicesOfBasis().data(), grad.data()));
    _envState->tag = detail::nextEnvTag();  // the gradient reuses the device buffer of the cached environment density
    return grad;
  }
  void notify() override final { _potential = nullptr; }
  const std::vector<double>& getEnergyParts() const { return _energyParts; }
  Functional getFunctional() { return _functional; }
  std::shared_ptr<GridController> getGridController() { return _grid; }

 private:
  friend class FDEPotentials<SCFMode>;
  void initEnvState() {
    std::vector<const void*> key{_grid.get()};
    for (auto& e : _env) key.push_back(e.get());
    _envState = detail::sharedEnvState(key);
  }
  const DensityMatrix& activeDensity() const { return (_combined ? _combined : _act)->getDensityMatrix(); }
  void envArguments(std::vector<int>& be, std::vector<const double*>& pe) {
    for (auto& e : _env) {
      be.push_back(e->getBasisController()->handle(_dev));
      pe.push_back(e->getDensityMatrix().data());
    }
  }
  void setResult(std::unique_ptr<FockMatrix> V, const std::vector<double>& parts) {
    _energyParts = parts;
    _energy = parts[0] - parts[1];  // E[rho_tot] - E[rho_act] - sum_env E[rho_env], :249, :282-286
    for (size_t i = 2; i < parts.size(); ++i) _energy -= parts[i];
    _potential = std::move(V);
  }

  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<DMC> _act;
  std::vector<std::shared_ptr<DMC>> _otherExact;
  std::shared_ptr<DMC> _combined;  // second constructor: P_act + projected exactly treated densities
  std::vector<std::shared_ptr<DMC>> _env;
  std::shared_ptr<GridController> _grid;
  Functional _functional;
  double _thr;
  int _func;
  std::shared_ptr<EnvWatcher> _watcher;
  std::shared_ptr<detail::EnvState> _envState;
  std::unique_ptr<FockMatrix> _potential;
  std::vector<double> _energyParts;
  double _energy = 0.0;
};

// potentials/bundles/FDEPotentials.h: the part of the bundle that lives on the grid.  FDEPotentials::getFockMatrix
// (FDEPotentials.cpp:43-61) adds naddXC->getMatrix() and naddKin->getMatrix() of the same active / environment densities; here
// both objects are evaluated in ONE device pass (sxc_build_nadd_multi: rho_act contracted once, both functionals on the same
// densities, one scatter per object) and each object receives its own matrix and energies, exactly as if it had been asked
// alone - later getMatrix() / getEnergy() calls on the objects are served from their caches.
template<Options::SCF_MODES SCFMode>
class FDEPotentials {
 public:
  FDEPotentials(std::shared_ptr<NAddFuncPotential<SCFMode>> naddXC, std::shared_ptr<NAddFuncPotential<SCFMode>> naddKin)
    : _xc(std::move(naddXC)), _kin(std::move(naddKin)) {
    if (_xc->_act != _kin->_act || _xc->_grid != _kin->_grid || _xc->_env != _kin->_env || _xc->_combined || _kin->_combined)
      throw SerenityError("FDEPotentials: the non-additive XC and kinetic potentials must share active density, environment and grid");
  }
  // sum of the two non-additive matrices (what the bundle adds to the Fock matrix)
  FockMatrix getNAddFockMatrix() {
    if (!_xc->_potential || !_kin->_potential) {
      auto& x = *_xc;
      const int nb = (int)x._act->getBasisController()->getNBasisFunctions();
      const int ns = detail::nspin<SCFMode>();
      const size_t nv = (size_t)nb * nb * ns, ne = 2 + x._env.size();
      std::vector<int> be;
      std::vector<const double*> pe;
      x.envArguments(be, pe);
      std::vector<double> V(2 * nv), parts(2 * ne);
      const int funcs[2] = {_xc->_func, _kin->_func};
      x._dev->buildNAddMulti(x._grid->handle(x._dev), 2, funcs, ns, x._act->getBasisController()->handle(x._dev),
                             x._act->getDensityMatrix().data(), (int)x._env.size(), be.data(), pe.data(), x._envState->tag, x._thr,
                             false, V.data(), parts.data());
      for (int k = 0; k < 2; ++k) {
        auto M = std::make_unique<FockMatrix>(nb, nb * ns);
        std::copy(V.begin() + k * nv, V.begin() + (k + 1) * nv, M->values.begin());
        (k == 0 ? _xc : _kin)->setResult(std::move(M), std::vector<double>(parts.begin() + k * ne, parts.begin() + (k + 1) * ne));
      }
    }
    FockMatrix F = _xc->getMatrix();
    const FockMatrix& K = _kin->getMatrix();
    for (size_t i = 0; i < F.values.size(); ++i) F.values[i] += K.values[i];
    return F;
  }

 private:
  std::shared_ptr<NAddFuncPotential<SCFMode>> _xc, _kin;
};

// potentials/ABFockMatrixConstruction/ABFuncPotential.h: the XC operator between two different basis
// sets A and B on one grid, from the sum of the densities of `dMats` (each in its own basis).  getMatrix() returns the
// nbf_A x nbf_B matrix (UNRESTRICTED: alpha and beta blocks side by side) and caches it until a density, the grid or a
// basis notifies (ABFuncPotential.cpp:40-52).
template<Options::SCF_MODES SCFMode>
class ABFuncPotential : public ObjectSensitive {
 public:
  ABFuncPotential(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<BasisController> basisA,
                  std::shared_ptr<BasisController> basisB, std::shared_ptr<GridController> grid,
                  std::vector<std::shared_ptr<DensityMatrixController<SCFMode>>> dMats, Functional functional,
                  double blockAveThreshold = 1e-11)
    : _dev(std::move(device)), _basisA(std::move(basisA)), _basisB(std::move(basisB)), _grid(std::move(grid)),
      _dMats(std::move(dMats)), _functional(std::move(functional)), _thr(blockAveThreshold),
      _func(detail::functionalHandle(*_dev, _functional)) {
    if (_dMats.empty()) throw SerenityError("ABFuncPotential: at least one density matrix controller is needed");
  }
  void registerSensitivity(const std::shared_ptr<ABFuncPotential>& self) {
    for (auto& d : _dMats) d->addSensitiveObject(self);
    _grid->addSensitiveObject(self);
  }
  Matrix& getMatrix() {  // ABFuncPotential.cpp:54-160
    if (!_abPotential) {
      const int nA = (int)_basisA->getNBasisFunctions(), nB = (int)_basisB->getNBasisFunctions();
      auto V = std::make_unique<Matrix>(nA, nB * detail::nspin<SCFMode>());
      std::vector<int> bc;
      std::vector<const double*> pc;
      for (auto& d : _dMats) {
        bc.push_back(d->getBasisController()->handle(_dev));
        pc.push_back(d->getDensityMatrix().data());
      }
      double e[2] = {0.0, 0.0};
      _dev->check(sxc_build_ab(_dev->get(), _grid->handle(_dev), _func, detail::nspin<SCFMode>(), _basisA->handle(_dev),
                               _basisB->handle(_dev), (int)bc.size(), bc.data(), pc.data(), _thr, V->data(), e));
      _abPotential = std::move(V);
    }
    return *_abPotential;
  }
  void notify() override final { _abPotential = nullptr; }

 private:
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<BasisController> _basisA, _basisB;
  std::shared_ptr<GridController> _grid;
  std::vector<std::shared_ptr<DensityMatrixController<SCFMode>>> _dMats;
  Functional _functional;
  double _thr;
  int _func;
  std::unique_ptr<Matrix> _abPotential;
};

// potentials/ABFockMatrixConstruction/ABNAddFuncPotential.h: the non-additive potential v[rho_act + sum rho_env] - v[rho_act]
// between two different basis sets A and B (ABNAddFuncPotential.cpp:66-176).  The environment density on the grid is kept by
// the reference until the object dies (:70-72); here the whole matrix is cached until notify().
template<Options::SCF_MODES SCFMode>
class ABNAddFuncPotential : public ObjectSensitive {
 public:
  ABNAddFuncPotential(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<DensityMatrixController<SCFMode>> actDMat,
                      std::shared_ptr<BasisController> basisA, std::shared_ptr<BasisController> basisB,
                      std::vector<std::shared_ptr<DensityMatrixController<SCFMode>>> envDMats,
                      std::shared_ptr<GridController> grid, Functional functional, double blockAveThreshold = 1e-11)
    : _dev(std::move(device)), _act(std::move(actDMat)), _basisA(std::move(basisA)), _basisB(std::move(basisB)),
      _env(std::move(envDMats)), _grid(std::move(grid)), _functional(std::move(functional)), _thr(blockAveThreshold),
      _func(detail::functionalHandle(*_dev, _functional)) {}
  void registerSensitivity(const std::shared_ptr<ABNAddFuncPotential>& self) {
    _act->addSensitiveObject(self);
    for (auto& d : _env) d->addSensitiveObject(self);
    _grid->addSensitiveObject(self);
  }
  Matrix& getMatrix() {
    if (!_abPotential) {
      const int nA = (int)_basisA->getNBasisFunctions(), nB = (int)_basisB->getNBasisFunctions();
      auto V = std::make_unique<Matrix>(nA, nB * detail::nspin<SCFMode>());
      std::vector<int> be;
      std::vector<const double*> pe;
      for (auto& d : _env) {
        be.push_back(d->getBasisController()->handle(_dev));
        pe.push_back(d->getDensityMatrix().data());
      }
      _dev->check(sxc_build_ab_nadd(_dev->get(), _grid->handle(_dev), _func, detail::nspin<SCFMode>(), _basisA->handle(_dev),
                                    _basisB->handle(_dev), _act->getBasisController()->handle(_dev),
                                    _act->getDensityMatrix().data(), (int)be.size(), be.data(), pe.data(), _thr, V->data()));
      _abPotential = std::move(V);
    }
    return *_abPotential;
  }
  void notify() override final { _abPotential = nullptr; }

 private:
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<DensityMatrixController<SCFMode>> _act;
  std::shared_ptr<BasisController> _basisA, _basisB;
  std::vector<std::shared_ptr<DensityMatrixController<SCFMode>>> _env;
  std::shared_ptr<GridController> _grid;
  Functional _functional;
  double _thr;
  int _func;
  std::unique_ptr<Matrix> _abPotential;
};

// postHF/LRSCF/Kernel/Kernel.h:50-208: the second functional derivatives of the subsystem XC functionals
// and of the non-additive XC / kinetic functionals on one grid.  Like the reference the object keeps a "total" set
// (non-additive functionals on the summed density, _pptot/_pgtot/_ggtot) and one set per subsystem (func_I - naddXC - naddKin
// on rho_I, _pp/_pg/_gg); getPP(I, J) = total + (I == J ? subsystem I : 0) (Kernel.cpp:170-230).  The data stays on the device;
// the handles are what KernelSigmavector contracts with.  SCFMode = UNRESTRICTED with RESTRICTED-halved densities is the
// reference's "ukernel" of the triplet case (Kernel.cpp:905-930).
template<Options::SCF_MODES SCFMode>
class Kernel {
 public:
  Kernel(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<GridController> grid,
         std::vector<std::shared_ptr<DensityMatrixController<SCFMode>>> dMats, std::vector<Functional> funcs,
         Functional naddXCFunc = Functional(), Functional naddKinFunc = Functional(), bool gga = true)
    : _dev(std::move(device)), _grid(std::move(grid)), _dMats(std::move(dMats)), _gga(gga) {
    if (_dMats.empty() || funcs.size() != _dMats.size())
      throw SerenityError("Kernel: one functional per subsystem is needed");
    const int g = _grid->handle(_dev);
    const int nspin = detail::nspin<SCFMode>();
    const bool embedded = !naddXCFunc.basicFunctionals.empty() || !naddKinFunc.basicFunctionals.empty();
    std::vector<int> bc;
    std::vector<const double*> pc;
    for (auto& d : _dMats) {
      bc.push_back(d->getBasisController()->handle(_dev));
      pc.push_back(d->getDensityMatrix().data());
    }
    // calculateDerivatives (Kernel.cpp:686-747)
    for (size_t I = 0; I < _dMats.size(); ++I) {
      int k = -1;
      _dev->check(sxc_kernel_create(_dev->get(), g, nspin, _gga ? 1 : 0, &k));
      _sub.push_back(k);
      _dev->check(sxc_kernel_add(_dev->get(), k, detail::functionalHandle(*_dev, funcs[I]), 1.0, 1, &bc[I], &pc[I]));
      if (!naddXCFunc.basicFunctionals.empty())
        _dev->check(sxc_kernel_add(_dev->get(), k, detail::functionalHandle(*_dev, naddXCFunc), -1.0, 1, &bc[I], &pc[I]));
      if (!naddKinFunc.basicFunctionals.empty())
        _dev->check(sxc_kernel_add(_dev->get(), k, detail::functionalHandle(*_dev, naddKinFunc), -1.0, 1, &bc[I], &pc[I]));
    }
    if (embedded) {
      _dev->check(sxc_kernel_create(_dev->get(), g, nspin, _gga ? 1 : 0, &_tot));
      if (!naddXCFunc.basicFunctionals.empty())
        _dev->check(sxc_kernel_add(_dev->get(), _tot, detail::functionalHandle(*_dev, naddXCFunc), 1.0, (int)bc.size(), bc.data(),
                                   pc.data()));
      if (!naddKinFunc.basicFunctionals.empty())
        _dev->check(sxc_kernel_add(_dev->get(), _tot, detail::functionalHandle(*_dev, naddKinFunc), 1.0, (int)bc.size(), bc.data(),
                                   pc.data()));
    }
  }
  // Mixed exact / approximate embedding, Kernel::calculateDerivativesMixedEmbedding (Kernel.cpp:752-888): subsystems whose
  // embedding mode is LEVELSHIFT or HUZINAGA are "exact" (their non-additive XC functional is naddXCExact and they carry no
  // non-additive kinetic term), all others "approximate" (naddXCApprox + naddKinFunc).  Three kinds of store:
  //   sub[I]  + func_I[rho_I]; exact: - naddXCExact[rho_I]; approximate: - naddXCApprox[rho_I] - naddKin[rho_I]      (:788-850)
  //   tot     + naddXCApprox[rho_all] + naddKin[rho_all]                                                                 (:857-870)
  //   exact   + naddXCExact[rho_ex] - naddXCApprox[rho_ex] - naddKin[rho_ex],  rho_ex = sum over the exact subsystems      (:871-887)
  // (the `samedensity` list of the reference, which drops duplicate densities from the sums, is not mirrored)
  Kernel(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<GridController> grid,
         std::vector<std::shared_ptr<DensityMatrixController<SCFMode>>> dMats, std::vector<Functional> funcs,
         std::vector<Options::KIN_EMBEDDING_MODES> embeddingModes, Functional naddXCExact, Functional naddXCApprox,
         Functional naddKinFunc, bool gga = true)
    : _dev(std::move(device)), _grid(std::move(grid)), _dMats(std::move(dMats)), _gga(gga), _modes(std::move(embeddingModes)) {
    if (_dMats.empty() || funcs.size() != _dMats.size() || _modes.size() != _dMats.size())
      throw SerenityError("Kernel: one functional and one embedding mode per subsystem are needed");
    const int g = _grid->handle(_dev);
    const int nspin = detail::nspin<SCFMode>();
    std::vector<int> bc, bex;
    std::vector<const double*> pc, pex;
    for (size_t I = 0; I < _dMats.size(); ++I) {
      const auto m = _modes[I];
      if (m == Options::KIN_EMBEDDING_MODES::FERMI_SHIFTED_HUZINAGA || m == Options::KIN_EMBEDDING_MODES::HOFFMANN ||
          m == Options::KIN_EMBEDDING_MODES::RECONSTRUCTION)
        throw SerenityError("Exact embedding mode not supported in list input yet!");  // Kernel.cpp:820-824
      bc.push_back(_dMats[I]->getBasisController()->handle(_dev));
      pc.push_back(_dMats[I]->getDensityMatrix().data());
      if (isExact(I)) {
        bex.push_back(bc.back());
        pex.push_back(pc.back());
      }
    }
    auto add = [&](int store, const Functional& f, double sign, int n, const int* b, const double* const* p) {
      if (!f.basicFunctionals.empty() && n > 0)
        _dev->check(sxc_kernel_add(_dev->get(), store, detail::functionalHandle(*_dev, f), sign, n, b, p));
    };
    for (size_t I = 0; I < _dMats.size(); ++I) {
      int k = -1;
      _dev->check(sxc_kernel_create(_dev->get(), g, nspin, _gga ? 1 : 0, &k));
      _sub.push_back(k);
      add(k, funcs[I], 1.0, 1, &bc[I], &pc[I]);
      if (isExact(I)) {
        add(k, naddXCExact, -1.0, 1, &bc[I], &pc[I]);
      } else {
        add(k, naddXCApprox, -1.0, 1, &bc[I], &pc[I]);
        add(k, naddKinFunc, -1.0, 1, &bc[I], &pc[I]);
      }
    }
    _dev->check(sxc_kernel_create(_dev->get(), g, nspin, _gga ? 1 : 0, &_tot));
    add(_tot, naddXCApprox, 1.0, (int)bc.size(), bc.data(), pc.data());
    add(_tot, naddKinFunc, 1.0, (int)bc.size(), bc.data(), pc.data());
    _dev->check(sxc_kernel_create(_dev->get(), g, nspin, _gga ? 1 : 0, &_exact));
    add(_exact, naddXCExact, 1.0, (int)bex.size(), bex.data(), pex.data());
    add(_exact, naddXCApprox, -1.0, (int)bex.size(), bex.data(), pex.data());
    add(_exact, naddKinFunc, -1.0, (int)bex.size(), bex.data(), pex.data());
  }
  ~Kernel() {
    for (int k : _sub) sxc_kernel_destroy(_dev->get(), k);
    if (_tot >= 0) sxc_kernel_destroy(_dev->get(), _tot);
    if (_exact >= 0) sxc_kernel_destroy(_dev->get(), _exact);
  }
  Kernel(const Kernel&) = delete;
  Kernel& operator=(const Kernel&) = delete;
  bool isGGA() const { return _gga; }
  unsigned int getNSystems() const { return (unsigned int)_dMats.size(); }
  std::shared_ptr<GridController> getGridController() const { return _grid; }
  std::shared_ptr<BasisController> getBasisController(unsigned I) const { return _dMats[I]->getBasisController(); }
  // device handles behind getPP/getPG/getGG(I, J, ...)
  std::vector<int> stores(unsigned I, unsigned J) const {
    std::vector<int> k;
    if (_tot >= 0) k.push_back(_tot);
    // the store of the exactly embedded subsystems enters between two of them with the same mode (Kernel.cpp:182-189)
    if (_exact >= 0 && _modes[I] == _modes[J] && isExact(I)) k.push_back(_exact);
    if (I == J) k.push_back(_sub[I]);
    return k;
  }
  bool mixedEmbeddingUsed() const { return _exact >= 0; }
  int exactStore() const { return _exact; }
  int totalStore() const { return _tot; }
  // Kernel::getPP(I, J, blockSize, iGridStart) (Kernel.cpp:170-230): first array(s) of the summed stores for one block
  std::vector<double> getPP(unsigned I, unsigned J, unsigned blockSize, unsigned iGridStart) const {
    const unsigned N = _grid->getNGridPoints();
    const int npp = SCFMode == Options::SCF_MODES::RESTRICTED ? 1 : 3;
    std::vector<double> out((size_t)npp * blockSize, 0.0);
    for (int k : stores(I, J)) {
      const int narr = sxc_kernel_num_arrays(_dev->get(), k);
      std::vector<double> all((size_t)narr * N);
      _dev->check(sxc_kernel_get(_dev->get(), k, all.data()));
      for (int a = 0; a < npp; ++a)
        for (unsigned p = 0; p < blockSize && iGridStart + p < N; ++p) out[(size_t)a * blockSize + p] += all[(size_t)a * N + iGridStart + p];
    }
    return out;
  }

 private:
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<GridController> _grid;
  std::vector<std::shared_ptr<DensityMatrixController<SCFMode>>> _dMats;
  bool isExact(size_t I) const {
    return !_modes.empty() &&
           (_modes[I] == Options::KIN_EMBEDDING_MODES::LEVELSHIFT || _modes[I] == Options::KIN_EMBEDDING_MODES::HUZINAGA);
  }
  bool _gga;
  std::vector<Options::KIN_EMBEDDING_MODES> _modes;  // empty: not the mixed-embedding variant
  std::vector<int> _sub;
  int _tot = -1, _exact = -1;
};

// postHF/LRSCF/Sigmavectors/KernelSigmavector.h:40-123.  calcF(I, J, D) returns the Fock-like matrices of
// KernelSigmavector.cpp:119-252 for the trial densities D (in the basis of subsystem J; UNRESTRICTED: alpha and beta matrices
// of every vector back to back); contractSupersystemDensity() is :60-117 (all subsystems' densities contracted once with the
// total-density kernel).  A RESTRICTED sigma vector built with an UNRESTRICTED kernel is the triplet case (:381-404).
template<Options::SCF_MODES SCFMode, Options::SCF_MODES KernelMode = SCFMode>
class KernelSigmavector {
 public:
  KernelSigmavector(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<Kernel<KernelMode>> kernel)
    : _dev(std::move(device)), _kernel(std::move(kernel)) {
    if (!_kernel) throw SerenityError("A kernel sigma vector was requested with no kernel present.");
  }
  // D[J] = trial densities of subsystem J (nvec x nspin matrices); afterwards calcF(I, I, ...) adds the supersystem part
  void contractSupersystemDensity(const std::vector<std::vector<Matrix>>& D) {
    if (_kernel->totalStore() < 0) throw SerenityError("KernelSigmavector: no embedding kernel to contract");
    const int tot = _kernel->totalStore();
    for (unsigned J = 0; J < D.size(); ++J) contract(J, {tot}, D[J], J != 0);
    _dev->check(sxc_kernel_response_copy(_dev->get(), _kernel->getGridController()->handle(_dev), 1));
    _supersystem = true;
  }
  std::vector<Matrix> calcF(unsigned I, unsigned J, const std::vector<Matrix>& densityMatrices) {
    if (I != J && _supersystem) return {};  // already inside the supersystem contraction (:142-145)
    std::vector<int> k = _kernel->stores(I, J);
    if (_supersystem) {  // the total-density part is in the saved supersystem response: add the subsystem store only
      k = {k.back()};
      _dev->check(sxc_kernel_response_copy(_dev->get(), _kernel->getGridController()->handle(_dev), 0));
    }
    if (k.empty()) return {};
    contract(J, k, densityMatrices, _supersystem);
    const int nb = (int)_kernel->getBasisController(I)->getNBasisFunctions();
    const size_t nmat = densityMatrices.size();
    std::vector<double> flat(nmat * (size_t)nb * nb);
    _dev->check(sxc_kernel_integrate(_dev->get(), _kernel->getGridController()->handle(_dev),
                                     _kernel->getBasisController(I)->handle(_dev), flat.data()));
    std::vector<Matrix> F;
    for (size_t m = 0; m < nmat; ++m) {
      Matrix M(nb, nb);
      std::copy(flat.begin() + m * (size_t)nb * nb, flat.begin() + (m + 1) * (size_t)nb * nb, M.values.begin());
      F.push_back(std::move(M));
    }
    return F;
  }

 private:
  static constexpr int mode() {
    return SCFMode == Options::SCF_MODES::UNRESTRICTED ? 2 : (KernelMode == Options::SCF_MODES::UNRESTRICTED ? 1 : 0);
  }
  void contract(unsigned J, const std::vector<int>& stores, const std::vector<Matrix>& D, bool accumulate) {
    const int nb = (int)_kernel->getBasisController(J)->getNBasisFunctions();
    const int nspin = detail::nspin<SCFMode>();
    if (D.empty() || D.size() % nspin) throw SerenityError("KernelSigmavector: nvec x nspin density matrices are needed");
    std::vector<double> flat;
    for (const Matrix& M : D) {
      if (M.rows() != nb || M.cols() != nb) throw SerenityError("KernelSigmavector: density matrix of the wrong dimension");
      flat.insert(flat.end(), M.values.begin(), M.values.end());
    }
    _dev->check(sxc_kernel_contract(_dev->get(), _kernel->getGridController()->handle(_dev),
                                    _kernel->getBasisController(J)->handle(_dev), (int)stores.size(), stores.data(), mode(),
                                    (int)D.size() / nspin, flat.data(), accumulate ? 1 : 0));
  }
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<Kernel<KernelMode>> _kernel;
  bool _supersystem = false;
};

}  // namespace Serenity
#endif
