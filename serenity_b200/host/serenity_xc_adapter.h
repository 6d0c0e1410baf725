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

// One device context per process and GPU (sxc_create); shared by all potentials of the process.
class XCDevice {
 public:
  explicit XCDevice(int device = 0) {
    const int rc = sxc_create(&_ctx, device);
    if (rc != SXC_OK)
      throw SerenityError("serenity_xc_b200: no usable CUDA device " + std::to_string(device) + " (status " +
                          std::to_string(rc) + "); the XC build has no CPU fallback");
  }
  ~XCDevice() { sxc_destroy(_ctx); }
  XCDevice(const XCDevice&) = delete;
  XCDevice& operator=(const XCDevice&) = delete;
  sxc_ctx* get() const { return _ctx; }
  void check(int rc) const {
    if (rc != SXC_OK) throw SerenityError(std::string("serenity_xc_b200: ") + sxc_last_error(_ctx));
  }

 private:
  sxc_ctx* _ctx = nullptr;
};

}  // namespace B200

// grid/GridController.h: 3 x N points (Eigen::Matrix3Xd, xyz interleaved) + N weights; uploads once per device.
class GridController : public Notifying {
 public:
  GridController(std::vector<double> xyz, std::vector<double> weights, int blocksize = 128)
    : _xyz(std::move(xyz)), _w(std::move(weights)), _blocksize(blocksize) {}
  const std::vector<double>& getGridPoints() const { return _xyz; }
  const std::vector<double>& getWeights() const { return _w; }
  unsigned int getNGridPoints() const { return (unsigned int)_w.size(); }
  int handle(const B200::XCDevice& dev) {  // lazily uploaded (RememberingFactory-style key: this object)
    if (_handle < 0) dev.check(sxc_set_grid(dev.get(), (int64_t)_w.size(), _xyz.data(), _w.data(), _blocksize, &_handle));
    return _handle;
  }

 private:
  std::vector<double> _xyz, _w;
  int _blocksize;
  int _handle = -1;
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
  int handle(const B200::XCDevice& dev) {
    if (_handle < 0)
      dev.check(sxc_add_basis(dev.get(), (int)_t.l.size(), _t.l.data(), _t.pure.data(), _t.nprim.data(), _t.firstBf.data(),
                              _t.centre.data(), _t.alpha.data(), _t.coeff.data(), _t.normfac.data(), _thr, &_handle));
    return _handle;
  }

 private:
  ShellTable _t;
  int _nbf;
  double _thr;
  int _handle = -1;
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
  int h = -1;
  dev.check(sxc_set_functional(dev.get(), (int)f.basicFunctionals.size(), f.basicFunctionals.data(), f.mixingFactors.data(), &h));
  return h;
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
      _dev->check(sxc_build_xc(_dev->get(), _grid->handle(*_dev), _dMatController->getBasisController()->handle(*_dev), _func,
                               detail::nspin<SCFMode>(), _dMatController->getDensityMatrix().data(), _thr, V->data(), &_energy,
                               &nel));
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
    _dev->check(sxc_xc_gradient(_dev->get(), _grid->handle(*_dev), basis->handle(*_dev), _func, detail::nspin<SCFMode>(),
                                _dMatController->getDensityMatrix().data(), basis->getNAtoms(),
                                basis->getAtomIndicesOfBasis().data(), grad.data()));
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
    _dev->check(sxc_density_on_grid(_dev->get(), _grid->handle(*_dev), _basis->handle(*_dev), P.data(), d.rho.data(), d.x.data(),
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
    _dev->check(sxc_scalar_to_matrix(_dev->get(), _grid->handle(*_dev), _basis->handle(*_dev), _thr, scalar.data(),
                                     gga ? gx.data() : nullptr, gga ? gy.data() : nullptr, gga ? gz.data() : nullptr, m.data()));
  }

 private:
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<BasisController> _basis;
  std::shared_ptr<GridController> _grid;
  double _thr;
};

// potentials/NAddFuncPotential.h:118-155 (first constructor; exact-exchange and solvation parts are ERI work and stay
// with the reference's ExchangeInteractionPotential).
template<Options::SCF_MODES SCFMode>
class NAddFuncPotential : public Potential<SCFMode>, public ObjectSensitive {
 public:
  NAddFuncPotential(std::shared_ptr<B200::XCDevice> device, std::shared_ptr<DensityMatrixController<SCFMode>> activeDMat,
                    std::vector<std::shared_ptr<DensityMatrixController<SCFMode>>> envDMats,
                    std::shared_ptr<GridController> grid, Functional functional, double blockAveThreshold = 1e-11)
    : _dev(std::move(device)), _act(std::move(activeDMat)), _env(std::move(envDMats)), _grid(std::move(grid)),
      _functional(std::move(functional)), _thr(blockAveThreshold), _func(detail::functionalHandle(*_dev, _functional)) {}

  struct EnvWatcher : ObjectSensitive {  // a changed environment density invalidates the cached rho_env on the grid
    explicit EnvWatcher(NAddFuncPotential* o) : owner(o) {}
    void notify() override {
      owner->_potential = nullptr;
      owner->_envFrozen = false;
    }
    NAddFuncPotential* owner;
  };
  void registerSensitivity(const std::shared_ptr<NAddFuncPotential>& self) {
    _act->addSensitiveObject(self);
    _grid->addSensitiveObject(self);
    _watcher = std::make_shared<EnvWatcher>(this);
    for (auto& e : _env) e->addSensitiveObject(_watcher);
  }

  FockMatrix& getMatrix() override final {  // NAddFuncPotential.cpp:192-300
    if (!_potential) {
      const int nb = (int)_act->getBasisController()->getNBasisFunctions();
      auto V = std::make_unique<FockMatrix>(nb, nb * detail::nspin<SCFMode>());
      std::vector<int> be;
      std::vector<const double*> pe;
      for (auto& e : _env) {
        be.push_back(e->getBasisController()->handle(*_dev));
        pe.push_back(e->getDensityMatrix().data());
      }
      _energyParts.assign(2 + _env.size(), 0.0);
      _dev->check(sxc_build_nadd(_dev->get(), _grid->handle(*_dev), _func, detail::nspin<SCFMode>(),
                                 _act->getBasisController()->handle(*_dev), _act->getDensityMatrix().data(), (int)_env.size(),
                                 be.data(), pe.data(), _envFrozen ? 1 : 0, _thr, V->data(), _energyParts.data()));
      _energy = _energyParts[0] - _energyParts[1];  // E[rho_tot] - E[rho_act] - sum_env E[rho_env], :249, :282-286
      for (size_t i = 2; i < _energyParts.size(); ++i) _energy -= _energyParts[i];
      _envFrozen = true;
      _potential = std::move(V);
    }
    return *_potential;
  }
  double getEnergy(const DensityMatrix& /*P*/) override final {
    if (!_potential) getMatrix();
    return _energy;
  }
  Matrix getGeomGradients() override final {  // NAddFuncPotential.cpp:329-493
    auto basis = _act->getBasisController();
    if (basis->getNAtoms() <= 0 || basis->getAtomIndicesOfBasis().size() != basis->getNBasisFunctions())
      throw SerenityError("NAddFuncPotential: Missed gradient element in gradient evaluation.");  // :367-369
    std::vector<int> be;
    std::vector<const double*> pe;
    for (auto& e : _env) {
      be.push_back(e->getBasisController()->handle(*_dev));
      pe.push_back(e->getDensityMatrix().data());
    }
    Matrix grad(basis->getNAtoms(), 3);
    _dev->check(sxc_nadd_gradient(_dev->get(), _grid->handle(*_dev), _func, detail::nspin<SCFMode>(), basis->handle(*_dev),
                                  _act->getDensityMatrix().data(), (int)_env.size(), be.data(), pe.data(), basis->getNAtoms(),
                                  basis->getAtomIndicesOfBasis().data(), grad.data()));
    _envFrozen = false;  // the gradient reuses the device buffer of the cached environment density
    return grad;
  }
  void notify() override final { _potential = nullptr; }
  const std::vector<double>& getEnergyParts() const { return _energyParts; }

 private:
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<DensityMatrixController<SCFMode>> _act;
  std::vector<std::shared_ptr<DensityMatrixController<SCFMode>>> _env;
  std::shared_ptr<GridController> _grid;
  Functional _functional;
  double _thr;
  int _func;
  std::shared_ptr<EnvWatcher> _watcher;
  std::unique_ptr<FockMatrix> _potential;
  std::vector<double> _energyParts;
  double _energy = 0.0;
  bool _envFrozen = false;
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
        bc.push_back(d->getBasisController()->handle(*_dev));
        pc.push_back(d->getDensityMatrix().data());
      }
      double e[2] = {0.0, 0.0};
      _dev->check(sxc_build_ab(_dev->get(), _grid->handle(*_dev), _func, detail::nspin<SCFMode>(), _basisA->handle(*_dev),
                               _basisB->handle(*_dev), (int)bc.size(), bc.data(), pc.data(), _thr, V->data(), e));
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
        be.push_back(d->getBasisController()->handle(*_dev));
        pe.push_back(d->getDensityMatrix().data());
      }
      _dev->check(sxc_build_ab_nadd(_dev->get(), _grid->handle(*_dev), _func, detail::nspin<SCFMode>(), _basisA->handle(*_dev),
                                    _basisB->handle(*_dev), _act->getBasisController()->handle(*_dev),
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
    const int g = _grid->handle(*_dev);
    const int nspin = detail::nspin<SCFMode>();
    const bool embedded = !naddXCFunc.basicFunctionals.empty() || !naddKinFunc.basicFunctionals.empty();
    std::vector<int> bc;
    std::vector<const double*> pc;
    for (auto& d : _dMats) {
      bc.push_back(d->getBasisController()->handle(*_dev));
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
  ~Kernel() {
    for (int k : _sub) sxc_kernel_destroy(_dev->get(), k);
    if (_tot >= 0) sxc_kernel_destroy(_dev->get(), _tot);
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
    if (I == J) k.push_back(_sub[I]);
    return k;
  }
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
  bool _gga;
  std::vector<int> _sub;
  int _tot = -1;
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
    _dev->check(sxc_kernel_response_copy(_dev->get(), _kernel->getGridController()->handle(*_dev), 1));
    _supersystem = true;
  }
  std::vector<Matrix> calcF(unsigned I, unsigned J, const std::vector<Matrix>& densityMatrices) {
    if (I != J && _supersystem) return {};  // already inside the supersystem contraction (:142-145)
    std::vector<int> k = _kernel->stores(I, J);
    if (_supersystem) {  // the total-density part is in the saved supersystem response: add the subsystem store only
      k = {k.back()};
      _dev->check(sxc_kernel_response_copy(_dev->get(), _kernel->getGridController()->handle(*_dev), 0));
    }
    if (k.empty()) return {};
    contract(J, k, densityMatrices, _supersystem);
    const int nb = (int)_kernel->getBasisController(I)->getNBasisFunctions();
    const size_t nmat = densityMatrices.size();
    std::vector<double> flat(nmat * (size_t)nb * nb);
    _dev->check(sxc_kernel_integrate(_dev->get(), _kernel->getGridController()->handle(*_dev),
                                     _kernel->getBasisController(I)->handle(*_dev), flat.data()));
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
    _dev->check(sxc_kernel_contract(_dev->get(), _kernel->getGridController()->handle(*_dev),
                                    _kernel->getBasisController(J)->handle(*_dev), (int)stores.size(), stores.data(), mode(),
                                    (int)D.size() / nspin, flat.data(), accumulate ? 1 : 0));
  }
  std::shared_ptr<B200::XCDevice> _dev;
  std::shared_ptr<Kernel<KernelMode>> _kernel;
  bool _supersystem = false;
};

}  // namespace Serenity
#endif
