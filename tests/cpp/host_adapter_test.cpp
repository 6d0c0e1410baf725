// Drives the C++ adapter (serenity_b200/host/serenity_xc_adapter.h) like Serenity's potential tests drive the reference
// classes (potentials/FuncPotential_test.cpp, NAddFuncPotential_test.cpp): build, cached second call, density change ->
// notify -> rebuild.  Inputs come from a flat binary file written by tests/test_cpp_host.py; results go back the same way.
//   host_adapter_test --expect-no-device          exit 0 iff constructing the device throws SerenityError
//   host_adapter_test <in.bin> <out.bin> [ngpu]
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "../../serenity_b200/host/serenity_xc_adapter.h"

using namespace Serenity;
using R = Options::SCF_MODES;

template <class T>
static std::vector<T> rd(std::ifstream& f) {
  int64_t n = 0;
  f.read(reinterpret_cast<char*>(&n), 8);
  std::vector<T> v((size_t)n);
  f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T));
  return v;
}
static void wr(std::ofstream& f, const double* p, int64_t n) {
  f.write(reinterpret_cast<const char*>(&n), 8);
  f.write(reinterpret_cast<const char*>(p), n * 8);
}
static std::shared_ptr<BasisController> readBasis(std::ifstream& f) {
  ShellTable t;
  t.l = rd<int>(f);
  t.pure = rd<int>(f);
  t.nprim = rd<int>(f);
  t.firstBf = rd<int>(f);
  t.centre = rd<double>(f);
  t.alpha = rd<double>(f);
  t.coeff = rd<double>(f);
  t.normfac = rd<double>(f);
  t.atomOfBf = rd<int>(f);
  t.nAtoms = t.atomOfBf.empty() ? 0 : *std::max_element(t.atomOfBf.begin(), t.atomOfBf.end()) + 1;
  const int nbf = (int)t.normfac.size();
  return std::make_shared<BasisController>(std::move(t), nbf);
}
static DensityMatrix readMatrix(std::ifstream& f, int nb) {
  DensityMatrix P(nb, nb);
  P.values = rd<double>(f);
  return P;
}

int main(int argc, char** argv) {
  if (argc == 2 && std::strcmp(argv[1], "--expect-no-device") == 0) {
    try {
      B200::XCDevice dev(0);
    } catch (const SerenityError& e) {
      std::cout << "SerenityError: " << e.what() << "\n";
      return 0;
    }
    return 1;
  }
  if (argc == 3 && std::strcmp(argv[1], "--resolve") == 0) {  // host-only: CompositeFunctionals::resolveFunctional
    try {
      double hfx = 0.0;
      Functional f = CompositeFunctionals::resolveFunctional(argv[2], &hfx);
      std::cout << argv[2] << " hfx=" << hfx;
      for (size_t i = 0; i < f.basicFunctionals.size(); ++i) std::cout << " " << f.basicFunctionals[i] << ":" << f.mixingFactors[i];
      std::cout << "\n";
      return 0;
    } catch (const SerenityError& e) {
      std::cout << "SerenityError: " << e.what() << "\n";
      return 3;
    }
  }
  if (argc != 3 && argc != 4) return 2;
  try {
    std::ifstream in(argv[1], std::ios::binary);
    std::ofstream out(argv[2], std::ios::binary);
    auto dev = std::make_shared<B200::XCDevice>(0);
    std::vector<double> xyz = rd<double>(in);  // (function arguments have no evaluation order: read first)
    std::vector<double> wts = rd<double>(in);
    auto grid = std::make_shared<GridController>(std::move(xyz), std::move(wts));
    Functional xc{rd<int>(in), rd<double>(in)};
    Functional kin{rd<int>(in), rd<double>(in)};
    auto basisA = readBasis(in);
    auto basisB = readBasis(in);
    const int nA = (int)basisA->getNBasisFunctions(), nB = (int)basisB->getNBasisFunctions();
    auto dA = std::make_shared<DensityMatrixController<R::RESTRICTED>>(basisA, readMatrix(in, nA));
    auto dB = std::make_shared<DensityMatrixController<R::RESTRICTED>>(basisB, readMatrix(in, nB));
    DensityMatrix PA2 = readMatrix(in, nA);
    Matrix DA = readMatrix(in, nA);  // trial densities of the LR-TDDFT sigma vector (row f-4)
    Matrix DB = readMatrix(in, nB);

    // KS-DFT potential of subsystem A on the common grid
    auto pot = std::make_shared<FuncPotential<R::RESTRICTED>>(dev, dA, grid, xc);
    pot->registerSensitivity(pot);
    FockMatrix& V1 = pot->getMatrix();
    if (&pot->getMatrix() != &V1) throw SerenityError("getMatrix() must return the cached matrix until notify()");
    wr(out, V1.data(), (int64_t)nA * nA);
    const double E1 = pot->getEnergy(dA->getDensityMatrix());
    wr(out, &E1, 1);
    // non-additive potentials (freeze-and-thaw: XC and kinetic objects, NAddFuncPotential.cpp:192)
    auto naddXC = std::make_shared<NAddFuncPotential<R::RESTRICTED>>(
        dev, dA, std::vector<std::shared_ptr<DensityMatrixController<R::RESTRICTED>>>{dB}, grid, xc);
    naddXC->registerSensitivity(naddXC);
    auto naddKin = std::make_shared<NAddFuncPotential<R::RESTRICTED>>(
        dev, dA, std::vector<std::shared_ptr<DensityMatrixController<R::RESTRICTED>>>{dB}, grid, kin);
    naddKin->registerSensitivity(naddKin);
    wr(out, naddXC->getMatrix().data(), (int64_t)nA * nA);
    double e = naddXC->getEnergy(dA->getDensityMatrix());
    wr(out, &e, 1);
    wr(out, naddKin->getMatrix().data(), (int64_t)nA * nA);
    e = naddKin->getEnergy(dA->getDensityMatrix());
    wr(out, &e, 1);
    // new active density: every potential that depends on it is invalidated and rebuilt (environment stays frozen)
    dA->setDensityMatrix(PA2);
    wr(out, pot->getMatrix().data(), (int64_t)nA * nA);
    e = pot->getEnergy(PA2);
    wr(out, &e, 1);
    wr(out, naddXC->getMatrix().data(), (int64_t)nA * nA);
    e = naddXC->getEnergy(PA2);
    wr(out, &e, 1);
    // XC nuclear gradient of the active system (FuncPotential_test.cpp:234-330 pattern)
    Matrix grad = pot->getGeomGradients();
    wr(out, grad.data(), (int64_t)grad.rows() * 3);
    // the XC operator between the two different basis sets from the sum of both densities (ABFuncPotential.cpp:54-160)
    auto abPot = std::make_shared<ABFuncPotential<R::RESTRICTED>>(
        dev, basisA, basisB, grid, std::vector<std::shared_ptr<DensityMatrixController<R::RESTRICTED>>>{dA, dB}, xc);
    abPot->registerSensitivity(abPot);
    Matrix& Vab = abPot->getMatrix();
    if (&abPot->getMatrix() != &Vab) throw SerenityError("ABFuncPotential::getMatrix() must cache");
    wr(out, Vab.data(), (int64_t)nA * nB);
    // the non-additive XC operator between the two basis sets (ABNAddFuncPotential.cpp:66-176)
    auto abNadd = std::make_shared<ABNAddFuncPotential<R::RESTRICTED>>(
        dev, dA, basisA, basisB, std::vector<std::shared_ptr<DensityMatrixController<R::RESTRICTED>>>{dB}, grid, xc);
    abNadd->registerSensitivity(abNadd);
    wr(out, abNadd->getMatrix().data(), (int64_t)nA * nB);
    // gradient of the non-additive XC potential over the active atoms (NAddFuncPotential.cpp:329-493)
    Matrix gradN = naddXC->getGeomGradients();
    wr(out, gradN.data(), (int64_t)gradN.rows() * 3);
    // subsystem-TDDFT kernel sigma vector of subsystem A (Kernel.cpp:686-747, KernelSigmavector.cpp:60-252): supersystem
    // contraction of both subsystems' trial densities with the non-additive kernel, then the intra-subsystem part
    {
      using DMC = DensityMatrixController<R::RESTRICTED>;
      auto kernel = std::make_shared<Kernel<R::RESTRICTED>>(dev, grid, std::vector<std::shared_ptr<DMC>>{dA, dB},
                                                            std::vector<Functional>{xc, xc}, xc, kin);
      KernelSigmavector<R::RESTRICTED> sigma(dev, kernel);
      sigma.contractSupersystemDensity({{DA}, {DB}});
      if (!sigma.calcF(0, 1, {DB}).empty()) throw SerenityError("calcF(I != J) must defer to the supersystem contraction");
      std::vector<Matrix> F = sigma.calcF(0, 0, {DA});
      wr(out, F[0].data(), (int64_t)nA * nA);
      std::vector<Matrix> FB = sigma.calcF(1, 1, {DB});  // every subsystem starts from the same supersystem contraction
      wr(out, FB[0].data(), (int64_t)nB * nB);
      // isolated system, two vectors at once
      auto iso = std::make_shared<Kernel<R::RESTRICTED>>(dev, grid, std::vector<std::shared_ptr<DMC>>{dA}, std::vector<Functional>{xc});
      KernelSigmavector<R::RESTRICTED> sigmaIso(dev, iso);
      std::vector<Matrix> F2 = sigmaIso.calcF(0, 0, {DA, PA2});
      wr(out, F2[0].data(), (int64_t)nA * nA);
      wr(out, F2[1].data(), (int64_t)nA * nA);
      std::vector<double> pp = iso->getPP(0, 0, 128, 256);
      wr(out, pp.data(), 128);
    }
    // stage-level classes: density on the grid -> functional data -> grid to matrix reproduce FuncPotential::getMatrix
    {
      DensityOnGridCalculator dens(dev, basisA, grid);
      DensityOnGrid d = dens.calcDensityAndGradientOnGrid(PA2);
      FunctionalLibrary flib(dev, grid);
      FunctionalData fd = flib.calcData(xc, d);
      ScalarOperatorToMatrixAdder adder(dev, basisA, grid);
      Matrix Vst(nA, nA);
      adder.addScalarOperatorToMatrix(Vst, fd.dFdRho, fd.dFdGradRhoX, fd.dFdGradRhoY, fd.dFdGradRhoZ);
      wr(out, Vst.data(), (int64_t)nA * nA);
      wr(out, &fd.energy, 1);
    }
    // ---- round-2 additions ------------------------------------------------------------------------------------------
    using DMC = DensityMatrixController<R::RESTRICTED>;
    const int ngpu = argc == 4 ? std::atoi(argv[3]) : 1;
    Matrix BtoA(nB, nA);
    BtoA.values = rd<double>(in);
    {
      // getLinearizedEnergy (NAddFuncPotential.cpp:180-189) of the cached non-additive XC potential
      const double lin = naddXC->getLinearizedEnergy(PA2, 0.5);
      wr(out, &lin, 1);
      // second constructor (NAddFuncPotential.cpp:105-176): subsystem B treated exactly, projected into the active basis; no
      // approximately treated environment
      auto comb = std::make_shared<NAddFuncPotential<R::RESTRICTED>>(dev, dA, std::vector<std::shared_ptr<DMC>>{dB},
                                                                     std::vector<std::shared_ptr<Matrix>>{std::make_shared<Matrix>(BtoA)},
                                                                     std::vector<std::shared_ptr<DMC>>{dB}, grid, xc);
      comb->registerSensitivity(comb);
      wr(out, comb->getMatrix().data(), (int64_t)nA * nA);
      const double ec = comb->getEnergy(PA2);
      wr(out, &ec, 1);
      // the bundle: both non-additive objects in one device pass; each keeps its own matrix and energy
      auto nx = std::make_shared<NAddFuncPotential<R::RESTRICTED>>(dev, dA, std::vector<std::shared_ptr<DMC>>{dB}, grid, xc);
      auto nk = std::make_shared<NAddFuncPotential<R::RESTRICTED>>(dev, dA, std::vector<std::shared_ptr<DMC>>{dB}, grid, kin);
      nx->registerSensitivity(nx);
      nk->registerSensitivity(nk);
      FDEPotentials<R::RESTRICTED> bundle(nx, nk);
      Matrix Fsum = bundle.getNAddFockMatrix();
      wr(out, Fsum.data(), (int64_t)nA * nA);
      wr(out, nx->getMatrix().data(), (int64_t)nA * nA);
      wr(out, nk->getMatrix().data(), (int64_t)nA * nA);
      const double ex = nx->getEnergy(PA2), ek = nk->getEnergy(PA2);
      wr(out, &ex, 1);
      wr(out, &ek, 1);
      // geometry steps: new grid and basis controllers every step, the old ones release their device copies
      for (int step = 0; step < 4; ++step) {
        auto g2 = std::make_shared<GridController>(grid->getGridPoints(), grid->getWeights());
        auto p2 = std::make_shared<FuncPotential<R::RESTRICTED>>(dev, dA, g2, xc);
        if (std::abs(p2->getEnergy(PA2) - pot->getEnergy(PA2)) > 1e-12) throw SerenityError("rebuilt grid gives another energy");
      }
    }
    {
      // mixed exact / approximate embedding (Kernel::calculateDerivativesMixedEmbedding, Kernel.cpp:752-888): subsystem A embedded
      // exactly (LEVELSHIFT), B approximately; the three kinds of store and the rule that picks them for a pair (I, J)
      using KM = Options::KIN_EMBEDDING_MODES;
      Kernel<R::RESTRICTED> mixed(dev, grid, std::vector<std::shared_ptr<DMC>>{dA, dB}, std::vector<Functional>{xc, xc},
                                  std::vector<KM>{KM::LEVELSHIFT, KM::NADD_FUNC}, /*naddXCExact*/ kin, /*naddXCApprox*/ xc, kin);
      if (!mixed.mixedEmbeddingUsed() || mixed.stores(0, 0).size() != 3 || mixed.stores(0, 1).size() != 1 || mixed.stores(1, 1).size() != 2)
        throw SerenityError("mixed-embedding kernel: wrong store selection");
      for (auto ij : {std::make_pair(0u, 0u), std::make_pair(0u, 1u), std::make_pair(1u, 1u)}) {
        std::vector<double> pp = mixed.getPP(ij.first, ij.second, 128, 256);
        wr(out, pp.data(), 128);
      }
    }
    if (ngpu > 1) {
      // one process, ngpu GPUs: the same potential classes on a group device (sxc_group: one worker thread and context per GPU,
      // ncclAllReduce inside the library)
      std::vector<int> devices(ngpu);
      for (int i = 0; i < ngpu; ++i) devices[i] = i;
      auto gdev = std::make_shared<B200::XCDevice>(devices);
      auto ggrid = std::make_shared<GridController>(grid->getGridPoints(), grid->getWeights());
      auto gA = std::make_shared<BasisController>(*basisA);
      auto gB = std::make_shared<BasisController>(*basisB);
      auto gdA = std::make_shared<DMC>(gA, PA2);
      auto gdB = std::make_shared<DMC>(gB, dB->getDensityMatrix());
      auto gpot = std::make_shared<FuncPotential<R::RESTRICTED>>(gdev, gdA, ggrid, xc);
      gpot->registerSensitivity(gpot);
      wr(out, gpot->getMatrix().data(), (int64_t)nA * nA);
      const double eg = gpot->getEnergy(PA2);
      wr(out, &eg, 1);
      auto gx = std::make_shared<NAddFuncPotential<R::RESTRICTED>>(gdev, gdA, std::vector<std::shared_ptr<DMC>>{gdB}, ggrid, xc);
      auto gk = std::make_shared<NAddFuncPotential<R::RESTRICTED>>(gdev, gdA, std::vector<std::shared_ptr<DMC>>{gdB}, ggrid, kin);
      FDEPotentials<R::RESTRICTED> gbundle(gx, gk);
      Matrix gF = gbundle.getNAddFockMatrix();
      wr(out, gF.data(), (int64_t)nA * nA);
      Matrix gg = gpot->getGeomGradients();
      wr(out, gg.data(), (int64_t)gg.rows() * 3);
      const double n = (double)gdev->nGPUs();
      wr(out, &n, 1);
    }
    // error convention: SerenityError, as the reference throws (here: a functional id the library does not implement)
    bool threw = false;
    try {
      FuncPotential<R::RESTRICTED> bad(dev, dA, grid, Functional{{9999}, {1.0}});
    } catch (const SerenityError&) {
      threw = true;
    }
    if (!threw) throw SerenityError("an unsupported functional must throw SerenityError");
  } catch (const std::exception& e) {
    std::cerr << "host_adapter_test failed: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
