/*
 * oracle.h - CPU restatement of Serenity's DFT numerical-integration path (TEST INFRASTRUCTURE).
 *
 * This is the parity oracle and the timed "host-core CPU baseline" of the serenity_b200 build.
 * It is NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product path (serenity_b200/csrc) never links or calls it.
 *
 * Parity status: the reference cannot be compiled here (Eigen3, libint2, xcfun, libxc, HDF5 absent), so this
 * is a restatement.  Rows 8a-1, 8a-2, 8a-5 are PINNED against the reference's own unit-test vectors
 * (the JSON files under tests/golden/).  The functional arithmetic (8a-4) lives in un-vendored xcfun (qcserenity/xcfun,
 * no tag pinned, cmake/ImportXCFun.cmake:13-17) / libxc 6.1.0: its published formulas are restated in
 * oracle_functionals.c and are "parity unpinned" at the 1e-9 Eh level (no per-functional KAT exists in the
 * reference); see DESIGN.md.
 *
 * Every function cites the reference file:line (relative to /root/reference/src) it follows.
 */
#ifndef ORACLE_H
#define ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* Shell table, as BasisController/Shell hand it to BasisFunctionOnGridController
 * (basis/Shell.h:179-181 coefficients = libint-renormalised contr[0].coeff; basis/Shell.cpp:37-47 norm factors). */
typedef struct {
  int nshell;
  int nbf;               /* number of (extended) basis functions */
  const int* l;          /* [nshell] angular momentum */
  const int* pure;       /* [nshell] 1 = spherical, 0 = Cartesian */
  const int* nprim;      /* [nshell] */
  const int* prim_off;   /* [nshell] offset into alpha/coeff */
  const int* first_bf;   /* [nshell] extendedIndex(shell) */
  const double* centre;  /* [3*nshell] bohr */
  const double* alpha;   /* primitives */
  const double* coeff;   /* renormalised contraction coefficients */
  const double* normfac; /* [nbf] Cartesian component factors (1 for spherical) */
} orc_basis;

/* Grid as GridController::getGridPoints()/getWeights() deliver it (grid/GridController.cpp:31-50):
 * xyz interleaved (Matrix3Xd column-major), consecutive `blocksize` points form a block. */
typedef struct {
  long npts;
  const double* xyz; /* [3*npts] */
  const double* w;   /* [npts]   */
  int blocksize;     /* settings grid.blocksize, default 128 */
} orc_grid;

/* BASIC_FUNCTIONALS enum values (dft/functionals/BasicFunctionals.h:39-...) of the supported kernels */
enum {
  ORC_NONE = 0,
  ORC_X_SLATER = 2,
  ORC_C_VWN = 45,
  ORC_K_TF = 66,
  ORC_X_B88 = 80,
  ORC_X_B88_CORR = 81,
  ORC_X_PBE = 135,
  ORC_C_LYP = 184,
  ORC_C_P86 = 193,
  ORC_C_PBE = 197,
  ORC_K_PW91 = 283,
  ORC_K_LLP = 286
};

typedef struct {
  int ncomp;
  const int* id;     /* BASIC_FUNCTIONALS values */
  const double* mix; /* mixing factors */
} orc_functional;

/* phase timers (seconds), labels as the reference's Timings (misc/Timing.cpp:75-106) */
typedef struct {
  double basis_on_grid;   /* "Tech. -    Basis On Grid Eval." (inside the two phases below, thread 0 only) */
  double density_on_grid; /* "Tech. -  Density On Grid Eval." */
  double functional;      /* "Tech. - XCFun Functional Eval." */
  double grid_to_matrix;  /* "Tech. -    Grid to Matrix Int." */
  double total;           /* "Active System - Functional Pot." */
} orc_timings;

int orc_nblocks(const orc_grid* g);
int orc_max_threads(void);
void orc_set_threads(int n);
/* optional BLAS back end for the dense products: a cblas_dgemm-compatible function pointer (NULL: built-in loops) */
void orc_set_dgemm(void* cblas_dgemm_fn);
int orc_has_dgemm(void);
/* phase probe of OpenMP thread 0: seconds in basis-function evaluation, seconds in the dense products, their flops */
void orc_probe_reset(void);
void orc_probe_get(double* out3);
int orc_functional_is_gga(const orc_functional* f);

/* 8a-1  BasisFunctionOnGridController::calculateBasisFunctionData (data/grid/BasisFunctionOnGridController.cpp:150-1105).
 * Outputs are n x nbf column-major (index mu*n + p) like functionValues; deriv = 0/1/2.
 * negligible[nbf]; values of negligible functions are written as 0. Returns the block size n. */
int orc_basis_block(const orc_basis* b, const orc_grid* g, double radial_thr, int deriv, int block, double* val,
                    double* dx, double* dy, double* dz, double* hxx, double* hxy, double* hxz, double* hyy,
                    double* hyz, double* hzz, int* negligible, double* centre);

/* 8a-2  MatrixOperatorToGridTransformer::transform (data/grid/MatrixOperatorToGridTransformer.cpp:37-198)
 * via DensityOnGridCalculator (DensityOnGridCalculator.cpp:55-65). P nbf x nbf column-major.
 * grad/hess pointers may be NULL. nonneg[nblocks] receives the non-negligible block flags (:190-197). */
void orc_density_on_grid(const orc_basis* b, const orc_grid* g, double radial_thr, const double* P, double* rho,
                         double* gx, double* gy, double* gz, double* hess6 /* [6*npts] xx,xy,xz,yy,yz,zz or NULL */,
                         int* nonneg);

/* 8a-4  FunctionalLibrary::calcData(GRADIENTS) through XCFun::calcData (dft/functionals/wrappers/XCFun.cpp:39-159),
 * RESTRICTED. Blocks of the literal 128 points, block skip at sum|rho| < n*1e-12, zero below rho < 1e-14.
 * Outputs zero-initialised here. gx..gz / dFdG* may be NULL for LDA. Returns E = sum_p w_p F_p (XCFun.cpp:752-765). */
double orc_functional_on_grid(const orc_functional* f, long npts, const double* w, const double* rho,
                              const double* gx, const double* gy, const double* gz, double* epuv, double* dFdRho,
                              double* dFdGx, double* dFdGy, double* dFdGz);

/* pointwise kernel (no thresholds): F, dF/drho, dF/dsigma of one basic functional, closed shell */
int orc_basic_functional(int id, double rho, double sigma, double* F, double* vrho, double* vsigma);

/* 8a-5  ScalarOperatorToMatrixAdder::addScalarOperatorToMatrix (data/grid/ScalarOperatorToMatrixAdder.cpp:52-116,
 * addBlock :179-303), same basis on both sides. gx==NULL selects the LDA variant. V (nbf x nbf col-major) is
 * ADDED to, as in the reference. */
void orc_scalar_to_matrix(const orc_basis* b, const orc_grid* g, double radial_thr, double block_ave_thr,
                          const double* v, const double* gx, const double* gy, const double* gz, double* V);
/* f-4: two-basis variant (basis A != basis B), ScalarOperatorToMatrixAdder.cpp:216-220 / :286-300; V is nbf_A x nbf_B */
void orc_scalar_to_matrix_ab(const orc_basis* bA, const orc_basis* bB, const orc_grid* g, double radial_thr,
                             double block_ave_thr, const double* v, const double* gx, const double* gy, const double* gz,
                             double* V);

/* 8a-6  FuncPotential::getMatrix / getEnergy (potentials/FuncPotential.cpp:74-111), RESTRICTED.
 * V is overwritten. nelec = sum_p w_p rho_p (gridAccuracyCheck, DensityMatrixDensityOnGridController.cpp:152-160). */
int orc_build_xc(const orc_basis* b, const orc_grid* g, const orc_functional* f, double radial_thr,
                 double block_ave_thr, const double* P, double* V, double* E, double* nelec, orc_timings* t);

/* 8a-7  NAddFuncPotential::getMatrix / getEnergy (potentials/NAddFuncPotential.cpp:192-300, :502-516) with
 * SupersystemDensityOnGridController::updateData (data/grid/SupersystemDensityOnGridController.cpp:95-193).
 * V_A (active basis) overwritten; E_nadd = E[tot] - E[act] - sum_env E[env]; E_parts = {E_tot, E_act, E_env...}. */
int orc_build_nadd(const orc_basis* bA, const double* PA, int nenv, const orc_basis* const* bE,
                   const double* const* PE, const orc_grid* g, const orc_functional* f, double radial_thr,
                   double block_ave_thr, double* VA, double* E_nadd, double* E_parts);

/* row f-3  FuncPotential<RESTRICTED / UNRESTRICTED>::getGeomGradients (potentials/FuncPotential.cpp:114-239): the
 * reference's double loop over the significant functions of every block, with second basis-function derivatives for GGAs.
 * nspin = 1: Pa = total density matrix, Pb ignored.  grad: natoms x 3 column-major, overwritten. */
int orc_xc_gradient(const orc_basis* b, const orc_grid* g, const orc_functional* f, double radial_thr, int nspin,
                    const double* Pa, const double* Pb, int natoms, const int* atom_of_bf, double* grad);
/* NAddFuncPotential::getGeomGradients (NAddFuncPotential.cpp:329-493), RESTRICTED; grad [natoms x 3] column-major */
int orc_nadd_gradient(const orc_basis* bA, const double* PA, int nenv, const orc_basis* const* bE, const double* const* PE,
                      const orc_grid* g, const orc_functional* f, double radial_thr, int natoms, const int* atom_of_bf,
                      double* grad);

/* ---- UNRESTRICTED (SCFMode = UNRESTRICTED: alpha/beta pairs, data/SpinPolarizedData.h) ------------------------ */
/* pointwise spin-polarised kernel: F and dF/d(rho_a, rho_b, s_aa, s_ab, s_bb) */
int orc_basic_functional_u(int id, double ra, double rb, double gaa, double gab, double gbb, double* F, double* d5);
/* 8a-4 unrestricted branch of XCFun::calcData (XCFun.cpp:100-112, :129-153): rho[2][npts], grad[2][3][npts] (may be
 * NULL for LDA) -> epuv[npts], dFdRho[2][npts], dFdGrad[2][3][npts].  Block skip only if BOTH spin densities are
 * below the block threshold (:133-137). */
double orc_functional_on_grid_u(const orc_functional* f, long npts, const double* w, const double* rho,
                                const double* grad, double* epuv, double* dFdRho, double* dFdGrad);
/* 8a-6 FuncPotential<UNRESTRICTED>::getMatrix: P = {P_alpha, P_beta}, V = {V_alpha, V_beta}; the grid -> matrix step
 * runs per spin with its own block-average test (ScalarOperatorToMatrixAdder.cpp:262-268 inside for_spin). */
int orc_build_xc_u(const orc_basis* b, const orc_grid* g, const orc_functional* f, double radial_thr,
                   double block_ave_thr, const double* Pa, const double* Pb, double* Va, double* Vb, double* E,
                   double* nelec);
/* 8a-7 NAddFuncPotential<UNRESTRICTED>: E_parts = {E_tot, E_act, E_env...} */
int orc_build_nadd_u(const orc_basis* bA, const double* PAa, const double* PAb, int nenv, const orc_basis* const* bE,
                     const double* const* PEa, const double* const* PEb, const orc_grid* g, const orc_functional* f,
                     double radial_thr, double block_ave_thr, double* VAa, double* VAb, double* E_nadd, double* E_parts);

/* ---- row f-4: second functional derivatives and the LR-TDDFT kernel contraction (oracle_kernel2.cpp) ------------- */
/* F, d5 and the 5 x 5 Hessian (row-major) of one basic functional w.r.t. (rho_a, rho_b, s_aa, s_ab, s_bb) */
int orc_basic_functional_d2(int id, double ra, double rb, double gaa, double gab, double gbb, double* F, double* d5,
                            double* h25);
/* test hook: density screen of storeDerivatives (the reference hard-codes 1e-8, Kernel.cpp:496, :606) */
void orc_kernel_set_screen(double thr);
/* Kernel<RESTRICTED>::storeDerivatives (postHF/LRSCF/Kernel/Kernel.cpp:476-520): store [10][npts] (pp, pg x y z,
 * gg xx xy xz yy yz zz; [1][npts] if !store_gga) += sign * d2F, then zeroed where rho < 1e-8 */
void orc_kernel_store_r(const orc_functional* f, long npts, const double* rho, const double* gx, const double* gy,
                        const double* gz, double sign, int store_gga, double* store);
/* Kernel<UNRESTRICTED>::storeDerivatives (:523-683): rho [2][npts], grad [2][3][npts] (NULL for LDA), store [33][npts] =
 * pp aa ab bb | pg {x,y,z} x {aa,ab,ba,bb} | gg {xx,xy,xz,yy,yz,zz} x {aa,ab,bb} ([3][npts] if !store_gga) */
void orc_kernel_store_u(const orc_functional* f, long npts, const double* rho, const double* grad, double sign,
                        int store_gga, double* store);
/* KernelSigmavector::contractKernel + contractBlock (postHF/LRSCF/Sigmavectors/KernelSigmavector.cpp:254-311, :360-497) for
 * one trial vector; mode 0 singlet / 1 triplet (UNRESTRICTED store) / 2 UNRESTRICTED; resp [4 * nspin][N] is added to */
void orc_kernel_contract(const orc_basis* b, const orc_grid* g, double radial_thr, double block_ave_thr, int mode, int gga,
                         const double* store, const double* D, double* resp);
/* KernelSigmavector::numericalIntegration (:313-358) + thread sum and F += F^T (:236-249) for one trial vector */
void orc_kernel_integrate(const orc_basis* b, const orc_grid* g, double radial_thr, double block_ave_thr, int gga, int nspin,
                          const double* resp, double* F);

/* ---- INT8-slice ("Ozaki") reference of the contractions (ozaki.c): groundwork for the tcgen05 kind::i8 path ------- */
/* rows of X [rows][cols] -> k signed 7-bit slices S [k][rows][cols] and one exponent per row */
void orc_ozaki_slice_rows(const double* X, int rows, int cols, int k, signed char* S, int* e);
/* acc [k][m][n] (INT32, exact): acc_d[a][b] = sum_{i+j=d} sum_c A_i[a][c] B_j[b][c]  (both operands K-contiguous) */
void orc_ozaki_gemm_i32(const signed char* A, const signed char* B, int k, int m, int n, int K, int* acc);
/* C[m][n] = 2^(eA_m + eB_n) sum_d 128^-(d+2) acc_d */
void orc_ozaki_combine(const int* acc, int k, int m, int n, const int* eA, const int* eB, double* C);

#ifdef __cplusplus
}
#endif
#endif
