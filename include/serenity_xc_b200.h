/*
 * serenity_xc_b200.h - C ABI of libserenity_xc_b200.so: the B200 (sm_100a) XC / embedding-potential build.
 *
 * Drop-in boundary directly underneath Serenity's unchanged Potential interface
 * (src/potentials/Potential.h:43-86).  An adapter replaces the bodies of
 *   FuncPotential<SCFMode>::getMatrix / getEnergy          src/potentials/FuncPotential.cpp:67-111
 *   NAddFuncPotential<SCFMode>::getMatrix / getEnergy      src/potentials/NAddFuncPotential.cpp:192-326
 * with calls to this library (see INTEGRATION.md and serenity_b200/host/).  All pointers are plain host
 * (or, for the *_device entry points, CUDA device) pointers owned by the caller and only touched during the
 * call; device state (grid, shell tables, screening plan, work buffers) persists in the context across SCF
 * iterations.  Functions return 0 or a negative sxc_status and never throw; sxc_last_error() gives the text
 * the adapter puts into SerenityError (src/misc/SerenityError.h:36).
 *
 * Conventions (identical to the reference):
 *   - grid points 3 x N column-major = xyz interleaved, as Eigen::Matrix3Xd (src/grid/Grid.h:39-81); weights [N];
 *     a block = `blocksize` consecutive points (BasisFunctionOnGridController.cpp:170-179)
 *   - matrices nb x nb column-major FP64 (Eigen::MatrixXd; src/data/matrices/MatrixInBasis.h:57)
 *   - RESTRICTED P is the total density matrix (occupation 2)
 *   - functional components are BASIC_FUNCTIONALS enum values (src/dft/functionals/BasicFunctionals.h:39-...)
 *     with mixing factors, as Functional::getBasicFunctionals()/getMixingFactors() return them
 *     (src/dft/functionals/wrappers/XCFun.cpp:721-745)
 * There is no CPU fallback: every entry point that computes fails with SXC_ERR_CUDA without a usable device.
 */
#ifndef SERENITY_XC_B200_H
#define SERENITY_XC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sxc_ctx sxc_ctx;
typedef struct sxc_shell_table sxc_shell_table; /* host-side shell table built from a basis-set file (row f-2) */

typedef enum {
  SXC_OK = 0,
  SXC_ERR_INVALID = -1,     /* bad argument / handle */
  SXC_ERR_CUDA = -2,        /* CUDA runtime error (text in sxc_last_error) */
  SXC_ERR_NOMEM = -3,       /* device or host allocation failed */
  SXC_ERR_UNSUPPORTED = -4  /* valid in the reference, not implemented here (e.g. l > 6, meta-GGA ids) */
} sxc_status;

/* supported BASIC_FUNCTIONALS (XCFun aliases: BasicFunctionals.cpp:1663-1790) */
enum {
  SXC_NONE = 0,
  SXC_X_SLATER = 2,     /* slaterx     */
  SXC_C_VWN = 45,       /* vwn5c       */
  SXC_K_TF = 66,        /* tfk         */
  SXC_X_B88 = 80,       /* beckex      */
  SXC_X_B88_CORR = 81,  /* beckecorrx  */
  SXC_X_PBE = 135,      /* pbex        */
  SXC_C_LYP = 184,      /* lypc        */
  SXC_C_P86 = 193,      /* p86c        */
  SXC_C_PBE = 197,      /* pbec        */
  SXC_K_PW91 = 283,     /* pw91k (Lembarki-Chermette) */
  SXC_K_LLP = 286       /* llp91k      */
};

/* kernels of one build, indices into sxc_stats.ms_kernel / n_kernel */
enum {
  SXC_T_SCREEN = 0,   /* k_screen: block prescreening                      (8a-1, :211-255) */
  SXC_T_BASIS = 1,    /* k_basis: phi, grad phi tiles                      (8a-1)           */
  SXC_T_DENSITY = 2,  /* k_density: phi_s P_s DMMA + rho / grad rho sums   (8a-2, 8a-3)     */
  SXC_T_FUNCTIONAL = 3, /* k_functional: LDA/GGA kernels, w F partial sums (8a-4)           */
  SXC_T_FORM_G = 4,   /* k_form_g: weights x potential, block test, G tile (8a-5)           */
  SXC_T_SCATTER = 5,  /* k_scatter: phi^T G + G^T phi DMMA, accumulation   (8a-5)           */
  SXC_T_FINISH = 6,   /* k_mirror + k_reduce_partials                                       */
  SXC_T_ALLREDUCE = 7, /* ncclAllReduce of [V | E | N] (multi-GPU, 8e)                      */
  SXC_T_COUNT = 8
};

/* work/timing counters of the last build on this context (for roofline reporting) */
typedef struct {
  int64_t npts;           /* grid points owned by this context (its shard) */
  int64_t nblocks;        /* blocks owned */
  int64_t sum_s;          /* sum_b s_b       (significant functions) */
  int64_t sum_ns;         /* sum_b n_b s_b   */
  int64_t sum_ns2;        /* sum_b n_b s_b^2 -> F_gemm = 4 * sum_ns2 */
  int64_t sum_ns2_padded; /* same with the padded tile sizes the kernels execute */
  int64_t sum_s2;         /* sum_b s_b^2     -> gather/accumulate bytes 16 * sum_s2 */
  int64_t s_max;
  int64_t nbf;
  int64_t workspace_bytes; /* device bytes of the phi / grad phi tile buffer */
  int32_t nchunks;         /* block chunks the build was pipelined in */
  int32_t kernel_launches; /* kernels launched by the last build */
  /* CUDA-event times (ms) and launch counts per kernel of the last build, summed over chunks and subsystems;
   * filled by the host-buffer entry points always and by the *_device ones after sxc_set_timing(ctx, 1) */
  float ms_kernel[SXC_T_COUNT];
  int32_t n_kernel[SXC_T_COUNT];
  float ms_total;          /* first kernel to last kernel of the build, on the device */
} sxc_stats;

/* ---- context ------------------------------------------------------------------------------------------- */
/* One context per process and GPU (one process per GPU; calls are serialised by the caller, matching the
 * reference's single SCF driver thread).  device = CUDA ordinal. */
int sxc_create(sxc_ctx** ctx, int device);
void sxc_destroy(sxc_ctx* ctx);
const char* sxc_last_error(const sxc_ctx* ctx);
/* run on the caller's stream (cudaStream_t passed as void*); NULL = the context's own (non-blocking) stream;
 * pass cudaStreamLegacy ((void*)0x1) to name the legacy default stream */
int sxc_set_stream(sxc_ctx* ctx, void* cuda_stream);
/* cap of the phi/grad-phi tile buffer in bytes (default: 40 % of free device memory at plan time) */
int sxc_set_workspace_limit(sxc_ctx* ctx, int64_t bytes);
/* Keep the phi / grad phi tiles resident between builds (off by default).  The reference re-evaluates the basis functions on
 * the grid twice per build (MatrixOperatorToGridTransformer.cpp:103, ScalarOperatorToMatrixAdder.cpp:69-70) because host memory
 * cannot hold them; they depend on grid and basis only, which are fixed during an SCF, and a B200 holds them (tetracene 3.7 GB,
 * peptide 25 GB of 180 GB).  With the cache on, sxc_build_xc / sxc_build_nadd skip the screening and basis kernels whenever the
 * workspace still holds the tiles of the same (grid, basis) pair from the previous call (single-chunk plans; any build with
 * another basis, or a new grid / shard, refills the workspace).  Results are bit-identical to the uncached build. */
int sxc_set_tile_cache(sxc_ctx* ctx, int on);
/* One-shot: the next *_device build waits for this cudaEvent_t (passed as void*) before it first reads P - i.e. after
 * the screening and basis kernels - so that the caller's asynchronous upload of P on another stream overlaps with them.
 * (The host-buffer entry points do the same internally.) */
int sxc_set_p_ready_event(sxc_ctx* ctx, void* cuda_event);
/* per-kernel CUDA-event timing of the *_device builds (off by default: events cost a few microseconds each);
 * sxc_get_stats() then synchronises on the last build's events */
int sxc_set_timing(sxc_ctx* ctx, int on);
/* Several contexts of ONE process that write their (identical, all-reduced) result matrix into the same caller buffer: this
 * context copies back only part `part` of `parts` of V in sxc_build_xc / sxc_build_nadd[_multi] (the matrix then crosses N PCIe
 * links at once).  Set by sxc_group_create for its workers; default 0 of 1 = the whole matrix. */
int sxc_set_output_slice(sxc_ctx* ctx, int part, int parts);

/* Optional page-locked host memory (cudaHostAlloc) for the P / V buffers of the host-buffer builds: with it their copies are
 * asynchronous DMA.  With ordinary (pageable) caller memory - what Eigen matrices in Serenity are - transfers of >= 1 MB are
 * staged by the library through its own page-locked buffer with a few host threads (SXC_COPY_THREADS, default 4; uploads behind
 * the basis kernel, downloads streamed in 512 KB pieces so that the memcpy overlaps the DMA); smaller ones are left to the driver. */
void* sxc_host_alloc(size_t bytes);
void sxc_host_free(void* p);

/* ---- multi-GPU ------------------------------------------------------------------ */
/* Grid blocks are sharded over the GPUs of one node, one context (= one process or host thread) per GPU; the only exchange is
 * ONE ncclAllReduce of the build's result buffer [V | E | N ...] (nspin*nb*nb + 2 doubles) over NVLink 5 / NVSwitch, issued by
 * the library on the build's stream right behind its last kernel - the device-side replacement of the serial reduction over
 * per-thread accumulators in ScalarOperatorToMatrixAdder.cpp:73-75 / :108-110.  NCCL is bound at run time (libnccl.so.2).
 *   rank 0:      sxc_comm_unique_id(id);  ship the 128 bytes to the other ranks by any means (MPI, file, torch store)
 *   every rank:  sxc_create(&ctx, local_device);  sxc_comm_init_rank(ctx, rank, world, id);
 * From then on every grid of the context is this rank's shard (sxc_set_grid_shard is applied automatically) and sxc_build_xc /
 * sxc_build_nadd / sxc_build_nadd_multi / sxc_xc_gradient / sxc_nadd_gradient / sxc_build_ab / sxc_build_ab_nadd /
 * sxc_kernel_integrate return the SUM over ranks on every rank; all ranks must make the same calls in the same order.  In the
 * host-buffer builds a rank that does not need the matrix may pass V = NULL (only E and nelec are copied back). */
#define SXC_COMM_ID_BYTES 128
int sxc_comm_unique_id(void* id128);
int sxc_comm_init_rank(sxc_ctx* ctx, int rank, int world, const void* id128);
int sxc_comm_destroy(sxc_ctx* ctx);
/* rank / world of the context's communicator (0 / 1 without one), all-reduces issued so far, NCCL version code; any pointer
 * may be NULL */
int sxc_comm_info(sxc_ctx* ctx, int* rank, int* world, int64_t* collectives, int* nccl_version);

/* ---- one process, N GPUs -------------------------------------------------------------------------------- */
/* The reference calls getMatrix() from its single SCF driver thread; a single-process host gets the multi-GPU build through a
 * group (the form SURVEY.md section 8b sketches as sxc_create(ctx, ngpu, devices)): one context and one host worker thread per
 * GPU, communicators created inside (ncclCommInitRank from every worker), every call below handed to all workers.  Handles
 * returned by the group name the same object on every context.  A group build returns the all-reduced result in the caller's
 * buffer, every context copying back its 1 / ngpu of the matrix over its own PCIe link (sxc_set_output_slice); P / V are caller-owned host buffers as in sxc_build_xc.  ngpu = 1 works (no NCCL needed). */
typedef struct sxc_group sxc_group;
int sxc_group_create(sxc_group** group, int ngpu, const int* devices /* NULL: 0 .. ngpu-1 */);
void sxc_group_destroy(sxc_group* group);
int sxc_group_size(const sxc_group* group);
sxc_ctx* sxc_group_ctx(sxc_group* group, int rank); /* introspection (sxc_get_stats, sxc_comm_info) */
const char* sxc_group_last_error(const sxc_group* group);
int sxc_group_set_grid(sxc_group* group, int64_t npts, const double* xyz, const double* w, int blocksize, int* grid);
int sxc_group_add_basis(sxc_group* group, int nshell, const int* l, const int* pure, const int* nprim, const int* first_bf,
                        const double* centre, const double* alpha, const double* coeff, const double* normfac,
                        double radial_threshold, int* basis);
int sxc_group_set_functional(sxc_group* group, int ncomp, const int* basic_id, const double* mix, int* func);
int sxc_group_release_grid(sxc_group* group, int grid);
int sxc_group_release_basis(sxc_group* group, int basis);
int sxc_group_build_xc(sxc_group* group, int grid, int basis, int func, int nspin, const double* P,
                       double block_ave_threshold, double* V, double* E, double* nelec);
int sxc_group_build_nadd_multi(sxc_group* group, int grid, int nfunc, const int* funcs, int nspin, int basis_act,
                               const double* P_act, int nenv, const int* basis_env, const double* const* P_env, int env_frozen,
                               double block_ave_threshold, int sum_matrices, double* V_act, double* E);
int sxc_group_xc_gradient(sxc_group* group, int grid, int basis, int func, int nspin, const double* P, int natoms,
                          const int* atom_of_bf, double* grad);

/* ---- inputs -------------------------------------------------------------------------------------------- */
/* replaces GridController::getGridPoints()/getWeights() (src/grid/GridController.cpp:31-50); re-upload only
 * on a Grid notify.  blocksize = settings grid.blocksize (128; 1..128 supported). */
int sxc_set_grid(sxc_ctx* ctx, int64_t npts, const double* xyz, const double* w, int blocksize, int* grid);
/* Multi-GPU: this context evaluates only the blocks of shard `rank` out of `world` (contiguous, cost-balanced
 * ranges of the block order, decided when the first basis is paired with the grid).  Partial V / E / N of all
 * ranks are summed by the caller (one all-reduce).  Requires blocksize == 128.  Default rank 0 of 1. */
int sxc_set_grid_shard(sxc_ctx* ctx, int grid, int rank, int world);

/* replaces the shell data BasisFunctionOnGridController reads from BasisController
 * (src/data/grid/BasisFunctionOnGridController.cpp:216-222,274-287): per shell l, spherical flag, primitives,
 * extendedIndex; coeff = libint-renormalised contr[0].coeff (src/basis/Shell.h:179-181); normfac [nbf] =
 * Shell::getNormFactors() for Cartesian shells (1.0 for spherical).  radial_threshold =
 * settings grid.basFuncRadialThreshold (1e-9). */
int sxc_add_basis(sxc_ctx* ctx, int nshell, const int* l, const int* pure, const int* nprim, const int* first_bf,
                  const double* centre /*3*nshell*/, const double* alpha, const double* coeff,
                  const double* normfac, double radial_threshold, int* basis);

/* replaces XCFun::getFunctional (src/dft/functionals/wrappers/XCFun.cpp:721-745); equal definitions share one handle */
int sxc_set_functional(sxc_ctx* ctx, int ncomp, const int* basic_id, const double* mix, int* func);
/* Release a grid (points, per-point work arrays, cached environment density, kernel stores on it), a basis (shell table) and
 * every screening plan that involves them; the handle value may be handed out again.  What ~GridController / ~BasisController
 * trigger in the adapter: geometry steps and regenerated FDE grids do not pile up in HBM.  Functional handles are a few
 * bytes, shared and kept until sxc_destroy (the call only validates the handle). */
int sxc_release_grid(sxc_ctx* ctx, int grid);
int sxc_release_basis(sxc_ctx* ctx, int basis);
int sxc_release_functional(sxc_ctx* ctx, int func);

/* ---- the hot path -------------------------------------------------------------------------------------- */
/* FuncPotential<SCFMode>::getMatrix + getEnergy (src/potentials/FuncPotential.cpp:67-111).
 * nspin = 1 (RESTRICTED): P, V nb x nb.  nspin = 2 (UNRESTRICTED): P = {P_alpha, P_beta} and V = {V_alpha, V_beta}
 * stored back to back (2 nb^2 doubles each, the alpha/beta pair of src/data/SpinPolarizedData.h).
 * V is overwritten, E = sum_p w_p F_p, nelec = sum_p w_p (rho_alpha + rho_beta).  block_ave_threshold = settings
 * grid.blockAveThreshold (1e-11), applied per spin as in the reference.  With a shard set by hand (sxc_set_grid_shard, no
 * communicator) V/E/nelec are this rank's partial sums; with a communicator they are the sums over all ranks and V may be
 * NULL on ranks that do not need the matrix. */
int sxc_build_xc(sxc_ctx* ctx, int grid, int basis, int func, int nspin, const double* P,
                 double block_ave_threshold, double* V, double* E, double* nelec);
/* same with device-resident P and result: d_VEN holds nspin*nb*nb doubles of V followed by E and nelec
 * (the buffer of the single all-reduce).  Asynchronous on the context's stream. */
int sxc_build_xc_device(sxc_ctx* ctx, int grid, int basis, int func, int nspin, const double* d_P,
                        double block_ave_threshold, double* d_VEN);

/* NAddFuncPotential::getMatrix + getEnergy (src/potentials/NAddFuncPotential.cpp:192-326) with the supersystem
 * density of SupersystemDensityOnGridController::updateData (SupersystemDensityOnGridController.cpp:95-193):
 * V_A = scatter of v[rho_A + sum rho_env] - v[rho_A] in the active basis; E[0] = E[rho_tot], E[1] = E[rho_A],
 * E[2+i] = E[rho_env_i]  (E_nadd = E[0] - E[1] - sum E[2+i]).  nspin = 2: every P / V is an {alpha, beta} pair
 * stored back to back.  env_frozen != 0: environment densities on the grid and their energies are kept from the previous
 * call with the same handles AND the same env_frozen value (they are frozen during one FDE SCF,
 * DensityOnGridFactory.cpp:41-48).  The value is the caller's tag of the environment state: the cache lives on the grid, so
 * two NAdd objects on one grid whose environments share basis handles but hold different density matrices must pass
 * different tags, and a caller whose environment density changed passes a new one (the adapter draws them from a counter;
 * the XC and kinetic objects of one FDE iteration see the same environment and share a tag). */
int sxc_build_nadd(sxc_ctx* ctx, int grid, int func, int nspin, int basis_act, const double* P_act, int nenv,
                   const int* basis_env, const double* const* P_env, int env_frozen, double block_ave_threshold,
                   double* V_act, double* E /*[2+nenv]*/);
int sxc_build_nadd_device(sxc_ctx* ctx, int grid, int func, int nspin, int basis_act, const double* d_P_act,
                          int nenv, const int* basis_env, const double* const* d_P_env, int env_frozen,
                          double block_ave_threshold, double* d_VE /* nspin*nbA*nbA + 2 + nenv */);

/* One device pass for the nfunc non-additive functionals of an FDE iteration: FDEPotentials::getFockMatrix
 * (src/potentials/bundles/FDEPotentials.cpp:43-61) adds naddXC->getMatrix() and naddKin->getMatrix(), two NAddFuncPotential
 * objects on the SAME active / environment densities and grid.  rho_act (the DMMA contraction) and sum rho_env are built once;
 * per functional E[k*(2+nenv) + 0] = E_k[rho_tot], [+1] = E_k[rho_act], [+2+i] = E_k[rho_env_i].
 *   sum_matrices != 0: V_act = sum_k V_nadd,k (nspin*nbA*nbA doubles): the potentials are summed on the grid and scattered
 *                      once - exactly what the bundle hands to the Fock matrix;
 *   sum_matrices == 0: V_act = nfunc matrices back to back (every object gets its own; one scatter each).
 * nfunc = 1, sum_matrices = 0 is sxc_build_nadd. */
int sxc_build_nadd_multi(sxc_ctx* ctx, int grid, int nfunc, const int* funcs, int nspin, int basis_act, const double* P_act,
                         int nenv, const int* basis_env, const double* const* P_env, int env_frozen,
                         double block_ave_threshold, int sum_matrices, double* V_act, double* E /*[nfunc*(2+nenv)]*/);
int sxc_build_nadd_multi_device(sxc_ctx* ctx, int grid, int nfunc, const int* funcs, int nspin, int basis_act,
                                const double* d_P_act, int nenv, const int* basis_env, const double* const* d_P_env,
                                int env_frozen, double block_ave_threshold, int sum_matrices,
                                double* d_VE /* (sum ? 1 : nfunc)*nspin*nbA*nbA + nfunc*(2+nenv) */);

/* FuncPotential<SCFMode>::getGeomGradients (src/potentials/FuncPotential.cpp:114-239), SURVEY.md row f-3: the XC
 * contribution to the nuclear gradient.  atom_of_bf[nbf] = BasisController::getAtomIndicesOfBasis() (:127); grad is the
 * nAtoms x 3 column-major matrix the reference returns (element (A, c) at A + c * natoms), overwritten.  P as in
 * sxc_build_xc.  With a shard set, grad is this rank's partial sum. */
int sxc_xc_gradient(sxc_ctx* ctx, int grid, int basis, int func, int nspin, const double* P, int natoms,
                    const int* atom_of_bf, double* grad);
/* NAddFuncPotential<SCFMode>::getGeomGradients (potentials/NAddFuncPotential.cpp:329-493): the same contraction with the
 * non-additive potential v[rho_act + sum rho_env] - v[rho_act] and the ACTIVE density matrix; grad [natoms x 3]
 * column-major over the atoms of the active system (atom_of_bf maps its basis functions to them). */
int sxc_nadd_gradient(sxc_ctx* ctx, int grid, int func, int nspin, int basis_act, const double* P_act, int nenv,
                      const int* basis_env, const double* const* P_env, int natoms, const int* atom_of_bf, double* grad);

/* ---- stage-level entry points (the reference classes one level below the Potentials) -------------------- */
/* DensityOnGridCalculator::calcDensityAndGradientOnGrid (DensityOnGridCalculator.cpp:55-65): host outputs [N];
 * gx/gy/gz may be NULL.  With a shard set only the owned points are filled (others 0). */
int sxc_density_on_grid(sxc_ctx* ctx, int grid, int basis, const double* P, double* rho, double* gx, double* gy,
                        double* gz);
/* SupersystemDensityOnGridController::updateData (data/grid/SupersystemDensityOnGridController.cpp:95-193): sum of the
 * densities (and gradients) of ndens subsystems, each in its own basis, on the common grid; added in the order given. */
int sxc_supersystem_density_on_grid(sxc_ctx* ctx, int grid, int ndens, const int* basis, const double* const* P, double* rho,
                                    double* gx, double* gy, double* gz);
/* BasisFunctionOnGridController::getBlockOnGridData (BasisFunctionOnGridController.cpp:122-131), derivative
 * level 1: n x nbf column-major values (index mu*n + p) and the negligible flags; returns n through *n_out. */
int sxc_basis_on_grid(sxc_ctx* ctx, int grid, int basis, int block, double* val, double* dx, double* dy,
                      double* dz, int* negligible, int* n_out);
/* derivative level 2 of the same call (BasisFunctionOnGridController.cpp:302-304, :381-440, :1081-1095; the six arrays
 * BasisFunctionBlockOnGridData::secondDerivativeValues xx, xy, xz, yy, yz, zz): n x nbf column-major each. */
int sxc_basis_hessian_on_grid(sxc_ctx* ctx, int grid, int basis, int block, double* hxx, double* hxy, double* hxz, double* hyy,
                              double* hyz, double* hzz, int* n_out);
/* DensityOnGridCalculator::calcDensityAndDerivativesOnGrid, second derivatives of the density
 * (MatrixOperatorToGridTransformer.cpp:166-188): host outputs [N] each. */
int sxc_density_hessian_on_grid(sxc_ctx* ctx, int grid, int basis, const double* P, double* hxx, double* hxy, double* hxz,
                                double* hyy, double* hyz, double* hzz);
/* FunctionalLibrary::calcData(GRADIENTS) (FunctionalLibrary.cpp:39-72 -> XCFun.cpp:39-159), RESTRICTED, on host
 * arrays of length npts; gx..gz and dFdG* may be NULL for LDA functionals. */
int sxc_functional_on_grid(sxc_ctx* ctx, int func, int64_t npts, const double* w, const double* rho,
                           const double* gx, const double* gy, const double* gz, double* epuv, double* dFdRho,
                           double* dFdGx, double* dFdGy, double* dFdGz, double* energy);
/* same, UNRESTRICTED (XCFun.cpp:100-112, XC_A_B_AX_AY_AZ_BX_BY_BZ): dens8 / out8 are [8][npts] with rows
 * rho_a, grad_a x y z, rho_b, grad_b x y z  ->  dF/drho_a, dF/dgrad_a x y z, dF/drho_b, dF/dgrad_b x y z.
 * has_grad = 0: the gradient rows are ignored (LDA). */
int sxc_functional_on_grid_u(sxc_ctx* ctx, int func, int64_t npts, const double* w, const double* dens8, int has_grad,
                             double* epuv, double* out8, double* energy);
/* ScalarOperatorToMatrixAdder::addScalarOperatorToMatrix (ScalarOperatorToMatrixAdder.cpp:52-116); gx == NULL
 * selects the LDA variant; the result is ADDED to V as in the reference. */
int sxc_scalar_to_matrix(sxc_ctx* ctx, int grid, int basis, double block_ave_threshold, const double* v,
                         const double* gx, const double* gy, const double* gz, double* V);

/* ---- two-basis operators -------------------------------------------------------------------- */
/* ScalarOperatorToMatrixAdder(basisFunctionOnGridControllerA, basisFunctionOnGridControllerB, ...)::addScalarOperatorToMatrix
 * with A != B (ScalarOperatorToMatrixAdder.cpp:216-220 LDA, :286-300 GGA): V is nbf_A x nbf_B column-major and is ADDED
 * to; gx == NULL selects the LDA variant.  Both bases must live on the same grid handle. */
int sxc_scalar_to_matrix_ab(sxc_ctx* ctx, int grid, int basis_a, int basis_b, double block_ave_threshold, const double* v,
                            const double* gx, const double* gy, const double* gz, double* V);
/* ABFuncPotential<SCFMode>::getMatrix (potentials/ABFockMatrixConstruction/ABFuncPotential.cpp:54-160): the densities of
 * the ndens (basis_c[i], P_c[i]) pairs are summed on the grid, the functional is evaluated once and scattered into the
 * nbf_A x nbf_B matrix (per spin: nspin of them back to back).  E[0] = E_xc of the summed density, E[1] = its integral. */
int sxc_build_ab(sxc_ctx* ctx, int grid, int func, int nspin, int basis_a, int basis_b, int ndens, const int* basis_c,
                 const double* const* P_c, double block_ave_threshold, double* V_ab, double* E);

/* ABNAddFuncPotential<SCFMode>::getMatrix (potentials/ABFockMatrixConstruction/ABNAddFuncPotential.cpp:66-176): the
 * non-additive potential v[rho_act + sum_i rho_env_i] - v[rho_act] of the active system (basis_act, P_act) and the nenv
 * environment densities, scattered into the nbf_A x nbf_B matrix (per spin: nspin of them back to back), overwritten. */
int sxc_build_ab_nadd(sxc_ctx* ctx, int grid, int func, int nspin, int basis_a, int basis_b, int basis_act, const double* P_act,
                      int nenv, const int* basis_env, const double* const* P_env, double block_ave_threshold, double* V_ab);

/* ---- grid construction ---------------------------------------------------------------------- */
/* The partition-weight step of GridFactory::produce (src/grid/construction/GridFactory.cpp:139-266), the O(N n_atoms^2)
 * part of the reference's grid set-up.  coords [natoms][3] (bohr); xyz 3 x npts interleaved = every atom's reference
 * grid already shifted to its nucleus; parent[p] = index of the atom point p belongs to; w in: atomic quadrature weights
 * (AtomGridFactory), out: molecular weights (0 where the SSF screen removes the point; the caller applies the
 * weightThreshold cut of :264 and the Hilbert sort).  flavour 0 = BECKE (aij = size adjustments a(j + natoms * l) of
 * :95-113, may be NULL; becke_smoothing = GridFactory's _smoothing, max(1, k)-fold iterated polynomial :324-347),
 * 1 = SSF (aij, becke_smoothing ignored), 2 = VORONOI (:194-203). */
int sxc_partition_weights(sxc_ctx* ctx, int flavour, int becke_smoothing, int natoms, const double* coords,
                          const double* aij, int64_t npts, const double* xyz, const int* parent, double* w);

/* ---- grid construction behind the ABI (row f-1) ---------------------------------------------------------- */
/* One atom's reference grid (AtomGridFactory::produce, grid/construction/AtomGridFactory.cpp:78-255): radial_type 0 = AHLRICHS,
 * 1 = BECKE; accuracy 1..7; pruned Lebedev shells; points relative to the nucleus.  H..Kr. */
typedef struct sxc_grid_points sxc_grid_points;
int sxc_atom_grid(int nuclear_charge, int accuracy, int radial_type, sxc_grid_points** out);
/* GridFactory::produce (grid/construction/GridFactory.cpp:52-321): atom grids shifted to the nuclei, partition weights on the
 * device (flavour 0 = BECKE, 1 = SSF, 2 = VORONOI), weight cut (points with w <= weight_threshold dropped, reference 1e-14) and
 * the Hilbert R-tree sort of grid/HilbertRTreeSorting.cpp:29-214 (hilbert_sort != 0).  The result feeds sxc_set_grid. */
int sxc_molecular_grid(sxc_ctx* ctx, int natoms, const int* nuclear_charges, const double* coords_bohr, int accuracy, int flavour,
                       int radial_type, int becke_smoothing, double weight_threshold, int hilbert_sort, sxc_grid_points** out);
/* the permutation HilbertRTreeSorting::sort applies (descending Hilbert index, ties in input order); order[k] = input index */
int sxc_hilbert_rtree_order(int64_t npts, const double* xyz, int64_t* order);
int64_t sxc_grid_points_size(const sxc_grid_points* g);
const double* sxc_grid_points_xyz(const sxc_grid_points* g);     /* 3 x N interleaved (Eigen::Matrix3Xd layout) */
const double* sxc_grid_points_weights(const sxc_grid_points* g);
void sxc_grid_points_free(sxc_grid_points* g);
const char* sxc_grid_last_error(void);
/* device time (ms, CUDA events) of the kernel inside the last sxc_partition_weights call */
double sxc_last_partition_ms(sxc_ctx* ctx);

/* ---- LR-TDDFT / subsystem-TDDFT kernel -------------------------------------------------------- */
/* A kernel store is one set of second functional derivatives on a grid, i.e. one of the _pp/_pg/_gg members of
 * Kernel<SCFMode> (src/postHF/LRSCF/Kernel/Kernel.h:150-200): nspin = 1: [10][N] = d2F/drho2, d2F/drho d(grad rho) x y z,
 * d2F/d(grad rho)2 xx xy xz yy yz zz ([1][N] if gga = 0); nspin = 2: [33][N] = pp aa ab bb | pg {x,y,z} x {aa,ab,ba,bb} |
 * gg {xx,xy,xz,yy,yz,zz} x {aa,ab,bb} ([3][N] if gga = 0; the reference sets gg.ba := gg.ab, Kernel.cpp:580-600).
 * Created zeroed (Kernel.cpp:58-115). */
int sxc_kernel_create(sxc_ctx* ctx, int grid, int nspin, int gga, int* kernel);
int sxc_kernel_destroy(sxc_ctx* ctx, int kernel);
/* One storeDerivatives call of Kernel<SCFMode>::calculateDerivatives (Kernel.cpp:476-683, :686-747): the densities of the
 * ndens (basis_c[i], P_c[i]) pairs are summed on the grid (a subsystem: ndens = 1; the total density of :693-712: all of
 * them), FunctionalLibrary::calcData(GRADIENTS, func, order 2) is evaluated on it (XCFun.cpp:39-159, rows :288-298 /
 * :486-528), added to the store with factor `sign` (pm) and the store is zeroed where that density is below the
 * reference's hard-coded 1e-8.  nspin = 2: every P_c[i] is an {alpha, beta} pair stored back to back. */
int sxc_kernel_add(sxc_ctx* ctx, int kernel, int func, double sign, int ndens, const int* basis_c, const double* const* P_c);
/* copies the store to the host: [sxc_kernel_num_arrays][N] (Kernel::getPP / getPG / getGG without the block cut) */
int sxc_kernel_get(sxc_ctx* ctx, int kernel, double* out);
int sxc_kernel_num_arrays(sxc_ctx* ctx, int kernel);
/* KernelSigmavector<SCFMode>::contractKernel + contractBlock (src/postHF/LRSCF/Sigmavectors/KernelSigmavector.cpp:254-311,
 * :360-497) for nvec trial vectors: D = nvec x nspin matrices nbf_J x nbf_J (column-major, back to back; symmetrised here
 * as calcF does, :201-208), contracted to rho~ / grad rho~ on the grid and multiplied with the sum of the nkern (1 to 3)
 * kernel stores = Kernel::getPP/PG/GG(I, J) (total-density store + subsystem store for I == J + the _ppExact set of mixed
 * exact/approximate embedding, Kernel.cpp:170-230).
 * mode 0: RESTRICTED singlet; 1: RESTRICTED triplet, stores are UNRESTRICTED ones (aa - ab, :381-404); 2: UNRESTRICTED.
 * The result stays on the device (per grid); accumulate != 0 adds to the previous one (the supersystem contraction over
 * all subsystems J, :96-117 and :214-226). */
int sxc_kernel_contract(sxc_ctx* ctx, int grid, int basis_j, int nkern, const int* kernels, int mode, int nvec,
                        const double* D, int accumulate);
/* same with device-resident trial densities (left untouched); asynchronous on the context's stream */
int sxc_kernel_contract_device(sxc_ctx* ctx, int grid, int basis_j, int nkern, const int* kernels, int mode, int nvec,
                               const double* d_D, int accumulate);
/* save != 0: keep a copy of the grid's contracted response (KernelSigmavector::_supersystem_scalar / _supersystem_gradient,
 * KernelSigmavector.h:105-110); save == 0: make that copy the current response again, so that calcF(I, I) of every
 * subsystem I starts from the same supersystem contraction (:214-226). */
int sxc_kernel_response_copy(sxc_ctx* ctx, int grid, int save);
/* KernelSigmavector<SCFMode>::numericalIntegration (:313-358) + the F += F^T of calcF (:236-249): the contracted response of
 * the grid integrated with the basis functions of system I.  F = nvec x nspin matrices nbf_I x nbf_I, overwritten.  With a
 * shard set F is this rank's partial sum. */
int sxc_kernel_integrate(sxc_ctx* ctx, int grid, int basis_i, double* F);
/* same with the result left on the device (nvec x nspin x nbf_I^2 doubles: the buffer of the single all-reduce of a
 * multi-GPU sigma build); asynchronous on the context's stream */
int sxc_kernel_integrate_device(sxc_ctx* ctx, int grid, int basis_i, double* d_F);
/* calcF(I, I) for an isolated system: contract + integrate */
int sxc_kernel_sigma(sxc_ctx* ctx, int grid, int basis, int nkern, const int* kernels, int mode, int nvec, const double* D,
                     double* F);

/* ---- basis-set file front end (host only, no device needed) ---------------------------------- */
/* BasisFunctionProvider::provideAtomWithBasisFunctions (src/basis/BasisFunctionProvider.cpp:32-140) for every atom of a
 * geometry + the Shell constructor (src/basis/Shell.cpp:29-47: libint2-renormalised coefficients, Cartesian norm factors) +
 * the extended indices of BasisController (src/basis/BasisController.cpp:62-68): parses the Turbomole-format file `path`
 * (the reference's data/basis/<LABEL>) for the entry "<element> <basis_label>" of each atom; shells are atom-major in file
 * order.  elements[natoms] = element symbols (any case), coords_bohr [natoms][3].  spherical = settings basis.makeSphericalBasis.
 * $ecp sections are not read.  Errors: status < 0 and the reference's SerenityError wording in sxc_host_last_error(). */
int sxc_shell_table_from_file(const char* path, const char* basis_label, int natoms, const char* const* elements,
                              const double* coords_bohr, int spherical, sxc_shell_table** out);
int sxc_shell_table_sizes(const sxc_shell_table* t, int* nshell, int* nprim_total, int* nbf);
/* copies the arrays sxc_add_basis takes (any pointer may be NULL) and BasisController::getAtomIndicesOfBasis() [nbf] */
int sxc_shell_table_copy(const sxc_shell_table* t, int* l, int* pure, int* nprim, int* first_bf, double* centre, double* alpha,
                         double* coeff, double* normfac, int* atom_of_bf);
void sxc_shell_table_free(sxc_shell_table* t);
/* sxc_add_basis fed from the table */
int sxc_add_basis_from_table(sxc_ctx* ctx, const sxc_shell_table* t, double radial_threshold, int* basis);
const char* sxc_host_last_error(void);

int sxc_get_stats(sxc_ctx* ctx, sxc_stats* out);
/* host-only helper behind sxc_set_grid_shard: splits n blocks into `world` contiguous ranges of nearly equal summed
 * cost; bounds[world + 1] receives the range starts (bounds[0] = 0, bounds[world] = n). */
int sxc_balance_ranges(int n, const double* cost, int world, int* bounds);
int sxc_abi_version(void);
/* host-only introspection for the CPU test-suite: the warp-tile round schedule of the scatter kernel for a block with n32 =
 * s_pad / 32 row groups, 40 bytes per round (ngroups, 7 x pad, group[8], slot_a[8], slot_b[8], kmask[8]; slot 0xff = idle
 * warp); returns the number of rounds (rounds40 may be NULL) */
int sxc_debug_scatter_schedule(int n32, unsigned char* rounds40, int max_rounds);

/* same for the TMA scatter kernel's schedule (v2): 64 bytes per round (ngroups, 7 x pad, group[8], slot_a[8], slot_b[8], kmask[8],
 * group_a[8], group_b[8], 8 x pad); ks = k-steps per K chunk (2 or 4), kmask bit k = the warp multiplies k-step k */
int sxc_debug_scatter_schedule2(int n32, int ks, unsigned char* rounds64, int max_rounds);

#ifdef __cplusplus
}
#endif
#endif /* SERENITY_XC_B200_H */
