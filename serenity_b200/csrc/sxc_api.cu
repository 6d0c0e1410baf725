// sxc_api.cu - context, screening plans, chunked pipeline and the C ABI of libserenity_xc_b200.so.
//
// Orchestrates rows 8a-1 ... 8a-7 of SURVEY.md on one B200:
//   k_screen -> k_basis -> k_density -> k_functional -> k_form_g -> k_scatter (-> k_mirror, k_reduce_partials)
// over chunks of 128-point blocks whose phi / grad phi tiles fit the workspace (default 40 % of free HBM), phi being
// evaluated ONCE per build (the reference evaluates it twice, MatrixOperatorToGridTransformer.cpp:103 and
// ScalarOperatorToMatrixAdder.cpp:69-70).  Host side is plain C++17; no torch types cross this boundary.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>
#include <string>
#include <utility>
#include <vector>

#include "../../include/serenity_xc_b200.h"
#include "basis_kernels.cuh"
#include "comm.h"
#include "density_kernel.cuh"
#include "functionals.cuh"
#include "host_copy.h"
#include "scatter_kernel.cuh"
#include "scatter_tma.cuh"
#include "scatter_fused.cuh"
#include "gradient_kernels.cuh"
#include "grid_kernels.cuh"
#include "kernel2.cuh"
#include "sxc_common.cuh"

using namespace sxc;

namespace {

// ---------------------------------------------------------------------------------------------- device memory
struct DevMem {
  void* p = nullptr;
  size_t bytes = 0;
  DevMem() = default;
  DevMem(const DevMem&) = delete;
  DevMem& operator=(const DevMem&) = delete;
  ~DevMem() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

struct Grid {
  long npts = 0;
  int blocksize = 128;
  int nblocks = 0;
  int nlit = 0;  // literal 128-point blocks of the functional evaluation
  DevMem xyzw;   // x[N] y[N] z[N] w[N]
  int rank = 0, world = 1;
  bool owner_fixed = false;
  int own_first = 0, own_count = 0;  // owned contiguous block range
  DevMem dens;     // rho, gx, gy, gz            [4][N]
  DevMem pot;      // v_rho, v_gx, v_gy, v_gz    [4][N]
  DevMem tot;      // supersystem density        [4][N]   (NAdd)
  DevMem envsum;   // sum of environment densities [4][N] (NAdd, cached while frozen)
  DevMem parts;    // e_part, n_part, e_part2 [3][nlit]
  std::vector<int> env_key;       // basis handles (+ nspin) the cached envsum belongs to
  std::map<int, std::vector<double>> env_energy;  // cached E[rho_env_i] per functional handle (this rank's partial sums): the
                                                  // XC and the kinetic NAdd objects of one FDE iteration alternate
  bool env_valid = false;
  DevMem resp;     // row f-4: contracted kernel x response density, [nvec][4 * nspin][N]
  DevMem resp_saved;  // copy kept by sxc_kernel_response_copy (the supersystem contraction, reused for every subsystem I)
  int saved_nvec = 0, saved_nspin = 0, saved_gga = 0;
  int resp_nvec = 0, resp_nspin = 0, resp_gga = 0;
  GridView view() const {
    GridView v;
    v.npts = npts;
    v.blocksize = blocksize;
    v.nblocks = nblocks;
    v.x = xyzw.as<double>();
    v.y = v.x + npts;
    v.z = v.y + npts;
    v.w = v.z + npts;
    return v;
  }
};

struct Basis {
  int nshell = 0, nbf = 0;
  double radial_thr = 1e-9;
  DevMem ints;  // l, pure, nprim, prim_off, first_bf, nfunc  [6][nshell]
  DevMem dbl;   // centre[3 nshell], alpha, coeff, normfac
  size_t nprim_total = 0;
  int lmax = 0;
  ShellView view() const {
    ShellView v;
    v.nshell = nshell;
    v.nbf = nbf;
    const int* i = ints.as<int>();
    v.l = i;
    v.pure = i + nshell;
    v.nprim = i + 2 * nshell;
    v.prim_off = i + 3 * nshell;
    v.first_bf = i + 4 * nshell;
    v.nfunc = i + 5 * nshell;
    const double* d = dbl.as<double>();
    v.centre = d;
    v.alpha = d + 3 * (size_t)nshell;
    v.coeff = v.alpha + nprim_total;
    v.normfac = v.coeff + nprim_total;
    v.radial_thr = radial_thr;
    v.exp_thr = -std::log(radial_thr);
    return v;
  }
};

struct Chunk {
  int slot0 = 0, nslots = 0;
  size_t doubles = 0;     // tile buffer size of the chunk
  int order_off = 0;      // offset into Plan::order (sorted slots of this chunk)
  int ditem_off = 0, nditems = 0;  // k_density work items
  int vitem_off = 0, nvitems = 0;  // k_vmat work items
  bool vitems_whole = true;        // every k_vmat work item is a whole block (k_vmat_fg can form G for it)
  int fblock_off = 0, nfblocks = 0;  // k_vmat_fg: the chunk's blocks with s > 0 in queue order (what its formers walk through)
  bool dens_split = false;         // some block's j-tiles are spread over several CTAs (outputs accumulated)
};

struct Plan {
  long serial = 0;  // unique per context: identifies whose tiles sit in the workspace (sxc_set_tile_cache)
  int grid = -1, basis = -1;
  int nown = 0;
  int nbf_pad = 0;
  int s_pad_max = 0;
  int vmat_variant = 0;  // scatter kernel the round templates / work items were made for
  DevMem block_id, nsig_shell, s, sig_shell, sig_c0, sig_bf, s_pad, phi_off, order, tpl, tpl_off, skip, ditems, vitems;
  DevMem gflag;   // [nown] k_vmat_fg: 0 = G of the block not formed yet, 1 = formed, 2 = block-average test failed
  DevMem vorder;  // plan slots of the blocks with s > 0 in k_vmat's queue order, chunk by chunk (the formers' list of k_vmat_fg)
  std::vector<int> h_s, h_s_pad;
  std::vector<Chunk> chunks;
  sxc_stats stats{};
  PlanView view() const {
    PlanView v;
    v.nown = nown;
    v.block_id = block_id.as<int>();
    v.nsig_shell = nsig_shell.as<int>();
    v.s = s.as<int>();
    v.sig_shell = sig_shell.as<int>();
    v.sig_c0 = sig_c0.as<int>();
    v.sig_bf = sig_bf.as<int>();
    v.nbf_pad = nbf_pad;
    v.s_pad = s_pad.as<int>();
    v.phi_off = phi_off.as<long long>();
    return v;
  }
};

// row f-4: second functional derivatives on a grid (one Kernel::_pp/_pg/_gg set, Kernel.h:150-200)
struct KernelStore {
  int grid = -1;
  int nspin = 1;
  int gga = 1;
  int narr = 0;
  DevMem data;  // [narr][N]
};

}  // namespace

struct sxc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  std::string err;
  std::vector<std::unique_ptr<Grid>> grids;
  std::vector<std::unique_ptr<Basis>> bases;
  std::vector<FuncView> funcs;
  std::map<std::pair<int, int>, std::unique_ptr<Plan>> plans;
  std::vector<std::unique_ptr<KernelStore>> kstores;
  std::map<int, std::vector<ScatterRound>> scatter_tpl;  // round templates per s_pad / 32
  std::map<std::pair<int, int>, std::vector<ScatterRound2>> scatter_tpl2;  // v2 templates per (s_pad / 32, k-steps per chunk)
  // which scatter kernel runs: 0 = k_vmat (cp.async producers), 8 / 16 = k_vmat_tma<8 / 16> (TMA producer); SXC_VMAT overrides
  // 24 = k_vmat_fg (G formed by helper warpgroups inside the TKP = 8 kernel; chunks cut into segments fall back to 8).  Opt-in:
  // measured against 16 on two kinds of box of the pool it is 3 % faster per build on one and equal (peptide: 1 % slower) on
  // the other (profiles/r02_scatter_fused.md), and the two-launch form keeps the tensor kernel's time free of HBM work
  int vmat_variant = 16;
  int dens_variant = 0;  // 0 = k_density (cp.async producers; 1-3 % faster as measured), 1 = k_density_tma; SXC_DENS overrides
  int dens_prefetch = 0;  // SXC_DPF: bit 0 = L2 prefetch of k_density's epilogue rows (measured: no effect), bit 1 = development
  int func_variant = 4;   // SXC_FUNC: 4 / 5 = k_functional_occ<4 / 5> (restricted functional kernel at 4 / 5 CTAs per SM), 0 = k_functional
                          // (measured: peptide 0.416 / 0.373 / 0.363 ms, fde (H2O)64 0.412 / 0.372 / 0.369, tetracene 0.090 / 0.081 / 0.082)
  int basis_variant = 0;  // SXC_BASIS: 1 = k_basis<1> (128 registers, one CTA per SM), 2 = k_basis<2> (64 registers, two CTAs), 0 = by size
  int seg_waves = 6;  // SXC_SEG_WAVES: a shard with fewer blocks than 3 waves of resident CTAs is cut into items for this many waves
  int fg_lead = 0;   // SXC_FG_LEAD (development): k_vmat_fg's queue opens with medium-sized blocks (get_plan)
  // SXC_FG_SEG = 1: k_vmat_fg also on shards whose blocks are cut into segments.  Measured on rank 0 of an 8- / 4-rank tetracene
  // run: 0.918 / 1.704 ms per build against 0.873 / 1.665 ms with k_form_g + k_vmat_tma, so such shards take the two launches
  int fg_segments = 0;
  int fg_mode = 0;   // SXC_FG_MODE: development switches of k_vmat_fg
  int smem_pad = 0;  // SXC_SMEM_PAD: extra dynamic shared memory per DMMA CTA (development: forces one CTA per SM)
  int dseg = 1, vseg = 1;  // pieces per block of the k_density / k_vmat work items (SXC_DSEG / SXC_VSEG; 1 = only when a shard is small)
  CUtensorMap tmap_v8{}, tmap_v16{};  // tile workspace as [rows] x [128 points], boxes 32 x 8 (SWIZZLE_64B) and 32 x 16 (128B)
  CUtensorMap tmap_d16{};             // boxes of 16 rows x 16 points (SWIZZLE_128B) for k_density_tma
  CUtensorMap tmap_rows{};            // boxes of scat3::HROWS rows x 32 points (no swizzle) for the G formers of k_vmat_fg
  void* tmap_ptr = nullptr;
  size_t tmap_bytes = 0;
  DevMem phi;     // tile workspace (one chunk)
  DevMem phi2;    // second tile workspace: basis B of the two-basis scatter (row f-4)
  DevMem dP;      // staged density matrices (host API)
  DevMem dD;      // symmetrised copy of device-resident trial densities (sxc_kernel_contract_device; never touched by the
                  // side-stream uploads of the host entry points)
  DevMem dOut;    // staged V | E | N (host API)
  DevMem scratch; // small device scalars
  DevMem counters; // work-queue heads of the persistent kernels
  static constexpr int NCOUNTERS = 1024;
  int counter_next = NCOUNTERS;
  int num_sms = 148;
  int64_t ws_limit = 0;
  // sxc_set_tile_cache: the phi / grad phi tiles of a single-chunk plan stay valid in the workspace across builds (grid and
  // basis are fixed during an SCF; 180 GB of HBM make the reference's recomputation unnecessary)
  bool tile_cache = false;
  long phi_owner = 0;    // serial of the plan whose tiles ctx->phi holds (0: none)
  long plan_serial = 0;
  float last_partition_ms = 0.f;  // device time of the last k_partition_weights launch
  sxc_stats stats{};
  int launches = 0;
  long collectives = 0;  // ncclAllReduce calls issued by this context
  bool attrs_set = false;
  cudaEvent_t p_ready = nullptr;     // one-shot: the next build waits for it before it first reads P
  cudaStream_t copy_stream = nullptr; // H2D of P in the host-buffer entry points (overlaps screening + basis)
  cudaEvent_t copy_done = nullptr;
  int timing = 0;             // events recorded during the current build: 0 none, 1 first-to-last kernel only, 2 per kernel
  bool timing_device = false; // sxc_set_timing: per-kernel events in every build (also the *_device entry points)
  struct PendingUpload {
    void* dst;
    const void* src;
    size_t bytes;
  };
  std::vector<PendingUpload> pending;  // host-buffer entry points: P matrices to upload when the build first needs them
  // staged transfers of caller-owned pageable memory (host_copy.h): worker threads, page-locked staging buffers, chunk events
  std::unique_ptr<HostCopier> copier;
  void* h_up = nullptr;    // staging of the uploads (P)
  void* h_down = nullptr;  // staging of the downloads (V)
  size_t h_up_bytes = 0, h_down_bytes = 0;
  std::vector<cudaEvent_t> chunk_events;
  int copy_threads = 4;    // SXC_COPY_THREADS (0: leave pageable transfers to the driver); measured on a 16-core box with the 18.9 MB
                           // matrix of (H2O)64: 4 threads 6.08 ms end to end, 8: 6.66, 16: 6.97 - the DMA engine delivers what ~3 memcpy threads move
  int out_part = 0, out_parts = 1;  // sxc_set_output_slice: this context copies back part out_part of out_parts of a result matrix
  struct Stamp {
    int slot;
    cudaEvent_t a, b;
  };
  std::vector<Stamp> stamps;  // events of the last build, collected lazily
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  // multi-GPU: communicator of the ranks that share the grid (sxc_comm_init_rank); builds on a sharded grid end with one
  // all-reduce of their result buffer
  nccl_comm_t comm = nullptr;
  int comm_rank = 0, comm_world = 1;
};

namespace {

int fail(sxc_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define CU(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? SXC_ERR_NOMEM : SXC_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                        \
  } while (0)

#define LAUNCH_CHECK()                                                                                      \
  do {                                                                                                      \
    ++ctx->launches;                                                                                        \
    cudaError_t e_ = cudaGetLastError();                                                                    \
    if (e_ != cudaSuccess)                                                                                  \
      return fail(ctx, SXC_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define TRY(call)          \
  do {                     \
    int rc_ = (call);      \
    if (rc_ != SXC_OK) return rc_; \
  } while (0)

// handles are slots of a vector; released slots (sxc_release_*) are reused
template <class T>
int store_handle(std::vector<std::unique_ptr<T>>& v, std::unique_ptr<T> obj) {
  for (size_t i = 0; i < v.size(); ++i)
    if (!v[i]) {
      v[i] = std::move(obj);
      return (int)i;
    }
  v.push_back(std::move(obj));
  return (int)v.size() - 1;
}

Grid* get_grid(sxc_ctx* ctx, int h) { return (h >= 0 && h < (int)ctx->grids.size()) ? ctx->grids[h].get() : nullptr; }
Basis* get_basis(sxc_ctx* ctx, int h) { return (h >= 0 && h < (int)ctx->bases.size()) ? ctx->bases[h].get() : nullptr; }

int set_kernel_attrs(sxc_ctx* ctx) {
  if (ctx->attrs_set) return SXC_OK;
  CU(cudaFuncSetAttribute(k_density, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CU(cudaFuncSetAttribute(k_grad_contract, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CU(cudaFuncSetAttribute(k_vmat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scat::smem_bytes_pipe()));
  CU(cudaFuncSetAttribute(k_vmat_ab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scat::smem_bytes()));
  CU(cudaFuncSetAttribute(k_density_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CU(cudaFuncSetAttribute(k_vmat_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CU(cudaFuncSetAttribute(k_vmat_tma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CU(cudaFuncSetAttribute(k_vmat_fg, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CU(cudaFuncSetAttribute(k_vmat_tma<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  ctx->attrs_set = true;
  return SXC_OK;
}

constexpr size_t STAGE_MIN_BYTES = (size_t)1 << 20;  // smaller transfers are left to the driver
constexpr size_t P_SLACK_BYTES = 1024;  // room behind the density matrices in ctx->dP: an all-gather of ceil(n / world) doubles per rank
constexpr size_t STAGE_CHUNK = (size_t)4 << 20;

bool is_pageable(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

int ensure_staging(sxc_ctx* ctx, void** buf, size_t* have, size_t need) {
  if (!ctx->copier) ctx->copier = std::make_unique<HostCopier>(std::max(0, ctx->copy_threads - 1));
  if (*have >= need) return SXC_OK;
  if (*buf) cudaFreeHost(*buf);
  *buf = nullptr;
  *have = 0;
  CU(cudaHostAlloc(buf, need, cudaHostAllocDefault));
  *have = need;
  return SXC_OK;
}

// caller memory -> device on `stream`: pageable sources of >= 1 MB go through the page-locked staging buffer at `stage_off`,
// chunk by chunk (worker threads fill chunk i + 1 while the DMA engine moves chunk i)
int staged_h2d(sxc_ctx* ctx, void* dst, const void* src, size_t bytes, size_t stage_off, cudaStream_t stream) {
  if (ctx->copy_threads <= 0 || bytes < STAGE_MIN_BYTES || !is_pageable(src)) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return SXC_OK;
  }
  char* st = static_cast<char*>(ctx->h_up) + stage_off;
  for (size_t off = 0; off < bytes; off += STAGE_CHUNK) {
    const size_t n = std::min(STAGE_CHUNK, bytes - off);
    ctx->copier->copy(st + off, static_cast<const char*>(src) + off, n);
    CU(cudaMemcpyAsync(static_cast<char*>(dst) + off, st + off, n, cudaMemcpyHostToDevice, stream));
  }
  return SXC_OK;
}

// device -> caller memory, complete on return (the stream is synchronised up to the copy).  Pageable destinations of >= 1 MB:
// the DMA engine delivers D2H_CHUNK pieces into the page-locked staging buffer, each followed by an event; the copier job is
// opened when the first piece has arrived and the calling thread only raises the "delivered" mark event by event (copying slices
// itself while the next event is pending), so the memcpy into the caller's pages overlaps with the rest of the transfer.
constexpr size_t D2H_CHUNK = (size_t)512 << 10;
int staged_d2h(sxc_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
  if (ctx->copy_threads <= 0 || bytes < STAGE_MIN_BYTES || !is_pageable(dst)) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return SXC_OK;
  }
  TRY(ensure_staging(ctx, &ctx->h_down, &ctx->h_down_bytes, bytes));
  const size_t nchunk = (bytes + D2H_CHUNK - 1) / D2H_CHUNK;
  while (ctx->chunk_events.size() < nchunk) {
    cudaEvent_t e = nullptr;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->chunk_events.push_back(e);
  }
  char* st = static_cast<char*>(ctx->h_down);
  for (size_t c = 0; c < nchunk; ++c) {
    const size_t off = c * D2H_CHUNK, n = std::min(D2H_CHUNK, bytes - off);
    CU(cudaMemcpyAsync(st + off, static_cast<const char*>(src) + off, n, cudaMemcpyDeviceToHost, stream));
    CU(cudaEventRecord(ctx->chunk_events[c], stream));
  }
  cudaError_t err = cudaEventSynchronize(ctx->chunk_events[0]);
  if (err != cudaSuccess) return fail(ctx, SXC_ERR_CUDA, "staged_d2h: %s", cudaGetErrorString(err));
  ctx->copier->open(dst, st, bytes, std::min(D2H_CHUNK, bytes));
  for (size_t c = 1; c < nchunk && err == cudaSuccess; ++c) {
    while ((err = cudaEventQuery(ctx->chunk_events[c])) == cudaErrorNotReady) {
      cudaGetLastError();  // "not ready" is recorded like an error; it must not surface in a later check
      if (!ctx->copier->help()) {  // nothing to copy: block until the piece is there
        err = cudaEventSynchronize(ctx->chunk_events[c]);
        break;
      }
    }
    if (err == cudaSuccess) ctx->copier->publish(std::min((c + 1) * D2H_CHUNK, bytes));
  }
  if (err != cudaSuccess) cudaGetLastError();
  ctx->copier->publish(bytes);  // (on an error too: the job has to drain before the buffers can be reused)
  ctx->copier->close();
  if (err != cudaSuccess) return fail(ctx, SXC_ERR_CUDA, "staged_d2h: %s", cudaGetErrorString(err));
  return SXC_OK;
}

// result matrix -> the caller's buffer: all of it, or this context's part when several contexts of one process share the buffer
// (sxc_set_output_slice: every GPU of a group then moves 1 / N of the matrix over its own PCIe link)
int result_d2h(sxc_ctx* ctx, double* dst, const double* src, size_t n) {
  const size_t lo = n * (size_t)ctx->out_part / (size_t)ctx->out_parts, hi = n * ((size_t)ctx->out_part + 1) / (size_t)ctx->out_parts;
  if (hi > lo) TRY(staged_d2h(ctx, dst + lo, src + lo, (hi - lo) * sizeof(double), ctx->stream));
  return SXC_OK;
}

// P is first read by the density phase: an upload handed over by the caller (sxc_set_p_ready_event) or pending from a host-buffer
// entry point is started / awaited only there.  For caller-owned (pageable) host memory cudaMemcpyAsync blocks the host while the
// driver stages the data; issued here, behind the launches of the screening and basis kernels, that host time and the DMA on the
// side stream overlap with those kernels.
int wait_p_ready(sxc_ctx* ctx) {
  if (!ctx->pending.empty()) {
    if (!ctx->copy_stream) {
      CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
    }
    size_t total = 0;
    for (const auto& u : ctx->pending) total += u.bytes;
    if (ctx->copy_threads > 0 && total >= STAGE_MIN_BYTES) TRY(ensure_staging(ctx, &ctx->h_up, &ctx->h_up_bytes, total));
    size_t off = 0;
    // With a communicator every rank holds the SAME host matrix: each uploads 1 / world of it over its own PCIe link and the
    // slices are all-gathered over NVLink (the matrix crosses PCIe once per build instead of once per rank; the decision
    // depends on sizes only, so every rank takes it alike).  The gather writes ceil(n / world) doubles per rank: up to world - 1
    // doubles of padding behind the matrix, which the next matrix's gather overwrites or P_SLACK_BYTES absorbs.
    const int W = ctx->comm ? ctx->comm_world : 1;
    const char* p0 = static_cast<const char*>(ctx->dP.p);
    for (const auto& u : ctx->pending) {
      const char* d0 = static_cast<const char*>(u.dst);
      const bool in_dp = p0 && d0 >= p0 && d0 + u.bytes + P_SLACK_BYTES <= p0 + ctx->dP.bytes;
      if (W > 1 && (size_t)W * sizeof(double) <= P_SLACK_BYTES && u.bytes >= STAGE_MIN_BYTES && u.bytes % sizeof(double) == 0 && in_dp) {
        const size_t n = u.bytes / sizeof(double), cnt = (n + W - 1) / W;
        const size_t lo = std::min(n, (size_t)ctx->comm_rank * cnt), hi = std::min(n, lo + cnt);
        if (hi > lo)
          TRY(staged_h2d(ctx, static_cast<double*>(u.dst) + lo, static_cast<const double*>(u.src) + lo, (hi - lo) * sizeof(double), off,
                         ctx->copy_stream));
        const int rc = nccl().AllGather(static_cast<double*>(u.dst) + (size_t)ctx->comm_rank * cnt, u.dst, cnt, NCCL_DOUBLE, ctx->comm,
                                        ctx->copy_stream);
        if (rc != NCCL_SUCCESS) return fail(ctx, SXC_ERR_CUDA, "ncclAllGather failed: %s", nccl().GetErrorString(rc));
      } else {
        TRY(staged_h2d(ctx, u.dst, u.src, u.bytes, off, ctx->copy_stream));
      }
      off += u.bytes;
    }
    ctx->pending.clear();
    CU(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
    ctx->p_ready = ctx->copy_done;
  }
  if (ctx->p_ready) {
    cudaEvent_t ev = ctx->p_ready;
    ctx->p_ready = nullptr;
    CU(cudaStreamWaitEvent(ctx->stream, ev, 0));
  }
  return SXC_OK;
}

// host-buffer entry points: queue the upload (the staging buffer ctx->dP is free: the previous host call synchronised)
int upload_async(sxc_ctx* ctx, void* dst, const void* src, size_t bytes) {
  ctx->pending.push_back({dst, src, bytes});
  return SXC_OK;
}
int upload_done(sxc_ctx*) { return SXC_OK; }

// error path of the host-buffer entry points: no armed upload event, no copy in flight, no uncollected time stamps survive
int abort_build(sxc_ctx* ctx, int rc) {
  ctx->p_ready = nullptr;
  ctx->pending.clear();
  ctx->timing = 0;
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->stream);
  for (auto& st : ctx->stamps) {
    ctx->event_pool.push_back(st.a);
    ctx->event_pool.push_back(st.b);
  }
  ctx->stamps.clear();
  return rc;
}

// scope guard of the host-buffer entry points: an upload that the build never consumed (early error return) is withdrawn
struct HostCall {
  sxc_ctx* c;
  ~HostCall() {
    if (!c->pending.empty() || c->p_ready) abort_build(c, 0);
  }
};

cudaEvent_t take_event(sxc_ctx* ctx) {
  cudaEvent_t e = nullptr;
  if (!ctx->event_pool.empty()) {
    e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}

// RAII bracket of one kernel (or kernel pair) of the build: two events on the build's stream when timing is on
struct PhaseTimer {
  sxc_ctx* ctx;
  int slot;
  cudaEvent_t a = nullptr;
  PhaseTimer(sxc_ctx* c, int s) : ctx(c), slot(s) {
    if (ctx->timing >= (s == SXC_T_COUNT ? 1 : 2)) {
      a = take_event(ctx);
      cudaEventRecord(a, ctx->stream);
    }
  }
  ~PhaseTimer() {
    if (a) {
      cudaEvent_t b = take_event(ctx);
      cudaEventRecord(b, ctx->stream);
      ctx->stamps.push_back({slot, a, b});
    }
  }
};
constexpr int T_TOTAL = SXC_T_COUNT;  // pseudo slot of the whole build

// host-buffer entry points always know the device time of the whole build (two events); per-kernel events are opt-in
int timing_mode(const sxc_ctx* ctx, bool host_call) { return ctx->timing_device ? 2 : (host_call ? 1 : 0); }

void begin_timing(sxc_ctx* ctx, int on) {
  for (auto& st : ctx->stamps) {
    ctx->event_pool.push_back(st.a);
    ctx->event_pool.push_back(st.b);
  }
  ctx->stamps.clear();
  ctx->timing = on;
  for (int i = 0; i < SXC_T_COUNT; ++i) {
    ctx->stats.ms_kernel[i] = 0.f;
    ctx->stats.n_kernel[i] = 0;
  }
  ctx->stats.ms_total = 0.f;
}

void collect_timers(sxc_ctx* ctx) {
  for (auto& t : ctx->stamps) {
    float ms = 0.f;
    cudaEventSynchronize(t.b);
    cudaEventElapsedTime(&ms, t.a, t.b);
    if (t.slot == T_TOTAL) {
      ctx->stats.ms_total += ms;
    } else {
      ctx->stats.ms_kernel[t.slot] += ms;
      ctx->stats.n_kernel[t.slot] += 1;
    }
    ctx->event_pool.push_back(t.a);
    ctx->event_pool.push_back(t.b);
  }
  ctx->stamps.clear();
}

// The one collective of the path: the partial [V | E | N ...] of the ranks' block ranges are summed in
// place, on the build's stream, right behind its last kernel.
int allreduce_result(sxc_ctx* ctx, const Grid& g, double* buf, size_t count) {
  if (!ctx->comm || g.world == 1) return SXC_OK;
  if (g.world != ctx->comm_world || g.rank != ctx->comm_rank)
    return fail(ctx, SXC_ERR_INVALID, "grid shard %d/%d does not match the communicator's rank %d/%d", g.rank, g.world,
                ctx->comm_rank, ctx->comm_world);
  PhaseTimer t(ctx, SXC_T_ALLREDUCE);
  const int rc = nccl().AllReduce(buf, buf, count, NCCL_DOUBLE, NCCL_SUM, ctx->comm, ctx->stream);
  if (rc != NCCL_SUCCESS) return fail(ctx, SXC_ERR_CUDA, "ncclAllReduce failed: %s", nccl().GetErrorString(rc));
  ++ctx->collectives;
  return SXC_OK;
}

// ---------------------------------------------------------------------------------------------- plan
int run_screen(sxc_ctx* ctx, const Grid& g, const Basis& b, const Plan& p) {
  if (p.nown == 0) return SXC_OK;
  PhaseTimer t(ctx, SXC_T_SCREEN);
  k_screen<<<p.nown, 128, 0, ctx->stream>>>(g.view(), b.view(), p.view());
  LAUNCH_CHECK();
  return SXC_OK;
}

int alloc_plan_arrays(sxc_ctx* ctx, Plan& p, const Basis& b, const std::vector<int>& block_ids) {
  p.nown = (int)block_ids.size();
  p.nbf_pad = ((b.nbf + SPAD - 1) / SPAD) * SPAD;
  const size_t n = std::max<size_t>(p.nown, 1);
  CU(p.block_id.ensure(n * sizeof(int)));
  CU(p.nsig_shell.ensure(n * sizeof(int)));
  CU(p.s.ensure(n * sizeof(int)));
  CU(p.sig_shell.ensure(n * b.nshell * sizeof(int)));
  CU(p.sig_c0.ensure(n * b.nshell * sizeof(int)));
  CU(p.sig_bf.ensure(n * p.nbf_pad * sizeof(int)));
  CU(p.s_pad.ensure(n * sizeof(int)));
  CU(p.phi_off.ensure(n * sizeof(long long)));
  CU(p.skip.ensure(n * sizeof(int)));
  if (p.nown)
    CU(cudaMemcpyAsync(p.block_id.p, block_ids.data(), p.nown * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  return SXC_OK;
}

int64_t workspace_limit(sxc_ctx* ctx) {
  if (ctx->ws_limit > 0) return ctx->ws_limit;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return (int64_t)4 << 30;
  // keep what is already reserved for the workspace in the balance
  return (int64_t)((free_b + ctx->phi.bytes) * 0.4);
}

// Scatter schedule of a block with n32 = s_pad / 32 row groups: the n32 (n32 + 1) / 2 upper-triangle warp tiles are
// walked in bands of two row groups (column-major inside a band) and cut into rounds of <= WARPS tiles that touch
// <= MAXG distinct groups; rounds of <= WARPS / 2 tiles give every tile to two warps (one k-step of each chunk each).
std::vector<ScatterRound> build_scatter_schedule(int n32) {
  std::vector<std::pair<int, int>> tiles;
  for (int b0 = 0; b0 < n32; b0 += 2)
    for (int j = b0; j < n32; ++j)
      for (int i = b0; i < std::min(b0 + 2, n32); ++i)
        if (i <= j) tiles.push_back({i, j});
  const int T = (int)tiles.size();
  const int R = (T + scat::WARPS - 1) / scat::WARPS;
  const int per = (T + R - 1) / R;
  std::vector<ScatterRound> out;
  std::vector<std::pair<int, int>> cur;
  std::vector<int> groups;
  auto flush = [&]() {
    if (cur.empty()) return;
    ScatterRound r;
    std::memset(&r, 0, sizeof(r));
    std::memset(r.ta, 0xff, sizeof(r.ta));
    std::memset(r.tb, 0xff, sizeof(r.tb));
    std::sort(groups.begin(), groups.end());
    r.ngroups = (unsigned char)groups.size();
    for (size_t k = 0; k < groups.size(); ++k) r.group[k] = (unsigned char)groups[k];
    auto slot = [&](int gidx) { return (unsigned char)(std::find(groups.begin(), groups.end(), gidx) - groups.begin()); };
    const int nt = (int)cur.size();
    const bool split = nt * 2 <= scat::WARPS;
    for (int w = 0; w < nt; ++w) {
      r.ta[w] = slot(cur[w].first);
      r.tb[w] = slot(cur[w].second);
      r.kmask[w] = split ? 1 : 3;
      if (split) {
        r.ta[nt + w] = r.ta[w];
        r.tb[nt + w] = r.tb[w];
        r.kmask[nt + w] = 2;
      }
    }
    out.push_back(r);
    cur.clear();
    groups.clear();
  };
  for (const auto& t : tiles) {
    std::vector<int> g2 = groups;
    if (std::find(g2.begin(), g2.end(), t.first) == g2.end()) g2.push_back(t.first);
    if (std::find(g2.begin(), g2.end(), t.second) == g2.end()) g2.push_back(t.second);
    if ((int)cur.size() >= per || (int)g2.size() > scat::MAXG) {
      flush();
      g2.clear();
      g2.push_back(t.first);
      if (t.second != t.first) g2.push_back(t.second);
    }
    cur.push_back(t);
    groups = g2;
  }
  flush();
  return out;
}

const std::vector<ScatterRound>& scatter_schedule(sxc_ctx* ctx, int n32) {
  auto it = ctx->scatter_tpl.find(n32);
  if (it != ctx->scatter_tpl.end()) return it->second;
  return ctx->scatter_tpl[n32] = build_scatter_schedule(n32);
}

// Scatter schedule v2 (k_vmat_tma): the strictly upper triangle is covered by 2 x 4 rectangles of warp tiles (8 tiles on 6 staged
// groups) band by band; what the rectangles leave over - the band's inner off-diagonal tile and the diagonal tiles, which cost
// 10/16 of a full tile - first fills the idle warps of partial rounds whose staged groups already cover it, then forms rounds
// of its own (diagonal tiles together: such a round lasts 10/16 of a full one).  Finally every round hands its idle warps to
// the most expensive tiles: a tile split over nw warps costs ceil(KS / nw) of its KS k-steps per chunk.
std::vector<ScatterRound2> build_scatter_schedule2(int n32, int KS) {
  using Tile = std::pair<int, int>;
  const int W = scat2::WARPS, MAXG = scat2::MAXG;
  std::vector<std::vector<Tile>> rounds;
  std::vector<Tile> loose;
  for (int b0 = 0; b0 < n32; b0 += 2) {
    std::vector<int> rows;
    for (int r = b0; r < std::min(b0 + 2, n32); ++r) rows.push_back(r);
    if (b0 + 1 < n32) loose.push_back({b0, b0 + 1});
    for (int c0 = b0 + 2; c0 < n32; c0 += 4) {
      std::vector<Tile> tl;
      for (int c = c0; c < std::min(c0 + 4, n32); ++c)
        for (int r : rows) tl.push_back({r, c});
      rounds.push_back(tl);
    }
    for (int r : rows) loose.push_back({r, r});
  }
  auto ngroups_with = [](const std::vector<Tile>& tl, const Tile* extra) {
    std::vector<int> g;
    auto add = [&](int x) {
      if (std::find(g.begin(), g.end(), x) == g.end()) g.push_back(x);
    };
    for (const Tile& t : tl) {
      add(t.first);
      add(t.second);
    }
    if (extra) {
      add(extra->first);
      add(extra->second);
    }
    return (int)g.size();
  };
  // off-diagonal loose tiles first: they are the expensive ones
  std::stable_sort(loose.begin(), loose.end(), [](const Tile& a, const Tile& b) { return (a.first == a.second) < (b.first == b.second); });
  for (auto& tl : rounds) {
    while ((int)tl.size() < W && !loose.empty()) {
      const int g0 = ngroups_with(tl, nullptr);
      int best = -1, best_add = 1 << 30;
      for (size_t k = 0; k < loose.size(); ++k) {
        const int ng = ngroups_with(tl, &loose[k]);
        if (ng <= MAXG && ng - g0 < best_add) {
          best_add = ng - g0;
          best = (int)k;
        }
      }
      if (best < 0) break;
      tl.push_back(loose[best]);
      loose.erase(loose.begin() + best);
    }
  }
  // the rest: diagonal tiles together (cheap rounds), then the remaining off-diagonal ones
  std::stable_sort(loose.begin(), loose.end(), [](const Tile& a, const Tile& b) { return (a.first != a.second) < (b.first != b.second); });
  std::vector<Tile> cur;
  for (const Tile& t : loose) {
    if ((int)cur.size() >= W || ngroups_with(cur, &t) > MAXG) {
      rounds.push_back(cur);
      cur.clear();
    }
    cur.push_back(t);
  }
  if (!cur.empty()) rounds.push_back(cur);

  std::vector<ScatterRound2> out;
  for (const auto& tl : rounds) {
    const int nt = (int)tl.size();
    std::vector<int> nw(nt, 1);
    auto cost = [&](int k) { return (tl[k].first == tl[k].second ? 10.0 : 16.0) * ((KS + nw[k] - 1) / nw[k]); };
    for (int idle = W - nt; idle > 0; --idle) {
      int k = 0;
      for (int j = 1; j < nt; ++j)
        if (cost(j) > cost(k)) k = j;
      if (nw[k] >= KS) break;
      ++nw[k];
    }
    ScatterRound2 r;
    std::memset(&r, 0, sizeof(r));
    std::memset(r.ta, 0xff, sizeof(r.ta));
    std::memset(r.tb, 0xff, sizeof(r.tb));
    std::vector<int> groups;
    auto slot = [&](int g) {
      auto it = std::find(groups.begin(), groups.end(), g);
      if (it == groups.end()) {
        groups.push_back(g);
        return (int)groups.size() - 1;
      }
      return (int)(it - groups.begin());
    };
    int w = 0;
    for (int k = 0; k < nt; ++k)
      for (int part = 0; part < nw[k]; ++part, ++w) {
        r.ta[w] = (unsigned char)slot(tl[k].first);
        r.tb[w] = (unsigned char)slot(tl[k].second);
        r.ga[w] = (unsigned char)tl[k].first;
        r.gb[w] = (unsigned char)tl[k].second;
        unsigned m = 0;
        for (int ks = part; ks < KS; ks += nw[k]) m |= 1u << ks;
        r.kmask[w] = (unsigned char)m;
      }
    r.ngroups = (unsigned char)groups.size();
    for (size_t k = 0; k < groups.size() && k < sizeof(r.group); ++k) r.group[k] = (unsigned char)groups[k];
    out.push_back(r);
  }
  return out;
}

const std::vector<ScatterRound2>& scatter_schedule2(sxc_ctx* ctx, int n32, int KS) {
  auto key = std::make_pair(n32, KS);
  auto it = ctx->scatter_tpl2.find(key);
  if (it != ctx->scatter_tpl2.end()) return it->second;
  return ctx->scatter_tpl2[key] = build_scatter_schedule2(n32, KS);
}

// ---- TMA: 2-D tensor maps over the tile workspace ([rows] x [128 points] of FP64), one per box shape -----------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_tile_map(sxc_ctx* ctx, const DevMem& buf, int box_points, int box_rows, CUtensorMapSwizzle swz, CUtensorMap* out) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(ctx, SXC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)BP, (cuuint64_t)(buf.bytes / (BP * sizeof(double)))};
  const cuuint64_t gstride[1] = {(cuuint64_t)BP * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_points, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  static const int l2mode = []() {  // development switch SXC_TMA_L2: 0 none, 1 = 64 B, 2 = 128 B (default), 3 = 256 B promotion
    const char* v = std::getenv("SXC_TMA_L2");
    return v ? std::atoi(v) : 2;
  }();
  const CUtensorMapL2promotion promo = l2mode == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                       : l2mode == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : l2mode == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                     : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  const CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, buf.p, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(ctx, SXC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
  return SXC_OK;
}

// maps of the main workspace, re-encoded when the buffer moved or grew
int ensure_tile_maps(sxc_ctx* ctx) {
  if (ctx->tmap_ptr == ctx->phi.p && ctx->tmap_bytes == ctx->phi.bytes) return SXC_OK;
  TRY(make_tile_map(ctx, ctx->phi, 8, 32, CU_TENSOR_MAP_SWIZZLE_64B, &ctx->tmap_v8));
  TRY(make_tile_map(ctx, ctx->phi, 16, 32, CU_TENSOR_MAP_SWIZZLE_128B, &ctx->tmap_v16));
  TRY(make_tile_map(ctx, ctx->phi, 16, 16, CU_TENSOR_MAP_SWIZZLE_128B, &ctx->tmap_d16));
  TRY(make_tile_map(ctx, ctx->phi, scat3::HPTS, scat3::HROWS, CU_TENSOR_MAP_SWIZZLE_NONE, &ctx->tmap_rows));
  ctx->tmap_ptr = ctx->phi.p;
  ctx->tmap_bytes = ctx->phi.bytes;
  return SXC_OK;
}

constexpr int GRAD_PLAN = 1 << 20;  // key offset of the 8-slot plans of the gradient path

int get_plan(sxc_ctx* ctx, int gh, int bh, Plan** out, int comps = TILE_COMPS) {
  Grid* g = get_grid(ctx, gh);
  Basis* b = get_basis(ctx, bh);
  if (!g || !b) return fail(ctx, SXC_ERR_INVALID, "invalid grid (%d) or basis (%d) handle", gh, bh);
  auto key = std::make_pair(gh, bh + (comps == TILE_COMPS ? 0 : GRAD_PLAN));
  auto it = ctx->plans.find(key);
  if (it != ctx->plans.end()) {
    *out = it->second.get();
    return SXC_OK;
  }
  TRY(set_kernel_attrs(ctx));
  // 1. ownership: contiguous block ranges balanced on n_b s_b^2, fixed by the first basis
  if (!g->owner_fixed) {
    if (g->world == 1) {
      g->own_first = 0;
      g->own_count = g->nblocks;
    } else {
      Plan tmp;
      std::vector<int> all(g->nblocks);
      std::iota(all.begin(), all.end(), 0);
      TRY(alloc_plan_arrays(ctx, tmp, *b, all));
      TRY(run_screen(ctx, *g, *b, tmp));
      std::vector<int> s(g->nblocks);
      CU(cudaMemcpyAsync(s.data(), tmp.s.p, s.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      std::vector<double> cost(g->nblocks);
      for (int i = 0; i < g->nblocks; ++i) {
        const double sp = std::ceil(std::max(s[i], 1) / (double)SPAD) * SPAD;
        cost[i] = 128.0 * sp * sp + 3.0e4 * sp + 1.0e5;  // GEMMs + basis evaluation + fixed per-block work
      }
      std::vector<int> bound(g->world + 1);
      sxc_balance_ranges(g->nblocks, cost.data(), g->world, bound.data());
      g->own_first = bound[g->rank];
      g->own_count = bound[g->rank + 1] - bound[g->rank];
    }
    g->owner_fixed = true;
  }
  // 2. screening of the owned blocks
  auto plan = std::make_unique<Plan>();
  Plan& p = *plan;
  p.serial = ++ctx->plan_serial;
  p.grid = gh;
  p.basis = bh;
  std::vector<int> ids(g->own_count);
  std::iota(ids.begin(), ids.end(), g->own_first);
  TRY(alloc_plan_arrays(ctx, p, *b, ids));
  TRY(run_screen(ctx, *g, *b, p));
  p.h_s.resize(p.nown);
  if (p.nown) CU(cudaMemcpyAsync(p.h_s.data(), p.s.p, p.nown * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  // 3. padded sizes, chunking, work orders
  p.h_s_pad.resize(p.nown);
  std::vector<long long> off(p.nown);
  const int64_t limit = workspace_limit(ctx);
  sxc_stats& st = p.stats;
  st = sxc_stats{};
  st.nbf = b->nbf;
  st.nblocks = p.nown;
  Chunk cur;
  cur.slot0 = 0;
  for (int q = 0; q < p.nown; ++q) {
    const int s = p.h_s[q];
    const int sp = std::max(SPAD, ((s + SPAD - 1) / SPAD) * SPAD);
    p.h_s_pad[q] = sp;
    p.s_pad_max = std::max(p.s_pad_max, sp);
    const size_t tile = (size_t)comps * sp * BP;
    if (cur.nslots > 0 && (int64_t)((cur.doubles + tile) * sizeof(double)) > limit) {
      p.chunks.push_back(cur);
      cur = Chunk();
      cur.slot0 = q;
    }
    off[q] = (long long)cur.doubles;
    cur.doubles += tile;
    cur.nslots++;
    const long first = (long)ids[q] * g->blocksize;
    const long n = std::min<long>(g->blocksize, g->npts - first);
    st.npts += n;
    st.sum_s += s;
    st.sum_ns += n * s;
    st.sum_ns2 += n * (int64_t)s * s;
    st.sum_ns2_padded += (int64_t)BP * ((s + 7) & ~7) * ((s + 7) & ~7);  // the DMMA kernels run over s rounded up to 8
    st.sum_s2 += (int64_t)s * s;
    st.s_max = std::max<int64_t>(st.s_max, s);
  }
  if (cur.nslots > 0) p.chunks.push_back(cur);
  if (g->blocksize != FUNC_BLOCK && p.chunks.size() > 1)
    return fail(ctx, SXC_ERR_UNSUPPORTED, "grid.blocksize != 128 needs the whole grid in one workspace chunk");
  st.nchunks = (int)p.chunks.size();
  std::vector<int> order(std::max(p.nown, 1));
  for (Chunk& c : p.chunks) {  // work order of a chunk: largest blocks first
    c.order_off = c.slot0;
    std::iota(order.begin() + c.slot0, order.begin() + c.slot0 + c.nslots, c.slot0);
    std::stable_sort(order.begin() + c.slot0, order.begin() + c.slot0 + c.nslots,
                     [&](int a, int bq) { return p.h_s_pad[a] > p.h_s_pad[bq]; });
    st.workspace_bytes = std::max<int64_t>(st.workspace_bytes, (int64_t)(c.doubles * sizeof(double)));
  }
  // scatter round templates for every s_pad / 32 up to the largest block
  const int n32max = p.s_pad_max / 32;
  std::vector<int> tpl_off(n32max + 2, 0);
  std::vector<ScatterRound> tpl;     // k_vmat
  std::vector<ScatterRound2> tpl2;   // k_vmat_tma<vmat_variant>
  p.vmat_variant = ctx->vmat_variant;
  for (int n32 = 1; n32 <= n32max; ++n32) {
    if (p.vmat_variant == 0) {
      tpl_off[n32] = (int)tpl.size();
      const auto& rs = scatter_schedule(ctx, n32);
      tpl.insert(tpl.end(), rs.begin(), rs.end());
    } else {
      tpl_off[n32] = (int)tpl2.size();
      const auto& rs = scatter_schedule2(ctx, n32, p.vmat_variant == 16 ? 4 : 2);
      tpl2.insert(tpl2.end(), rs.begin(), rs.end());
    }
  }
  tpl_off[n32max + 1] = (int)(p.vmat_variant == 0 ? tpl.size() : tpl2.size());
  tpl_off[0] = 0;
  CU(p.order.ensure(order.size() * sizeof(int)));
  // work items of the DMMA kernels: whole blocks, unless that leaves fewer than ~3 waves of the resident CTAs (strong
  // scaling of a small grid over many GPUs): then blocks are cut into segments of j-tiles / rounds
  std::vector<WorkItem> ditems, vitems;
  std::vector<int> vq;  // plan slots of the blocks behind vitems, each once
  // (threshold: fewer blocks than 3 waves of the resident CTAs; granularity: items for ~seg_waves waves - measured on rank 0 of an
  // 8-rank tetracene run: 3 waves 0.892 ms per build, 4: 0.882, 6: 0.875; whole blocks: 1.07)
  const int target = 3 * 2 * ctx->num_sms;
  const int seg_target = ctx->seg_waves * 2 * ctx->num_sms;
  for (Chunk& c : p.chunks) {
    long tot_jt = 0, tot_r = 0;
    for (int k = 0; k < c.nslots; ++k) {
      const int q = order[c.slot0 + k];
      if (p.h_s[q] == 0) continue;
      const int n32 = p.h_s_pad[q] / 32;
      tot_jt += (n32 + dens::NJW - 1) / dens::NJW;
      tot_r += tpl_off[n32 + 1] - tpl_off[n32];
    }
    const int seg_jt = (int)std::max<long>(1, tot_jt / seg_target), seg_r = (int)std::max<long>(1, tot_r / seg_target);
    c.ditem_off = (int)ditems.size();
    c.vitem_off = (int)vitems.size();
    c.fblock_off = (int)vq.size();
    // k_vmat_fg forms the G of a CTA's next item while its DMMA warps contract the current one; only the FIRST item of every CTA
    // is formed with the tensor pipe idle.  Largest-first would make those first items the most expensive ones to form (all
    // formers start together and share HBM).  Instead the queue opens with the smallest blocks whose contraction still covers the
    // forming of the largest block that follows (contraction ~ s^2, forming ~ s: s_first^2 >= ~100 s_max from the measured
    // rates), as many as there are resident CTAs; everything else stays largest-first.
    std::vector<int> vorder(order.begin() + c.slot0, order.begin() + c.slot0 + c.nslots);
    if (p.vmat_variant == 24 && ctx->fg_lead && c.nslots > 0) {
      const int G = 2 * ctx->num_sms;
      const double smax = p.h_s_pad[vorder[0]];
      const double s_first = std::sqrt(100.0 * smax);
      int hi = 0;  // blocks [0, hi) have s_pad >= s_first (descending order)
      while (hi < c.nslots && p.h_s_pad[vorder[hi]] >= s_first) ++hi;
      if (hi >= 2 * G) std::rotate(vorder.begin(), vorder.begin() + (hi - G), vorder.begin() + hi);
    }
    for (int k = 0; k < c.nslots; ++k) {
      const int q = order[c.slot0 + k];
      const int n32 = p.h_s_pad[q] / 32;
      const int njt = p.h_s[q] == 0 ? 0 : (n32 + dens::NJW - 1) / dens::NJW;
      // (dseg / vseg > 1: every block is cut into that many pieces even on a full GPU - CTAs that run at the same time then work on
      // the same few blocks, whose tiles stay in L2 between the re-reads of the j-tiles / rounds)
      const int want_d = ctx->dseg > 1 ? std::min(njt, ctx->dseg) : 1;
      if (want_d <= 1 && (njt <= seg_jt || c.nslots >= target)) {
        ditems.push_back(WorkItem{q, 0, (short)njt, 0, 1});
      } else {
        c.dens_split = true;
        const int nseg = std::max(want_d, (c.nslots >= target) ? 1 : (njt + seg_jt - 1) / seg_jt);
        for (int sgi = 0; sgi < nseg; ++sgi)  // equal-sized segments
          ditems.push_back(WorkItem{q, (short)((long)njt * sgi / nseg), (short)((long)njt * (sgi + 1) / nseg), (short)sgi, (short)nseg});
      }
    }
    for (int k = 0; k < c.nslots; ++k) {
      const int q = vorder[k];
      if (p.h_s[q] == 0) continue;
      const int n32 = p.h_s_pad[q] / 32;
      const int nr = tpl_off[n32 + 1] - tpl_off[n32];
      const int want_v = ctx->vseg > 1 ? std::min(nr, ctx->vseg) : 1;
      vq.push_back(q);
      if (want_v <= 1 && (nr <= seg_r || c.nslots >= target)) {
        vitems.push_back(WorkItem{q, 0, (short)nr, 0, 1});
      } else {
        c.vitems_whole = false;
        const int nseg = std::max(want_v, (c.nslots >= target) ? 1 : (nr + seg_r - 1) / seg_r);
        for (int sgi = 0; sgi < nseg; ++sgi)
          vitems.push_back(WorkItem{q, (short)((long)nr * sgi / nseg), (short)((long)nr * (sgi + 1) / nseg), (short)sgi, (short)nseg});
      }
    }
    c.nfblocks = (int)vq.size() - c.fblock_off;
    c.nditems = (int)ditems.size() - c.ditem_off;
    c.nvitems = (int)vitems.size() - c.vitem_off;
  }
  CU(p.gflag.ensure(std::max<size_t>(p.nown, 1) * sizeof(int)));
  CU(p.vorder.ensure(std::max<size_t>(vq.size(), 1) * sizeof(int)));
  if (!vq.empty()) CU(cudaMemcpyAsync(p.vorder.p, vq.data(), vq.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CU(p.ditems.ensure(std::max<size_t>(ditems.size(), 1) * sizeof(WorkItem)));
  CU(p.vitems.ensure(std::max<size_t>(vitems.size(), 1) * sizeof(WorkItem)));
  if (!ditems.empty())
    CU(cudaMemcpyAsync(p.ditems.p, ditems.data(), ditems.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, ctx->stream));
  if (!vitems.empty())
    CU(cudaMemcpyAsync(p.vitems.p, vitems.data(), vitems.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, ctx->stream));
  CU(p.tpl.ensure(std::max<size_t>(tpl.size(), 1) * sizeof(ScatterRound) + std::max<size_t>(tpl2.size(), 1) * sizeof(ScatterRound2)));
  CU(p.tpl_off.ensure(tpl_off.size() * sizeof(int)));
  if (p.nown) {
    CU(cudaMemcpyAsync(p.s_pad.p, p.h_s_pad.data(), p.nown * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(p.phi_off.p, off.data(), p.nown * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(p.order.p, order.data(), p.nown * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (!tpl.empty())
    CU(cudaMemcpyAsync(p.tpl.p, tpl.data(), tpl.size() * sizeof(ScatterRound), cudaMemcpyHostToDevice, ctx->stream));
  if (!tpl2.empty())
    CU(cudaMemcpyAsync(p.tpl.p, tpl2.data(), tpl2.size() * sizeof(ScatterRound2), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(p.tpl_off.p, tpl_off.data(), tpl_off.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (std::max(dens::smem_bytes_pipe(p.s_pad_max), dens::smem_bytes_tma(p.s_pad_max)) > 227 * 1024 || p.s_pad_max / 32 > 255)
    return fail(ctx, SXC_ERR_UNSUPPORTED, "more than %d significant functions in one block", p.s_pad_max);
  *out = plan.get();
  ctx->plans[key] = std::move(plan);
  return SXC_OK;
}

// per-point SoA arrays: rows rho, gx, gy, gz per spin ([4 * nspin][N])
int ensure_point_arrays(sxc_ctx* ctx, Grid& g, bool nadd, int nspin, int nparts = 3, int npot = 1) {
  const size_t n4 = (size_t)4 * nspin * std::max<long>(g.npts, 1) * sizeof(double);
  if (g.dens.bytes < n4) {
    CU(g.dens.ensure(n4));
    CU(cudaMemsetAsync(g.dens.p, 0, n4, ctx->stream));
  }
  if (g.pot.bytes < n4 * npot) {  // (npot potentials side by side: the summed multi-functional scatter)
    CU(g.pot.ensure(n4 * npot));
    CU(cudaMemsetAsync(g.pot.p, 0, n4 * npot, ctx->stream));
  }
  CU(g.parts.ensure((size_t)std::max(3, nparts) * std::max(g.nlit, 1) * sizeof(double)));
  if (nadd) {
    if (g.tot.bytes < n4) {
      CU(g.tot.ensure(n4));
      CU(cudaMemsetAsync(g.tot.p, 0, n4, ctx->stream));
    }
    if (g.envsum.bytes < n4) {
      CU(g.envsum.ensure(n4));
      g.env_valid = false;
    }
  }
  return SXC_OK;
}

// phases of one chunk ------------------------------------------------------------------------------------------
bool tiles_cached(const sxc_ctx* ctx, const Plan& p) {
  return ctx->tile_cache && p.chunks.size() == 1 && p.serial != 0 && ctx->phi_owner == p.serial;
}

int phase_basis(sxc_ctx* ctx, const Grid& g, const Basis& b, const Plan& p, const Chunk& c, DevMem* buf = nullptr) {
  DevMem& phi = buf ? *buf : ctx->phi;
  if (!buf) ctx->phi_owner = p.chunks.size() == 1 ? p.serial : 0;
  CU(phi.ensure(c.doubles * sizeof(double)));
  PhaseTimer t(ctx, SXC_T_BASIS);
  // 64 registers / two CTAs per SM pay off when shells are small (def2-SVP: (H2O)64 0.90 -> 0.79 ms, peptide 2.21 -> 1.87 ms;
  // def2-TZVP tetracene unchanged) and the chunk has enough blocks for two waves of them; a small shard is faster with one
  const int variant = ctx->basis_variant ? ctx->basis_variant : (c.nslots >= 4 * ctx->num_sms ? 2 : 1);
  if (variant == 2)
    k_basis<2><<<c.nslots, BASIS_GROUPS * BP, 0, ctx->stream>>>(g.view(), b.view(), p.view(), c.slot0,
                                                               p.order.as<int>() + c.order_off, phi.as<double>());
  else
    k_basis<1><<<c.nslots, BASIS_GROUPS * BP, 0, ctx->stream>>>(g.view(), b.view(), p.view(), c.slot0,
                                                               p.order.as<int>() + c.order_off, phi.as<double>());
  LAUNCH_CHECK();
  return SXC_OK;
}

__global__ void k_zero_blocks(long N, int blocksize, int ncomp, const int* __restrict__ block_id, double* __restrict__ out) {
  const long first = (long)block_id[blockIdx.x] * blocksize;
  const long n = min((long)blocksize, N - first);
  for (int c = 0; c < ncomp; ++c)
    for (long i = threadIdx.x; i < n; i += blockDim.x) out[(size_t)c * N + first + i] = 0.0;
}

int phase_density(sxc_ctx* ctx, const Grid& g, const Basis& b, const Plan& p, const Chunk& c, const double* dP,
                  double* dens4, bool with_grad, int* nonneg, int a_slot = 0, int e_slot0 = 0) {
  const long N = g.npts;
  if (c.nditems == 0) return SXC_OK;
  PhaseTimer t(ctx, SXC_T_DENSITY);
  if (c.dens_split) {  // blocks shared by several CTAs accumulate into their outputs: clear the chunk's points first
    k_zero_blocks<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, with_grad ? 4 : 1, p.block_id.as<int>() + c.slot0, dens4);
    LAUNCH_CHECK();
  }
  if (ctx->dens_variant == 0 || a_slot != 0 || e_slot0 != 0) {
    k_density<<<c.nditems, dens::PTHREADS, dens::smem_bytes_pipe(p.s_pad_max), ctx->stream>>>(
        g.view(), p.view(), b.nbf, dP, p.ditems.as<WorkItem>() + c.ditem_off, ctx->phi.as<double>(), dens4,
        with_grad ? dens4 + N : nullptr, with_grad ? dens4 + 2 * N : nullptr, with_grad ? dens4 + 3 * N : nullptr, nonneg,
        a_slot, e_slot0, ctx->dens_prefetch);
  } else {
    TRY(ensure_tile_maps(ctx));
    k_density_tma<<<c.nditems, dens::PTHREADS, dens::smem_bytes_tma(p.s_pad_max) + ctx->smem_pad, ctx->stream>>>(
        ctx->tmap_d16, g.view(), p.view(), b.nbf, dP, p.ditems.as<WorkItem>() + c.ditem_off, dens4,
        with_grad ? dens4 + N : nullptr, with_grad ? dens4 + 2 * N : nullptr, with_grad ? dens4 + 3 * N : nullptr, nonneg);
  }
  LAUNCH_CHECK();
  return SXC_OK;
}

// functional on the literal blocks covered by the chunk (or on all of them when blocksize != 128);
// dens / pot are [4 * nspin][N]
int phase_functional(sxc_ctx* ctx, const Grid& g, const Plan& p, const Chunk& c, const FuncView& f, int nspin,
                     const double* dens, double sign, int accumulate, double* pot, double* e_part, double* n_part) {
  const long N = g.npts;
  const GridView gv = g.view();
  const bool lit_is_block = g.blocksize == FUNC_BLOCK;
  const int nb = lit_is_block ? c.nslots : g.nlit;
  if (nb == 0) return SXC_OK;
  const int* list = lit_is_block ? p.block_id.as<int>() + c.slot0 : nullptr;
  PhaseTimer t(ctx, SXC_T_FUNCTIONAL);
  if (nspin == 2) {
    k_functional_u<<<nb, FUNC_BLOCK, 0, ctx->stream>>>(f, N, list, gv.w, dens, sign, accumulate, nullptr, pot, e_part,
                                                       n_part);
  } else {
    auto kern = ctx->func_variant == 4 ? k_functional_occ<4> : ctx->func_variant == 5 ? k_functional_occ<5> : k_functional;
    kern<<<nb, FUNC_BLOCK, 0, ctx->stream>>>(f, N, list, gv.w, dens, dens + N, dens + 2 * N, dens + 3 * N, sign, accumulate, nullptr,
                                             pot, f.gga ? pot + N : nullptr, f.gga ? pot + 2 * N : nullptr,
                                             f.gga ? pot + 3 * N : nullptr, e_part, n_part);
  }
  LAUNCH_CHECK();
  return SXC_OK;
}

int next_counter(sxc_ctx* ctx, int** out) {
  if (ctx->counter_next >= sxc_ctx::NCOUNTERS) {  // recycle the slots (one per persistent launch)
    CU(cudaMemsetAsync(ctx->counters.p, 0, sxc_ctx::NCOUNTERS * sizeof(int), ctx->stream));
    ctx->counter_next = 0;
  }
  *out = ctx->counters.as<int>() + ctx->counter_next++;
  return SXC_OK;
}

// k_form_g is bandwidth-bound: aim at >= 8 resident CTAs per SM by letting up to 8 CTAs share a block's function rows
dim3 form_g_grid(const sxc_ctx* ctx, int nslots) {
  const int want = 8 * ctx->num_sms;
  const int split = std::max(1, std::min(8, (want + std::max(nslots, 1) - 1) / std::max(nslots, 1)));
  return dim3((unsigned)nslots, (unsigned)split);
}

int phase_scatter(sxc_ctx* ctx, const Grid& g, const Basis& b, const Plan& p, const Chunk& c, bool gga,
                  double block_ave_thr, const double* pot4, double* dW, int npot = 1, size_t pot_stride = 0) {
  const long N = g.npts;
  if (c.nslots == 0) return SXC_OK;
  const double* v_gx = gga ? pot4 + N : nullptr;
  const double* v_gy = gga ? pot4 + 2 * N : nullptr;
  const double* v_gz = gga ? pot4 + 3 * N : nullptr;
  const int grid = std::min(c.nvitems, 2 * ctx->num_sms);
  if (p.vmat_variant == 24 && (c.vitems_whole || ctx->fg_segments)) {
    // one launch: the helper warpgroups of all CTAs form G block by block in queue order while the DMMA warps contract the work
    // items (blocks, or segments of their rounds on a small shard) whose G is ready (scatter_fused.cuh)
    if (c.nvitems == 0) return SXC_OK;
    int *counter = nullptr, *fcounter = nullptr;
    TRY(next_counter(ctx, &counter));
    TRY(next_counter(ctx, &fcounter));
    TRY(ensure_tile_maps(ctx));
    PhaseTimer t(ctx, SXC_T_SCATTER);
    CU(cudaMemsetAsync(p.gflag.as<int>() + c.slot0, 0, (size_t)c.nslots * sizeof(int), ctx->stream));
    k_vmat_fg<<<grid, scat3::THREADS, scat3::smem_bytes(p.s_pad_max), ctx->stream>>>(
        ctx->tmap_v8, ctx->tmap_rows, g.view(), p.view(), b.nbf, p.vitems.as<WorkItem>() + c.vitem_off, c.nvitems,
        p.vorder.as<int>() + c.fblock_off, c.nfblocks, counter, fcounter, p.tpl.as<ScatterRound2>(), p.tpl_off.as<int>(), p.s_pad_max,
        block_ave_thr, 0.5, pot4, v_gx, v_gy, v_gz, npot, pot_stride, ctx->phi.as<double>(), dW, p.gflag.as<int>(), ctx->fg_mode);
    LAUNCH_CHECK();
    return SXC_OK;
  }
  {
    PhaseTimer t(ctx, SXC_T_FORM_G);
    k_form_g<<<form_g_grid(ctx, c.nslots), 256, 0, ctx->stream>>>(g.view(), p.view(), p.order.as<int>() + c.order_off, block_ave_thr, 0.5,
                                                pot4, v_gx, v_gy, v_gz, npot, pot_stride, ctx->phi.as<double>(), p.skip.as<int>());
    LAUNCH_CHECK();
  }
  int* counter = nullptr;
  TRY(next_counter(ctx, &counter));
  PhaseTimer t(ctx, SXC_T_SCATTER);
  if (c.nvitems == 0) return SXC_OK;
  if (p.vmat_variant == 0) {
    k_vmat<<<grid, scat::PTHREADS, scat::smem_bytes_pipe(), ctx->stream>>>(
        p.view(), b.nbf, p.vitems.as<WorkItem>() + c.vitem_off, c.nvitems, counter, p.skip.as<int>(), p.tpl.as<ScatterRound>(),
        p.tpl_off.as<int>(), ctx->phi.as<double>(), dW);
  } else {
    TRY(ensure_tile_maps(ctx));
    if (p.vmat_variant == 16)
      k_vmat_tma<16><<<grid, scat2::THREADS, scat2::smem_bytes<16>(p.s_pad_max) + ctx->smem_pad, ctx->stream>>>(
          ctx->tmap_v16, p.view(), b.nbf, p.vitems.as<WorkItem>() + c.vitem_off, c.nvitems, counter, p.skip.as<int>(),
          p.tpl.as<ScatterRound2>(), p.tpl_off.as<int>(), p.s_pad_max, dW);
    else if (p.vmat_variant == 83)
      k_vmat_tma<8, 3><<<grid, scat2::THREADS, scat2::smem_bytes<8>(p.s_pad_max) + ctx->smem_pad, ctx->stream>>>(
          ctx->tmap_v8, p.view(), b.nbf, p.vitems.as<WorkItem>() + c.vitem_off, c.nvitems, counter, p.skip.as<int>(),
          p.tpl.as<ScatterRound2>(), p.tpl_off.as<int>(), p.s_pad_max, dW);
    else
      k_vmat_tma<8><<<grid, scat2::THREADS, scat2::smem_bytes<8>(p.s_pad_max) + ctx->smem_pad, ctx->stream>>>(
          ctx->tmap_v8, p.view(), b.nbf, p.vitems.as<WorkItem>() + c.vitem_off, c.nvitems, counter, p.skip.as<int>(),
          p.tpl.as<ScatterRound2>(), p.tpl_off.as<int>(), p.s_pad_max, dW);
  }
  LAUNCH_CHECK();
  return SXC_OK;
}

// two-basis scatter of one spin: tiles of A in ctx->phi, of B in ctx->phi2 (both plans single-chunk, same blocks / slots)
int phase_scatter_ab(sxc_ctx* ctx, const Grid& g, const Basis& bA, const Plan& pA, const Plan& pB, bool gga,
                     double block_ave_thr, const double* pot4, double* dW) {
  const long N = g.npts;
  const Chunk& cA = pA.chunks[0];
  const Chunk& cB = pB.chunks[0];
  if (cA.nslots == 0) return SXC_OK;
  {
    PhaseTimer t(ctx, SXC_T_FORM_G);  // G_A = grad_A (no scalar part), G_B = a phi_B + grad_B
    k_form_g<<<form_g_grid(ctx, cA.nslots), 256, 0, ctx->stream>>>(g.view(), pA.view(), pA.order.as<int>() + cA.order_off, block_ave_thr, 0.0,
                                                 pot4, gga ? pot4 + N : nullptr, gga ? pot4 + 2 * N : nullptr,
                                                 gga ? pot4 + 3 * N : nullptr, 1, 0, ctx->phi.as<double>(), pA.skip.as<int>());
    LAUNCH_CHECK();
    k_form_g<<<form_g_grid(ctx, cB.nslots), 256, 0, ctx->stream>>>(g.view(), pB.view(), pB.order.as<int>() + cB.order_off, block_ave_thr, 1.0,
                                                 pot4, gga ? pot4 + N : nullptr, gga ? pot4 + 2 * N : nullptr,
                                                 gga ? pot4 + 3 * N : nullptr, 1, 0, ctx->phi2.as<double>(), pB.skip.as<int>());
    LAUNCH_CHECK();
  }
  int* counter = nullptr;
  TRY(next_counter(ctx, &counter));
  PhaseTimer t(ctx, SXC_T_SCATTER);
  const int grid = std::min(cA.nslots, 2 * ctx->num_sms);
  k_vmat_ab<<<grid, scat::THREADS, scat::smem_bytes(), ctx->stream>>>(
      pA.view(), pB.view(), bA.nbf, pA.order.as<int>() + cA.order_off, cA.nslots, counter, pA.skip.as<int>(),
      pB.skip.as<int>(), ctx->phi.as<double>(), ctx->phi2.as<double>(), dW);
  LAUNCH_CHECK();
  return SXC_OK;
}

int ab_plans(sxc_ctx* ctx, int gh, int bA, int bB, Plan** pa, Plan** pb) {
  if (!get_grid(ctx, gh) || !get_basis(ctx, bA) || !get_basis(ctx, bB))
    return fail(ctx, SXC_ERR_INVALID, "invalid grid (%d) or basis (%d, %d) handle", gh, bA, bB);
  TRY(get_plan(ctx, gh, bA, pa));
  TRY(get_plan(ctx, gh, bB, pb));
  if ((*pa)->chunks.size() > 1 || (*pb)->chunks.size() > 1)
    return fail(ctx, SXC_ERR_UNSUPPORTED, "the two-basis scatter needs the tiles of both bases resident at once "
                                           "(raise sxc_set_workspace_limit)");
  return SXC_OK;
}

int finish_matrix(sxc_ctx* ctx, int nbf, double* dW) {
  dim3 blk(32, 8), grd((nbf + 31) / 32, (nbf + 7) / 8);
  PhaseTimer t(ctx, SXC_T_FINISH);
  k_mirror<<<grd, blk, 0, ctx->stream>>>(nbf, dW);
  LAUNCH_CHECK();
  return SXC_OK;
}

// mirror nmat matrices + nred ordered partial-sum reductions in one launch (k_finish)
int finish_build(sxc_ctx* ctx, int nbf, int nmat, double* dV, const double* part, int nlit, int nred, int pair_stride,
                 double* out) {
  dim3 blk(32, 8), grd((nbf + 31) / 32, (nbf + 7) / 8, std::max(nmat, 1));
  PhaseTimer t(ctx, SXC_T_FINISH);
  k_finish<<<grd, blk, 0, ctx->stream>>>(nmat > 0 ? nbf : 0, dV, part, nlit, nred, pair_stride, out);
  LAUNCH_CHECK();
  return SXC_OK;
}

int reduce_to(sxc_ctx* ctx, const double* part, int n, double* out) {
  PhaseTimer t(ctx, SXC_T_FINISH);
  k_reduce_partials<<<1, 256, 0, ctx->stream>>>(part, n, 1.0, 0, out);
  LAUNCH_CHECK();
  return SXC_OK;
}

// ---------------------------------------------------------------------------------------------- builds
int build_xc_device(sxc_ctx* ctx, int gh, int bh, int fh, int nspin, const double* dP, double thr, double* dVEN,
                    bool timed) {
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 (RESTRICTED) or 2 (UNRESTRICTED)");
  if (fh < 0 || fh >= (int)ctx->funcs.size()) return fail(ctx, SXC_ERR_INVALID, "invalid functional handle %d", fh);
  Plan* pp = nullptr;
  TRY(get_plan(ctx, gh, bh, &pp));
  Plan& p = *pp;
  Grid& g = *get_grid(ctx, gh);
  Basis& b = *get_basis(ctx, bh);
  const FuncView f = ctx->funcs[fh];
  TRY(ensure_point_arrays(ctx, g, false, nspin));
  ctx->stats = p.stats;
  begin_timing(ctx, timing_mode(ctx, timed));
  const int launches0 = ctx->launches;
  const size_t nb2 = (size_t)b.nbf * b.nbf;
  const long N = g.npts;
  double* parts = g.parts.as<double>();
  double* dens = g.dens.as<double>();
  double* pot = g.pot.as<double>();
  {
    PhaseTimer t_all(ctx, T_TOTAL);
    CU(cudaMemsetAsync(dVEN, 0, (nspin * nb2 + 2) * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(parts, 0, (size_t)3 * std::max(g.nlit, 1) * sizeof(double), ctx->stream));
    // The reference repeats the block prescreening with every evaluation (calculateBasisFunctionData :211-255); its outcome is a
    // pure function of grid and basis, both immutable behind their handles, so the lists made by k_screen when the plan was
    // created are used by every build.  With sxc_set_tile_cache the tiles of the previous build of the same plan are reused too.
    const bool cached = tiles_cached(ctx, p);
    for (const Chunk& c : p.chunks) {
      if (!cached) TRY(phase_basis(ctx, g, b, p, c));
      TRY(wait_p_ready(ctx));
      // UNRESTRICTED: the same block data is contracted with P_alpha and P_beta (MatrixOperatorToGridTransformer.h:128-146)
      for (int sp = 0; sp < nspin; ++sp)
        TRY(phase_density(ctx, g, b, p, c, dP + sp * nb2, dens + (size_t)4 * sp * N, true, nullptr));
      TRY(phase_functional(ctx, g, p, c, f, nspin, dens, 1.0, 0, pot, parts, parts + g.nlit));
      if (f.ncomp > 0)  // per spin, each with its own block-average test (ScalarOperatorToMatrixAdder.cpp:262-268)
        for (int sp = 0; sp < nspin; ++sp)
          TRY(phase_scatter(ctx, g, b, p, c, f.gga != 0, thr, pot + (size_t)4 * sp * N, dVEN + sp * nb2));
    }
    TRY(finish_build(ctx, b.nbf, nspin, dVEN, parts, g.nlit, 2, 2, dVEN + nspin * nb2));
    TRY(allreduce_result(ctx, g, dVEN, nspin * nb2 + 2));
  }
  ctx->timing = 0;
  ctx->stats.kernel_launches = ctx->launches - launches0;
  return SXC_OK;
}

// out = a + b (b may be null) on the blocks of a chunk, ncomp rows of length N
__global__ void k_add4(long N, int blocksize, int ncomp, const int* __restrict__ block_id, const double* __restrict__ a,
                       const double* __restrict__ b, double* __restrict__ out) {
  const long first = (long)block_id[blockIdx.x] * blocksize;
  const long n = min((long)blocksize, N - first);
  for (int c = 0; c < ncomp; ++c)
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
      const size_t k = (size_t)c * N + first + i;
      out[k] = a[k] + (b ? b[k] : 0.0);
    }
}

std::vector<int> env_cache_key(int nenv, const int* bE, int nspin, int tag) {
  std::vector<int> key(bE, bE + nenv);
  key.push_back(nspin);
  key.push_back(tag);
  return key;
}
// the grid holds sum rho_env of exactly this environment state and E[rho_env_i] of every requested functional
bool env_cache_hit(const Grid& g, int nfunc, const int* fhs, int nenv, const int* bE, int nspin, int tag) {
  if (!tag || !g.env_valid || g.env_key != env_cache_key(nenv, bE, nspin, tag)) return false;
  for (int k = 0; k < nfunc; ++k)
    if (!g.env_energy.count(fhs[k])) return false;
  return true;
}

// One device pass for nfunc functionals on the same densities (FDEPotentials::getFockMatrix, potentials/bundles/FDEPotentials.cpp:43-61,
// adds the non-additive XC and the non-additive kinetic potential of the same active / environment densities): rho_act and
// sum rho_env are built ONCE; per functional the energies E[rho_tot], E[rho_act], E[rho_env_i] are kept apart.
//   sum_mode != 0: the potentials are summed on the grid and scattered once      -> dVE = [nspin nbA^2 | nfunc x (2 + nenv)]
//   sum_mode == 0: one scatter per functional, every object gets its own matrix   -> dVE = [nfunc x nspin nbA^2 | nfunc x (2 + nenv)]
// nfunc == 1 is NAddFuncPotential<SCFMode>::getMatrix + getEnergy (NAddFuncPotential.cpp:192-326).
int build_nadd_device(sxc_ctx* ctx, int gh, int nfunc, const int* fhs, int nspin, int bA, const double* dPA, int nenv,
                      const int* bE, const double* const* dPE, int frozen, double thr, int sum_mode, double* dVE, bool timed) {
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 (RESTRICTED) or 2 (UNRESTRICTED)");
  if (nfunc < 1 || nfunc > 8) return fail(ctx, SXC_ERR_INVALID, "1 to 8 functionals per pass");
  for (int k = 0; k < nfunc; ++k)
    if (fhs[k] < 0 || fhs[k] >= (int)ctx->funcs.size()) return fail(ctx, SXC_ERR_INVALID, "invalid functional handle %d", fhs[k]);
  if (nenv < 0) return fail(ctx, SXC_ERR_INVALID, "nenv < 0");
  Grid* gp = get_grid(ctx, gh);
  Basis* ba = get_basis(ctx, bA);
  if (!gp || !ba) return fail(ctx, SXC_ERR_INVALID, "invalid grid or active basis handle");
  Grid& g = *gp;
  Plan* pa = nullptr;
  TRY(get_plan(ctx, gh, bA, &pa));  // the active system fixes the block ownership
  for (int i = 0; i < nenv; ++i) {
    Plan* pe = nullptr;
    if (!get_basis(ctx, bE[i])) return fail(ctx, SXC_ERR_INVALID, "invalid environment basis handle %d", bE[i]);
    TRY(get_plan(ctx, gh, bE[i], &pe));
  }
  TRY(ensure_point_arrays(ctx, g, true, nspin, 2 * nfunc, sum_mode ? nfunc : 1));
  CU(ctx->scratch.ensure(64 * sizeof(double)));
  const long N = g.npts;
  const int ncomp = 4 * nspin;
  const int nlit = std::max(g.nlit, 1);
  const size_t nb2 = (size_t)ba->nbf * ba->nbf;
  const int nmat = sum_mode ? 1 : nfunc;
  const size_t nV = (size_t)nmat * nspin * nb2;
  const int ne = 2 + nenv;  // energies per functional
  double* parts = g.parts.as<double>();
  double* dens = g.dens.as<double>();
  double* pot = g.pot.as<double>();
  ctx->stats = pa->stats;
  begin_timing(ctx, timing_mode(ctx, timed));
  const int launches0 = ctx->launches;
  bool any_gga = false, any_comp = false;
  for (int k = 0; k < nfunc; ++k) {
    any_gga |= ctx->funcs[fhs[k]].gga != 0;
    any_comp |= ctx->funcs[fhs[k]].ncomp > 0;
  }
  {
    PhaseTimer t_all(ctx, T_TOTAL);
    CU(cudaMemsetAsync(dVE, 0, (nV + (size_t)nfunc * ne) * sizeof(double), ctx->stream));
    TRY(wait_p_ready(ctx));

    // environment: rho_env on the supersystem grid, summed; E[rho_env_i] (NAddEnergyHelper, NAddFuncPotential.cpp:502-516)
    // (the summed density does not depend on the functional, the environment energies do: a functional seen for the first time
    // with a frozen environment repeats the pass once, afterwards all NAdd objects of an iteration are served from the cache).
    // The cache is per grid; `frozen` is the caller's tag of the environment STATE (any non-zero value): two NAdd objects on
    // one grid whose environments share basis handles but hold different densities must use different tags.
    const std::vector<int> key = env_cache_key(nenv, bE, nspin, frozen);
    const bool same_env = frozen && g.env_valid && g.env_key == key;
    const bool reuse = env_cache_hit(g, nfunc, fhs, nenv, bE, nspin, frozen);
    if (!same_env) g.env_energy.clear();
    if (!reuse) {
      CU(cudaMemsetAsync(g.envsum.p, 0, (size_t)ncomp * N * sizeof(double), ctx->stream));
      for (int k = 0; k < nfunc; ++k) g.env_energy[fhs[k]].assign(nenv, 0.0);
      if ((size_t)nenv * nfunc > 64) return fail(ctx, SXC_ERR_UNSUPPORTED, "more than 64 environment energies per pass");
      for (int i = 0; i < nenv; ++i) {
        Basis* be = get_basis(ctx, bE[i]);
        const size_t ne2 = (size_t)be->nbf * be->nbf;
        Plan* pe = nullptr;
        TRY(get_plan(ctx, gh, bE[i], &pe));
        CU(cudaMemsetAsync(parts, 0, (size_t)nfunc * nlit * sizeof(double), ctx->stream));
        for (const Chunk& c : pe->chunks) {
          TRY(phase_basis(ctx, g, *be, *pe, c));
          for (int sp = 0; sp < nspin; ++sp)
            TRY(phase_density(ctx, g, *be, *pe, c, dPE[i] + sp * ne2, dens + (size_t)4 * sp * N, true, nullptr));
          if (pe->nown) {
            PhaseTimer t(ctx, SXC_T_DENSITY);
            k_add4<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, ncomp, pe->block_id.as<int>() + c.slot0,
                                                      g.envsum.as<double>(), dens, g.envsum.as<double>());
            LAUNCH_CHECK();
          }
          // energy only: the potential goes to g.pot and is overwritten later
          for (int k = 0; k < nfunc; ++k)
            TRY(phase_functional(ctx, g, *pe, c, ctx->funcs[fhs[k]], nspin, dens, 1.0, 0, pot, parts + (size_t)k * nlit, nullptr));
        }
        for (int k = 0; k < nfunc; ++k)
          TRY(reduce_to(ctx, parts + (size_t)k * nlit, g.nlit, ctx->scratch.as<double>() + (size_t)k * nenv + i));
      }
      if (nenv > 0) {
        std::vector<double> h((size_t)nenv * nfunc);
        CU(cudaMemcpyAsync(h.data(), ctx->scratch.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < nfunc; ++k)
          for (int i = 0; i < nenv; ++i) g.env_energy[fhs[k]][i] = h[(size_t)k * nenv + i];
      }
      g.env_key = key;
      g.env_valid = true;
    }
    if (nenv > 0)
      for (int k = 0; k < nfunc; ++k)
        CU(cudaMemcpyAsync(dVE + nV + (size_t)k * ne + 2, g.env_energy[fhs[k]].data(), nenv * sizeof(double),
                           cudaMemcpyHostToDevice, ctx->stream));

    // active system: rho_A, rho_tot = rho_A + sum_env, v = v[rho_tot] - v[rho_A]  (NAddFuncPotential.cpp:197-225)
    Plan& p = *pa;
    CU(cudaMemsetAsync(parts, 0, (size_t)2 * nfunc * nlit * sizeof(double), ctx->stream));
    // summed matrix: the functionals' potentials sit side by side ([nfunc][4 nspin][N]) and are added inside k_form_g, each under
    // its own block-average test (an LDA functional next to a GGA one leaves its gradient rows zero)
    const size_t pot_stride = (size_t)ncomp * N;
    if (sum_mode && nfunc > 1) CU(cudaMemsetAsync(pot, 0, (size_t)nfunc * pot_stride * sizeof(double), ctx->stream));
    const bool cached = tiles_cached(ctx, p);  // (frozen environment: the active system's tiles survive from call to call)
    for (const Chunk& c : p.chunks) {
      if (!cached) TRY(phase_basis(ctx, g, *ba, p, c));
      for (int sp = 0; sp < nspin; ++sp)
        TRY(phase_density(ctx, g, *ba, p, c, dPA + sp * nb2, dens + (size_t)4 * sp * N, true, nullptr));
      if (p.nown) {
        PhaseTimer t(ctx, SXC_T_DENSITY);
        k_add4<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, ncomp, p.block_id.as<int>() + c.slot0, dens,
                                                  g.envsum.as<double>(), g.tot.as<double>());
        LAUNCH_CHECK();
      }
      for (int k = 0; k < nfunc; ++k) {
        const FuncView f = ctx->funcs[fhs[k]];
        double* pot_k = pot + ((sum_mode && nfunc > 1) ? (size_t)k * pot_stride : 0);
        TRY(phase_functional(ctx, g, p, c, f, nspin, g.tot.as<double>(), 1.0, 0, pot_k, parts + (size_t)(2 * k) * nlit, nullptr));
        TRY(phase_functional(ctx, g, p, c, f, nspin, dens, -1.0, 1, pot_k, parts + (size_t)(2 * k + 1) * nlit, nullptr));
        if (!sum_mode && f.ncomp > 0)
          for (int sp = 0; sp < nspin; ++sp)
            TRY(phase_scatter(ctx, g, *ba, p, c, f.gga != 0, thr, pot + (size_t)4 * sp * N,
                              dVE + ((size_t)k * nspin + sp) * nb2));
      }
      if (sum_mode && any_comp)
        for (int sp = 0; sp < nspin; ++sp)
          TRY(phase_scatter(ctx, g, *ba, p, c, any_gga, thr, pot + (size_t)4 * sp * N, dVE + sp * nb2, nfunc, pot_stride));
    }
    TRY(finish_build(ctx, ba->nbf, nmat * nspin, dVE, parts, g.nlit, 2 * nfunc, ne, dVE + nV));
    TRY(allreduce_result(ctx, g, dVE, nV + (size_t)nfunc * ne));
  }
  ctx->timing = 0;
  ctx->stats.kernel_launches = ctx->launches - launches0;
  return SXC_OK;
}

// XC nuclear gradient (row f-3): FuncPotential<SCFMode>::getGeomGradients (FuncPotential.cpp:114-239).
// d_gfunc [nbf][3] receives t[nu, c] (gradient_kernels.cuh); the caller folds functions into atoms with the factor -2.
// With nenv > 0 it is NAddFuncPotential<SCFMode>::getGeomGradients (NAddFuncPotential.cpp:329-493): the potential on the
// grid is v[rho_act + sum rho_env] - v[rho_act] (:331-341), contracted with the ACTIVE density matrix.
int build_gradient_device(sxc_ctx* ctx, int gh, int bh, int fh, int nspin, const double* dP, double* d_gfunc, int nenv = 0,
                          const int* bE = nullptr, const double* const* dPE = nullptr) {
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 (RESTRICTED) or 2 (UNRESTRICTED)");
  if (fh < 0 || fh >= (int)ctx->funcs.size()) return fail(ctx, SXC_ERR_INVALID, "invalid functional handle %d", fh);
  if (!get_grid(ctx, gh) || !get_basis(ctx, bh)) return fail(ctx, SXC_ERR_INVALID, "invalid grid (%d) or basis (%d) handle", gh, bh);
  // (the regular plan of the active basis fixes the block ownership before any environment plan is made)
  Plan* pp = nullptr;
  TRY(get_plan(ctx, gh, bh, &pp, GRAD_TILE_COMPS));
  Plan& p = *pp;
  Grid& g = *get_grid(ctx, gh);
  Basis& b = *get_basis(ctx, bh);
  const FuncView f = ctx->funcs[fh];
  const bool nadd = nenv > 0;
  for (int i = 0; i < nenv; ++i) {
    Plan* pe = nullptr;
    if (!get_basis(ctx, bE[i])) return fail(ctx, SXC_ERR_INVALID, "invalid environment basis handle %d", bE[i]);
    TRY(get_plan(ctx, gh, bE[i], &pe));
  }
  TRY(ensure_point_arrays(ctx, g, nadd, nspin));
  ctx->stats = p.stats;
  begin_timing(ctx, timing_mode(ctx, false));
  const int launches0 = ctx->launches;
  const size_t nb2 = (size_t)b.nbf * b.nbf;
  const long N = g.npts;
  double* parts = g.parts.as<double>();
  double* dens = g.dens.as<double>();
  double* pot = g.pot.as<double>();
  {
    PhaseTimer t_all(ctx, T_TOTAL);
    CU(cudaMemsetAsync(d_gfunc, 0, (size_t)b.nbf * 3 * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(parts, 0, (size_t)3 * std::max(g.nlit, 1) * sizeof(double), ctx->stream));
    if (nadd) {  // sum of the environment densities on the grid (SupersystemDensityOnGridController::updateData)
      TRY(wait_p_ready(ctx));
      g.env_valid = false;  // the cache of sxc_build_nadd is overwritten
      CU(cudaMemsetAsync(g.envsum.p, 0, (size_t)4 * nspin * N * sizeof(double), ctx->stream));
      for (int i = 0; i < nenv; ++i) {
        Basis* be = get_basis(ctx, bE[i]);
        const size_t ne2 = (size_t)be->nbf * be->nbf;
        Plan* pe = nullptr;
        TRY(get_plan(ctx, gh, bE[i], &pe));
        for (const Chunk& c : pe->chunks) {
          if (c.nslots == 0) continue;
          TRY(phase_basis(ctx, g, *be, *pe, c));
          for (int sp = 0; sp < nspin; ++sp)
            TRY(phase_density(ctx, g, *be, *pe, c, dPE[i] + sp * ne2, dens + (size_t)4 * sp * N, true, nullptr));
          PhaseTimer t(ctx, SXC_T_DENSITY);
          k_add4<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, 4 * nspin, pe->block_id.as<int>() + c.slot0,
                                                    g.envsum.as<double>(), dens, g.envsum.as<double>());
          LAUNCH_CHECK();
        }
      }
    }
    for (const Chunk& c : p.chunks) {
      if (c.nslots == 0) continue;
      TRY(phase_basis(ctx, g, b, p, c));
      TRY(wait_p_ready(ctx));
      for (int sp = 0; sp < nspin; ++sp)
        TRY(phase_density(ctx, g, b, p, c, dP + sp * nb2, dens + (size_t)4 * sp * N, true, nullptr));
      if (nadd) {  // v = v[rho_tot] - v[rho_act]
        {
          PhaseTimer t(ctx, SXC_T_DENSITY);
          k_add4<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, 4 * nspin, p.block_id.as<int>() + c.slot0, dens,
                                                    g.envsum.as<double>(), g.tot.as<double>());
          LAUNCH_CHECK();
        }
        TRY(phase_functional(ctx, g, p, c, f, nspin, g.tot.as<double>(), 1.0, 0, pot, parts, nullptr));
        TRY(phase_functional(ctx, g, p, c, f, nspin, dens, -1.0, 1, pot, parts + g.nlit, nullptr));
      } else {
        TRY(phase_functional(ctx, g, p, c, f, nspin, dens, 1.0, 0, pot, parts, parts + g.nlit));
      }
      if (f.ncomp == 0) continue;
      for (int sp = 0; sp < nspin; ++sp) {
        const double* pot4 = pot + (size_t)4 * sp * N;
        const bool gga = f.gga != 0;
        {  // K = a phi + b . grad phi into slot 4 (no block-average test in the gradient: threshold 0)
          PhaseTimer t(ctx, SXC_T_FORM_G);
          k_form_g<<<form_g_grid(ctx, c.nslots), 256, 0, ctx->stream>>>(g.view(), p.view(), p.order.as<int>() + c.order_off, 0.0, 1.0, pot4,
                                                      gga ? pot4 + N : nullptr, gga ? pot4 + 2 * N : nullptr,
                                                      gga ? pot4 + 3 * N : nullptr, 1, 0, ctx->phi.as<double>(), p.skip.as<int>());
          LAUNCH_CHECK();
        }
        PhaseTimer t(ctx, SXC_T_SCATTER);
        k_grad_contract<<<c.nslots, dens::THREADS, dens::smem_bytes(p.s_pad_max), ctx->stream>>>(
            p.view(), b.nbf, dP + sp * nb2, p.order.as<int>() + c.order_off, ctx->phi.as<double>(), 4, 1, d_gfunc);
        LAUNCH_CHECK();
        if (gga) {
          k_hessq<<<c.nslots, BASIS_GROUPS * BP, 0, ctx->stream>>>(g.view(), b.view(), p.view(), c.slot0,
                                                                   p.order.as<int>() + c.order_off, pot4 + N, pot4 + 2 * N,
                                                                   pot4 + 3 * N, ctx->phi.as<double>());
          LAUNCH_CHECK();
          k_grad_contract<<<c.nslots, dens::THREADS, dens::smem_bytes(p.s_pad_max), ctx->stream>>>(
              p.view(), b.nbf, dP + sp * nb2, p.order.as<int>() + c.order_off, ctx->phi.as<double>(), 0, 5, d_gfunc);
          LAUNCH_CHECK();
        }
      }
    }
    TRY(allreduce_result(ctx, g, d_gfunc, (size_t)b.nbf * 3));
  }
  ctx->timing = 0;
  ctx->stats.kernel_launches = ctx->launches - launches0;
  return SXC_OK;
}

}  // namespace

// ================================================================================================ C ABI
// ABFuncPotential<SCFMode>::getMatrix (potentials/ABFockMatrixConstruction/ABFuncPotential.cpp:54-160): the densities of all
// (basis_C, P_C) pairs are summed on the grid, the functional is evaluated once, its potential is scattered into the
// nbf_A x nbf_B matrix.  dVE = [nspin * nA * nB | E_xc | N_el].
int build_ab_device(sxc_ctx* ctx, int gh, int fh, int nspin, int bA, int bB, int ndens, const int* bC,
                    const double* const* dPC, double thr, double* dVE) {
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 (RESTRICTED) or 2 (UNRESTRICTED)");
  if (fh < 0 || fh >= (int)ctx->funcs.size()) return fail(ctx, SXC_ERR_INVALID, "invalid functional handle %d", fh);
  if (ndens <= 0) return fail(ctx, SXC_ERR_INVALID, "at least one density matrix is needed");
  Plan *pa = nullptr, *pb = nullptr;
  TRY(ab_plans(ctx, gh, bA, bB, &pa, &pb));
  Grid& g = *get_grid(ctx, gh);
  Basis& ba = *get_basis(ctx, bA);
  Basis& bb = *get_basis(ctx, bB);
  const FuncView f = ctx->funcs[fh];
  for (int i = 0; i < ndens; ++i) {
    Plan* pc = nullptr;
    if (!get_basis(ctx, bC[i])) return fail(ctx, SXC_ERR_INVALID, "invalid density basis handle %d", bC[i]);
    TRY(get_plan(ctx, gh, bC[i], &pc));
  }
  TRY(ensure_point_arrays(ctx, g, true, nspin));
  const long N = g.npts;
  const int ncomp = 4 * nspin;
  const size_t nab = (size_t)ba.nbf * bb.nbf;
  double* parts = g.parts.as<double>();
  double* dens = g.dens.as<double>();
  double* tot = g.tot.as<double>();
  double* pot = g.pot.as<double>();
  ctx->stats = pa->stats;
  begin_timing(ctx, timing_mode(ctx, true));
  const int launches0 = ctx->launches;
  {
    PhaseTimer t_all(ctx, T_TOTAL);
    CU(cudaMemsetAsync(dVE, 0, (nspin * nab + 2) * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(parts, 0, (size_t)3 * std::max(g.nlit, 1) * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(tot, 0, (size_t)ncomp * N * sizeof(double), ctx->stream));
    TRY(wait_p_ready(ctx));
    for (int i = 0; i < ndens; ++i) {  // rho += rho_C, grad rho += grad rho_C  (ABFuncPotential.cpp:66-90)
      Basis& bc = *get_basis(ctx, bC[i]);
      const size_t nc2 = (size_t)bc.nbf * bc.nbf;
      Plan* pc = nullptr;
      TRY(get_plan(ctx, gh, bC[i], &pc));
      for (const Chunk& c : pc->chunks) {
        TRY(phase_basis(ctx, g, bc, *pc, c));
        for (int sp = 0; sp < nspin; ++sp)
          TRY(phase_density(ctx, g, bc, *pc, c, dPC[i] + sp * nc2, dens + (size_t)4 * sp * N, true, nullptr));
        if (pc->nown) {
          PhaseTimer t(ctx, SXC_T_DENSITY);
          k_add4<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, ncomp, pc->block_id.as<int>() + c.slot0, tot, dens, tot);
          LAUNCH_CHECK();
        }
      }
    }
    TRY(phase_functional(ctx, g, *pa, pa->chunks.empty() ? Chunk() : pa->chunks[0], f, nspin, tot, 1.0, 0, pot, parts,
                         parts + g.nlit));
    if (f.ncomp > 0 && !pa->chunks.empty()) {
      TRY(phase_basis(ctx, g, ba, *pa, pa->chunks[0]));
      TRY(phase_basis(ctx, g, bb, *pb, pb->chunks[0], &ctx->phi2));
      for (int sp = 0; sp < nspin; ++sp)
        TRY(phase_scatter_ab(ctx, g, ba, *pa, *pb, f.gga != 0, thr, pot + (size_t)4 * sp * N, dVE + sp * nab));
    }
    TRY(reduce_to(ctx, parts, g.nlit, dVE + nspin * nab));
    TRY(reduce_to(ctx, parts + g.nlit, g.nlit, dVE + nspin * nab + 1));
    TRY(allreduce_result(ctx, g, dVE, nspin * nab + 2));
  }
  ctx->timing = 0;
  ctx->stats.kernel_launches = ctx->launches - launches0;
  return SXC_OK;
}

// ABNAddFuncPotential<SCFMode>::getMatrix (potentials/ABFockMatrixConstruction/ABNAddFuncPotential.cpp:66-176): the
// non-additive potential v[rho_act + sum rho_env] - v[rho_act] (:150-170) scattered into the nbf_A x nbf_B matrix.
int build_ab_nadd_device(sxc_ctx* ctx, int gh, int fh, int nspin, int bA, int bB, int bAct, const double* dPact, int nenv,
                         const int* bE, const double* const* dPE, double thr, double* dV) {
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 (RESTRICTED) or 2 (UNRESTRICTED)");
  if (fh < 0 || fh >= (int)ctx->funcs.size()) return fail(ctx, SXC_ERR_INVALID, "invalid functional handle %d", fh);
  if (nenv < 0) return fail(ctx, SXC_ERR_INVALID, "nenv < 0");
  Plan *pa = nullptr, *pb = nullptr, *pact = nullptr;
  TRY(ab_plans(ctx, gh, bA, bB, &pa, &pb));
  if (!get_basis(ctx, bAct)) return fail(ctx, SXC_ERR_INVALID, "invalid active basis handle %d", bAct);
  TRY(get_plan(ctx, gh, bAct, &pact));
  for (int i = 0; i < nenv; ++i) {
    Plan* pe = nullptr;
    if (!get_basis(ctx, bE[i])) return fail(ctx, SXC_ERR_INVALID, "invalid environment basis handle %d", bE[i]);
    TRY(get_plan(ctx, gh, bE[i], &pe));
  }
  Grid& g = *get_grid(ctx, gh);
  Basis& ba = *get_basis(ctx, bA);
  Basis& bb = *get_basis(ctx, bB);
  const FuncView f = ctx->funcs[fh];
  TRY(ensure_point_arrays(ctx, g, true, nspin));
  const long N = g.npts;
  const int ncomp = 4 * nspin;
  const size_t nab = (size_t)ba.nbf * bb.nbf;
  double* parts = g.parts.as<double>();
  double* dens = g.dens.as<double>();
  double* tot = g.tot.as<double>();
  double* pot = g.pot.as<double>();
  ctx->stats = pa->stats;
  begin_timing(ctx, timing_mode(ctx, true));
  const int launches0 = ctx->launches;
  {
    PhaseTimer t_all(ctx, T_TOTAL);
    CU(cudaMemsetAsync(dV, 0, nspin * nab * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(parts, 0, (size_t)3 * std::max(g.nlit, 1) * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(tot, 0, (size_t)ncomp * N * sizeof(double), ctx->stream));
    g.env_valid = false;  // g.tot / g.dens are reused: the cache of sxc_build_nadd is gone
    TRY(wait_p_ready(ctx));
    // tot = sum of the environment densities (:73-104), then + rho_act; dens keeps rho_act (:107-147)
    for (int i = 0; i <= nenv; ++i) {
      const bool act = i == nenv;
      const int bh = act ? bAct : bE[i];
      const double* dP = act ? dPact : dPE[i];
      Basis& bc = *get_basis(ctx, bh);
      const size_t nc2 = (size_t)bc.nbf * bc.nbf;
      Plan* pc = nullptr;
      TRY(get_plan(ctx, gh, bh, &pc));
      for (const Chunk& c : pc->chunks) {
        TRY(phase_basis(ctx, g, bc, *pc, c));
        for (int sp = 0; sp < nspin; ++sp)
          TRY(phase_density(ctx, g, bc, *pc, c, dP + sp * nc2, dens + (size_t)4 * sp * N, true, nullptr));
        if (pc->nown) {
          PhaseTimer t(ctx, SXC_T_DENSITY);
          k_add4<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, ncomp, pc->block_id.as<int>() + c.slot0, tot, dens, tot);
          LAUNCH_CHECK();
        }
      }
    }
    if (f.ncomp > 0 && !pa->chunks.empty()) {
      const Chunk& c0 = pa->chunks[0];
      TRY(phase_functional(ctx, g, *pa, c0, f, nspin, tot, 1.0, 0, pot, parts, nullptr));
      TRY(phase_functional(ctx, g, *pa, c0, f, nspin, dens, -1.0, 1, pot, parts + g.nlit, nullptr));
      TRY(phase_basis(ctx, g, ba, *pa, c0));
      TRY(phase_basis(ctx, g, bb, *pb, pb->chunks[0], &ctx->phi2));
      for (int sp = 0; sp < nspin; ++sp)
        TRY(phase_scatter_ab(ctx, g, ba, *pa, *pb, f.gga != 0, thr, pot + (size_t)4 * sp * N, dV + sp * nab));
    }
    TRY(allreduce_result(ctx, g, dV, nspin * nab));
  }
  ctx->timing = 0;
  ctx->stats.kernel_launches = ctx->launches - launches0;
  return SXC_OK;
}

extern "C" {

int sxc_abi_version(void) { return 5; }

// host-only: the k_vmat round schedule of a block with n32 row groups (40 bytes per round, struct ScatterRound of
// scatter_kernel.cuh: ngroups, 7 pad, group[8], ta[8], tb[8], kmask[8]); returns the number of rounds
int sxc_debug_scatter_schedule(int n32, unsigned char* rounds40, int max_rounds) {
  if (n32 < 1 || n32 > 255) return SXC_ERR_INVALID;
  const std::vector<ScatterRound> r = build_scatter_schedule(n32);
  if (rounds40)
    for (int i = 0; i < (int)r.size() && i < max_rounds; ++i) std::memcpy(rounds40 + (size_t)40 * i, &r[i], 40);
  return (int)r.size();
}

// host-only: the k_vmat_tma round schedule (64 bytes per round, struct ScatterRound2 of scatter_tma.cuh) for a block with n32 row
// groups and ks k-steps per K chunk (2: chunks of 8 points, 4: chunks of 16 points); returns the number of rounds
int sxc_debug_scatter_schedule2(int n32, int ks, unsigned char* rounds64, int max_rounds) {
  if (n32 < 1 || n32 > 255 || (ks != 2 && ks != 4)) return SXC_ERR_INVALID;
  const std::vector<ScatterRound2> r = build_scatter_schedule2(n32, ks);
  if (rounds64)
    for (int i = 0; i < (int)r.size() && i < max_rounds; ++i) std::memcpy(rounds64 + (size_t)64 * i, &r[i], 64);
  return (int)r.size();
}

// host-only: contiguous ranges [bounds[r], bounds[r+1]) of nearly equal summed cost
int sxc_balance_ranges(int n, const double* cost, int world, int* bounds) {
  if (n < 0 || world < 1 || !bounds || (n > 0 && !cost)) return SXC_ERR_INVALID;
  double total = 0.0;
  for (int i = 0; i < n; ++i) total += cost[i];
  for (int r = 0; r <= world; ++r) bounds[r] = n;
  bounds[0] = 0;
  double acc = 0.0;
  int r = 1;
  for (int i = 0; i < n && r < world; ++i) {
    acc += cost[i];
    while (r < world && acc >= total * r / world) bounds[r++] = i + 1;
  }
  return SXC_OK;
}

int sxc_create(sxc_ctx** out, int device) {
  if (!out) return SXC_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return SXC_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return SXC_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SXC_ERR_CUDA;
  if (prop.major < 10) return SXC_ERR_UNSUPPORTED;  // sm_100a code only
  auto* ctx = new sxc_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (ctx->counters.ensure(sxc_ctx::NCOUNTERS * sizeof(int)) != cudaSuccess) {
    delete ctx;
    return SXC_ERR_NOMEM;
  }
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return SXC_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  if (const char* v = std::getenv("SXC_VMAT")) {  // development switch: 0 = cp.async kernel, 8 / 16 = TMA kernel with that K chunk
    const int k = std::atoi(v);
    if (k == 0 || k == 8 || k == 16 || k == 24 || k == 83) ctx->vmat_variant = k;  // 83: TKP 8 with a 3-stage ring (development)
  }
  if (const char* v = std::getenv("SXC_DENS")) ctx->dens_variant = std::atoi(v) ? 1 : 0;
  if (const char* v = std::getenv("SXC_COPY_THREADS")) ctx->copy_threads = std::max(0, std::min(16, std::atoi(v)));
  if (const char* v = std::getenv("SXC_DPF")) ctx->dens_prefetch = std::atoi(v);
  if (const char* v = std::getenv("SXC_FUNC")) ctx->func_variant = std::atoi(v);
  if (const char* v = std::getenv("SXC_BASIS")) ctx->basis_variant = std::max(0, std::min(2, std::atoi(v)));
  if (const char* v = std::getenv("SXC_SEG_WAVES")) ctx->seg_waves = std::max(1, std::atoi(v));
  if (const char* v = std::getenv("SXC_FG_LEAD")) ctx->fg_lead = std::atoi(v);
  if (const char* v = std::getenv("SXC_FG_SEG")) ctx->fg_segments = std::atoi(v);
  if (const char* v = std::getenv("SXC_FG_MODE")) ctx->fg_mode = std::atoi(v);
  if (const char* v = std::getenv("SXC_SMEM_PAD")) ctx->smem_pad = std::max(0, std::atoi(v));
  if (const char* v = std::getenv("SXC_DSEG")) ctx->dseg = std::max(1, std::atoi(v));
  if (const char* v = std::getenv("SXC_VSEG")) ctx->vseg = std::max(1, std::atoi(v));
  *out = ctx;
  return SXC_OK;
}

void sxc_destroy(sxc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) nccl().CommDestroy(ctx->comm);
  ctx->plans.clear();
  ctx->grids.clear();
  ctx->bases.clear();
  ctx->kstores.clear();
  ctx->phi.release();
  ctx->phi2.release();
  ctx->dP.release();
  ctx->dD.release();
  ctx->dOut.release();
  ctx->scratch.release();
  ctx->counters.release();
  for (auto& st : ctx->stamps) {
    cudaEventDestroy(st.a);
    cudaEventDestroy(st.b);
  }
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->chunk_events) cudaEventDestroy(e);
  if (ctx->h_up) cudaFreeHost(ctx->h_up);
  if (ctx->h_down) cudaFreeHost(ctx->h_down);
  ctx->copier.reset();
  if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* sxc_last_error(const sxc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sxc_set_stream(sxc_ctx* ctx, void* s) {
  if (!ctx) return SXC_ERR_INVALID;
  ctx->stream = s ? static_cast<cudaStream_t>(s) : ctx->own_stream;
  return SXC_OK;
}

int sxc_set_workspace_limit(sxc_ctx* ctx, int64_t bytes) {
  if (!ctx) return SXC_ERR_INVALID;
  ctx->ws_limit = bytes;
  ctx->plans.clear();  // chunking depends on the limit
  return SXC_OK;
}

int sxc_set_grid(sxc_ctx* ctx, int64_t npts, const double* xyz, const double* w, int blocksize, int* grid) {
  if (!ctx || !grid || npts <= 0 || !xyz || !w) return fail(ctx, SXC_ERR_INVALID, "sxc_set_grid: bad arguments");
  if (blocksize < 1 || blocksize > BP)
    return fail(ctx, SXC_ERR_UNSUPPORTED, "grid.blocksize %d outside 1..128", blocksize);
  CU(cudaSetDevice(ctx->device));
  auto g = std::make_unique<Grid>();
  g->npts = npts;
  g->blocksize = blocksize;
  g->nblocks = (int)((npts + blocksize - 1) / blocksize);
  g->nlit = (int)((npts + FUNC_BLOCK - 1) / FUNC_BLOCK);
  std::vector<double> soa((size_t)4 * npts);
  for (int64_t i = 0; i < npts; ++i) {  // Matrix3Xd interleaved -> SoA for coalesced loads
    soa[i] = xyz[3 * i];
    soa[npts + i] = xyz[3 * i + 1];
    soa[2 * npts + i] = xyz[3 * i + 2];
    soa[3 * npts + i] = w[i];
  }
  CU(g->xyzw.ensure(soa.size() * sizeof(double)));
  CU(cudaMemcpyAsync(g->xyzw.p, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->comm && blocksize == FUNC_BLOCK) {  // a context with a communicator evaluates its own shard of every grid
    g->rank = ctx->comm_rank;
    g->world = ctx->comm_world;
  }
  *grid = store_handle(ctx->grids, std::move(g));
  return SXC_OK;
}

int sxc_set_grid_shard(sxc_ctx* ctx, int grid, int rank, int world) {
  Grid* g = ctx ? get_grid(ctx, grid) : nullptr;
  if (!g || world < 1 || rank < 0 || rank >= world) return fail(ctx, SXC_ERR_INVALID, "sxc_set_grid_shard: bad arguments");
  if (world > 1 && g->blocksize != FUNC_BLOCK)
    return fail(ctx, SXC_ERR_UNSUPPORTED, "sharding needs grid.blocksize == 128");
  g->rank = rank;
  g->world = world;
  g->owner_fixed = false;
  g->env_valid = false;
  for (auto it = ctx->plans.begin(); it != ctx->plans.end();)
    it = (it->first.first == grid) ? ctx->plans.erase(it) : std::next(it);
  return SXC_OK;
}

int sxc_add_basis(sxc_ctx* ctx, int nshell, const int* l, const int* pure, const int* nprim, const int* first_bf,
                  const double* centre, const double* alpha, const double* coeff, const double* normfac,
                  double radial_threshold, int* basis) {
  if (!ctx || !basis || nshell <= 0 || !l || !pure || !nprim || !first_bf || !centre || !alpha || !coeff || !normfac)
    return fail(ctx, SXC_ERR_INVALID, "sxc_add_basis: bad arguments");
  CU(cudaSetDevice(ctx->device));
  auto b = std::make_unique<Basis>();
  b->nshell = nshell;
  b->radial_thr = radial_threshold;
  std::vector<int> ints((size_t)6 * nshell);
  size_t np = 0;
  int nbf = 0;
  for (int s = 0; s < nshell; ++s) {
    if (l[s] < 0 || l[s] > LMAX) return fail(ctx, SXC_ERR_UNSUPPORTED, "shell %d: angular momentum %d > %d", s, l[s], LMAX);
    const int nf = pure[s] ? 2 * l[s] + 1 : (l[s] + 1) * (l[s] + 2) / 2;
    ints[s] = l[s];
    ints[nshell + s] = pure[s] ? 1 : 0;
    ints[2 * nshell + s] = nprim[s];
    ints[3 * nshell + s] = (int)np;
    ints[4 * nshell + s] = first_bf[s];
    ints[5 * nshell + s] = nf;
    np += nprim[s];
    nbf = std::max(nbf, first_bf[s] + nf);
    b->lmax = std::max(b->lmax, l[s]);
  }
  b->nbf = nbf;
  b->nprim_total = np;
  std::vector<double> d((size_t)3 * nshell + 2 * np + nbf);
  std::memcpy(d.data(), centre, sizeof(double) * 3 * nshell);
  std::memcpy(d.data() + 3 * nshell, alpha, sizeof(double) * np);
  std::memcpy(d.data() + 3 * nshell + np, coeff, sizeof(double) * np);
  std::memcpy(d.data() + 3 * nshell + 2 * np, normfac, sizeof(double) * nbf);
  CU(b->ints.ensure(ints.size() * sizeof(int)));
  CU(b->dbl.ensure(d.size() * sizeof(double)));
  CU(cudaMemcpyAsync(b->ints.p, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(b->dbl.p, d.data(), d.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *basis = store_handle(ctx->bases, std::move(b));
  return SXC_OK;
}

int sxc_set_functional(sxc_ctx* ctx, int ncomp, const int* ids, const double* mix, int* func) {
  if (!ctx || !func || ncomp < 0 || (ncomp > 0 && (!ids || !mix)))
    return fail(ctx, SXC_ERR_INVALID, "sxc_set_functional: bad arguments");
  FuncView f{};
  for (int i = 0; i < ncomp; ++i) {
    if (ids[i] == 0) continue;  // BASIC_FUNCTIONALS::NONE is skipped, XCFun.cpp:729-730
    if (!functional_id_supported(ids[i]))
      return fail(ctx, SXC_ERR_UNSUPPORTED, "basic functional %d is not implemented", ids[i]);
    if (f.ncomp == MAX_COMP) return fail(ctx, SXC_ERR_UNSUPPORTED, "more than %d functional components", MAX_COMP);
    f.id[f.ncomp] = ids[i];
    f.mix[f.ncomp] = mix[i];
    f.gga |= functional_id_is_gga(ids[i]) ? 1 : 0;
    ++f.ncomp;
  }
  // identical definitions share one handle (FunctionalLibrary::calcData asks per call; handles never pile up)
  for (size_t i = 0; i < ctx->funcs.size(); ++i) {
    const FuncView& o = ctx->funcs[i];
    bool same = o.ncomp == f.ncomp && o.gga == f.gga;
    for (int k = 0; same && k < f.ncomp; ++k) same = o.id[k] == f.id[k] && o.mix[k] == f.mix[k];
    if (same) {
      *func = (int)i;
      return SXC_OK;
    }
  }
  ctx->funcs.push_back(f);
  *func = (int)ctx->funcs.size() - 1;
  return SXC_OK;
}

namespace {
// plans (screening lists, work items) that involve a released grid or basis go with it; tiles in the workspace lose their owner
void drop_plans(sxc_ctx* ctx, int grid, int basis) {
  for (auto it = ctx->plans.begin(); it != ctx->plans.end();) {
    const Plan& p = *it->second;
    if ((grid >= 0 && p.grid == grid) || (basis >= 0 && p.basis == basis)) {
      if (ctx->phi_owner == p.serial) ctx->phi_owner = 0;
      it = ctx->plans.erase(it);
    } else {
      ++it;
    }
  }
}
}  // namespace

int sxc_release_grid(sxc_ctx* ctx, int grid) {
  if (!ctx || !get_grid(ctx, grid)) return fail(ctx, SXC_ERR_INVALID, "sxc_release_grid: invalid grid handle %d", grid);
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  drop_plans(ctx, grid, -1);
  for (auto& ks : ctx->kstores)
    if (ks && ks->grid == grid) ks.reset();
  ctx->grids[grid].reset();  // frees the points, the per-point work arrays and the cached environment density
  return SXC_OK;
}

int sxc_release_basis(sxc_ctx* ctx, int basis) {
  if (!ctx || !get_basis(ctx, basis)) return fail(ctx, SXC_ERR_INVALID, "sxc_release_basis: invalid basis handle %d", basis);
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  drop_plans(ctx, -1, basis);
  for (auto& g : ctx->grids)  // a cached environment density that was built with this basis is no longer identifiable
    if (g && std::find(g->env_key.begin(), g->env_key.end(), basis) != g->env_key.end()) g->env_valid = false;
  ctx->bases[basis].reset();
  return SXC_OK;
}

int sxc_release_functional(sxc_ctx* ctx, int func) {
  if (!ctx || func < 0 || func >= (int)ctx->funcs.size())
    return fail(ctx, SXC_ERR_INVALID, "sxc_release_functional: invalid functional handle %d", func);
  return SXC_OK;  // definitions are a few bytes, shared between equal requests and kept until sxc_destroy
}

// page-locked host memory for callers that can place P / V in it (the copies of the host-buffer builds are then true DMA)
void* sxc_host_alloc(size_t bytes) {
  void* p = nullptr;
  return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void sxc_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---- multi-GPU --------------------------------------------------------------------------------------------------
int sxc_comm_unique_id(void* id128) {
  if (!id128) return SXC_ERR_INVALID;
  if (!nccl().load()) return SXC_ERR_UNSUPPORTED;
  nccl_unique_id id;
  if (nccl().GetUniqueId(&id) != NCCL_SUCCESS) return SXC_ERR_CUDA;
  std::memcpy(id128, &id, sizeof(id));
  return SXC_OK;
}

int sxc_comm_init_rank(sxc_ctx* ctx, int rank, int world, const void* id128) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, SXC_ERR_INVALID, "sxc_comm_init_rank: bad arguments");
  if (ctx->comm) return fail(ctx, SXC_ERR_INVALID, "sxc_comm_init_rank: the context already has a communicator");
  if (!nccl().load()) return fail(ctx, SXC_ERR_UNSUPPORTED, "%s", nccl().error.c_str());
  CU(cudaSetDevice(ctx->device));
  nccl_unique_id id;
  std::memcpy(&id, id128, sizeof(id));
  const int rc = nccl().CommInitRank(&ctx->comm, world, id, rank);
  if (rc != NCCL_SUCCESS) {
    ctx->comm = nullptr;
    return fail(ctx, SXC_ERR_CUDA, "ncclCommInitRank failed: %s", nccl().GetErrorString(rc));
  }
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  for (size_t gh = 0; gh < ctx->grids.size(); ++gh)  // grids that are already there become shards
    if (ctx->grids[gh] && ctx->grids[gh]->blocksize == FUNC_BLOCK) TRY(sxc_set_grid_shard(ctx, (int)gh, rank, world));
  return SXC_OK;
}

int sxc_comm_destroy(sxc_ctx* ctx) {
  if (!ctx) return SXC_ERR_INVALID;
  if (ctx->comm) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl().CommDestroy(ctx->comm);
    ctx->comm = nullptr;
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
  }
  return SXC_OK;
}

int sxc_comm_info(sxc_ctx* ctx, int* rank, int* world, int64_t* collectives, int* nccl_version) {
  if (!ctx) return SXC_ERR_INVALID;
  if (rank) *rank = ctx->comm_rank;
  if (world) *world = ctx->comm_world;
  if (collectives) *collectives = ctx->collectives;
  if (nccl_version) {
    *nccl_version = 0;
    if (ctx->comm) nccl().GetVersion(nccl_version);
  }
  return SXC_OK;
}

int sxc_build_xc_device(sxc_ctx* ctx, int grid, int basis, int func, int nspin, const double* d_P, double thr,
                        double* d_VEN) {
  if (!ctx || !d_P || !d_VEN) return fail(ctx, SXC_ERR_INVALID, "sxc_build_xc_device: bad arguments");
  CU(cudaSetDevice(ctx->device));
  return build_xc_device(ctx, grid, basis, func, nspin, d_P, thr, d_VEN, false);
}

int sxc_build_xc(sxc_ctx* ctx, int grid, int basis, int func, int nspin, const double* P, double thr, double* V,
                 double* E, double* nelec) {
  if (!ctx || !P || !E) return fail(ctx, SXC_ERR_INVALID, "sxc_build_xc: bad arguments");
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 or 2");
  Basis* b = get_basis(ctx, basis);
  if (!b) return fail(ctx, SXC_ERR_INVALID, "invalid basis handle %d", basis);
  CU(cudaSetDevice(ctx->device));
  const size_t nv = (size_t)nspin * b->nbf * b->nbf;
  CU(ctx->dP.ensure(nv * sizeof(double) + P_SLACK_BYTES));
  CU(ctx->dOut.ensure((nv + 2) * sizeof(double)));
  HostCall host_guard{ctx};
  // SXC_TRACE=1 (development): host-side timeline of the call on stderr, microseconds since entry
  static const bool trace = std::getenv("SXC_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto us = [&] { return (long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count(); };
  TRY(upload_async(ctx, ctx->dP.p, P, nv * sizeof(double)));
  TRY(upload_done(ctx));
  int rc = build_xc_device(ctx, grid, basis, func, nspin, ctx->dP.as<double>(), thr, ctx->dOut.as<double>(), true);
  ctx->timing = 0;
  if (rc != SXC_OK) return abort_build(ctx, rc);
  const long t_launched = us();
  std::vector<double> tail(2);
  CU(cudaMemcpyAsync(tail.data(), ctx->dOut.as<double>() + nv, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  const long t_built = us();
  if (V) TRY(result_d2h(ctx, V, ctx->dOut.as<double>(), nv));
  const long t_copied = us();
  CU(cudaStreamSynchronize(ctx->stream));
  collect_timers(ctx);
  if (trace)
    std::fprintf(stderr, "sxc_build_xc: launches issued %ld us, build done %ld us, V copied %ld us, return %ld us (device time %.1f us)\n",
                 t_launched, t_built, t_copied, us(), 1e3 * ctx->stats.ms_total);
  *E = tail[0];
  if (nelec) *nelec = tail[1];
  return SXC_OK;
}

int sxc_build_nadd_multi_device(sxc_ctx* ctx, int grid, int nfunc, const int* funcs, int nspin, int basis_act,
                                const double* d_P_act, int nenv, const int* basis_env, const double* const* d_P_env,
                                int env_frozen, double thr, int sum_matrices, double* d_VE) {
  if (!ctx || !funcs || !d_P_act || !d_VE || (nenv > 0 && (!basis_env || !d_P_env)))
    return fail(ctx, SXC_ERR_INVALID, "sxc_build_nadd_multi_device: bad arguments");
  CU(cudaSetDevice(ctx->device));
  return build_nadd_device(ctx, grid, nfunc, funcs, nspin, basis_act, d_P_act, nenv, basis_env, d_P_env, env_frozen, thr,
                           sum_matrices, d_VE, false);
}

int sxc_build_nadd_device(sxc_ctx* ctx, int grid, int func, int nspin, int basis_act, const double* d_P_act, int nenv,
                          const int* basis_env, const double* const* d_P_env, int env_frozen, double thr, double* d_VE) {
  return sxc_build_nadd_multi_device(ctx, grid, 1, &func, nspin, basis_act, d_P_act, nenv, basis_env, d_P_env, env_frozen, thr,
                                     0, d_VE);
}

int sxc_build_nadd_multi(sxc_ctx* ctx, int grid, int nfunc, const int* funcs, int nspin, int basis_act, const double* P_act,
                         int nenv, const int* basis_env, const double* const* P_env, int env_frozen, double thr,
                         int sum_matrices, double* V_act, double* E) {
  if (!ctx || !funcs || !P_act || !E || nfunc < 1 || nenv < 0 || (nenv > 0 && (!basis_env || !P_env)))
    return fail(ctx, SXC_ERR_INVALID, "sxc_build_nadd_multi: bad arguments");
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 or 2");
  Basis* ba = get_basis(ctx, basis_act);
  Grid* g = get_grid(ctx, grid);
  if (!ba || !g) return fail(ctx, SXC_ERR_INVALID, "invalid grid (%d) or active basis (%d) handle", grid, basis_act);
  CU(cudaSetDevice(ctx->device));
  const size_t nvA = (size_t)nspin * ba->nbf * ba->nbf;
  const size_t nV = (sum_matrices ? 1 : (size_t)nfunc) * nvA;
  const size_t nE = (size_t)nfunc * (2 + nenv);
  size_t total = nvA;
  std::vector<size_t> offs(nenv);
  for (int i = 0; i < nenv; ++i) {
    Basis* be = get_basis(ctx, basis_env[i]);
    if (!be) return fail(ctx, SXC_ERR_INVALID, "invalid environment basis handle %d", basis_env[i]);
    offs[i] = total;
    total += (size_t)nspin * be->nbf * be->nbf;
  }
  CU(ctx->dP.ensure(total * sizeof(double) + P_SLACK_BYTES));
  CU(ctx->dOut.ensure((nV + nE) * sizeof(double)));
  HostCall host_guard{ctx};
  TRY(upload_async(ctx, ctx->dP.p, P_act, nvA * sizeof(double)));
  // a frozen environment whose density and energies are cached on the grid is not uploaded again
  const bool env_cached = env_cache_hit(*g, nfunc, funcs, nenv, basis_env, nspin, env_frozen);
  std::vector<const double*> dpe(nenv);
  for (int i = 0; i < nenv; ++i) {
    Basis* be = get_basis(ctx, basis_env[i]);
    dpe[i] = ctx->dP.as<double>() + offs[i];
    if (!env_cached)
      TRY(upload_async(ctx, ctx->dP.as<double>() + offs[i], P_env[i], (size_t)nspin * be->nbf * be->nbf * sizeof(double)));
  }
  int rc = build_nadd_device(ctx, grid, nfunc, funcs, nspin, basis_act, ctx->dP.as<double>(), nenv, basis_env, dpe.data(),
                             env_frozen, thr, sum_matrices, ctx->dOut.as<double>(), true);
  ctx->timing = 0;
  if (rc != SXC_OK) return abort_build(ctx, rc);
  CU(cudaMemcpyAsync(E, ctx->dOut.as<double>() + nV, nE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (V_act) TRY(result_d2h(ctx, V_act, ctx->dOut.as<double>(), nV));
  CU(cudaStreamSynchronize(ctx->stream));
  collect_timers(ctx);
  return SXC_OK;
}

int sxc_build_nadd(sxc_ctx* ctx, int grid, int func, int nspin, int basis_act, const double* P_act, int nenv,
                   const int* basis_env, const double* const* P_env, int env_frozen, double thr, double* V_act, double* E) {
  return sxc_build_nadd_multi(ctx, grid, 1, &func, nspin, basis_act, P_act, nenv, basis_env, P_env, env_frozen, thr, 0, V_act, E);
}

int sxc_density_on_grid(sxc_ctx* ctx, int grid, int basis, const double* P, double* rho, double* gx, double* gy,
                        double* gz) {
  if (!ctx || !P || !rho) return fail(ctx, SXC_ERR_INVALID, "sxc_density_on_grid: bad arguments");
  CU(cudaSetDevice(ctx->device));
  Plan* pp = nullptr;
  TRY(get_plan(ctx, grid, basis, &pp));
  Grid& g = *get_grid(ctx, grid);
  Basis& b = *get_basis(ctx, basis);
  TRY(ensure_point_arrays(ctx, g, false, 1));
  const size_t nb2 = (size_t)b.nbf * b.nbf;
  const long N = g.npts;
  CU(ctx->dP.ensure(nb2 * sizeof(double) + P_SLACK_BYTES));
  CU(cudaMemcpyAsync(ctx->dP.p, P, nb2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(g.dens.p, 0, (size_t)4 * N * sizeof(double), ctx->stream));
  for (const Chunk& c : pp->chunks) {
    TRY(phase_basis(ctx, g, b, *pp, c));
    TRY(phase_density(ctx, g, b, *pp, c, ctx->dP.as<double>(), g.dens.as<double>(), gx != nullptr, nullptr));
  }
  const double* d = g.dens.as<double>();
  CU(cudaMemcpyAsync(rho, d, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (gx) {
    CU(cudaMemcpyAsync(gx, d + N, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(gy, d + 2 * N, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(gz, d + 3 * N, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return SXC_OK;
}

int sxc_basis_on_grid(sxc_ctx* ctx, int grid, int basis, int block, double* val, double* dx, double* dy, double* dz,
                      int* negligible, int* n_out) {
  if (!ctx || !val || !negligible) return fail(ctx, SXC_ERR_INVALID, "sxc_basis_on_grid: bad arguments");
  CU(cudaSetDevice(ctx->device));
  Plan* pp = nullptr;
  TRY(get_plan(ctx, grid, basis, &pp));
  Plan& p = *pp;
  Grid& g = *get_grid(ctx, grid);
  Basis& b = *get_basis(ctx, basis);
  const int q = block - g.own_first;
  if (q < 0 || q >= p.nown) return fail(ctx, SXC_ERR_INVALID, "block %d is not owned by this context", block);
  const Chunk* ch = nullptr;
  for (const Chunk& c : p.chunks)
    if (q >= c.slot0 && q < c.slot0 + c.nslots) ch = &c;
  ctx->phi_owner = 0;  // the single block is evaluated into the workspace: cached tiles are gone
  CU(ctx->phi.ensure(ch->doubles * sizeof(double)));
  k_basis<1><<<1, BASIS_GROUPS * BP, 0, ctx->stream>>>(g.view(), b.view(), p.view(), q, nullptr, ctx->phi.as<double>());
  LAUNCH_CHECK();
  const int s = p.h_s[q], sp = p.h_s_pad[q];
  std::vector<double> tile((size_t)4 * sp * BP);
  std::vector<int> sig(std::max(s, 1));
  std::vector<long long> off(1);
  CU(cudaMemcpyAsync(off.data(), p.phi_off.as<long long>() + q, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpyAsync(tile.data(), ctx->phi.as<double>() + off[0], tile.size() * sizeof(double), cudaMemcpyDeviceToHost,
                     ctx->stream));
  if (s) CU(cudaMemcpyAsync(sig.data(), p.sig_bf.as<int>() + (size_t)q * p.nbf_pad, s * sizeof(int), cudaMemcpyDeviceToHost,
                            ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const long first = (long)block * g.blocksize;
  const int n = (int)std::min<long>(g.blocksize, g.npts - first);
  if (n_out) *n_out = n;
  double* outs[4] = {val, dx, dy, dz};
  for (int c = 0; c < 4; ++c)
    if (outs[c]) std::fill(outs[c], outs[c] + (size_t)n * b.nbf, 0.0);
  for (int i = 0; i < b.nbf; ++i) negligible[i] = 1;
  for (int c = 0; c < s; ++c) {
    negligible[sig[c]] = 0;
    for (int comp = 0; comp < 4; ++comp) {
      if (!outs[comp]) continue;
      const double* src = tile.data() + ((size_t)comp * sp + c) * BP;
      std::copy(src, src + n, outs[comp] + (size_t)sig[c] * n);
    }
  }
  return SXC_OK;
}

// SupersystemDensityOnGridController::updateData (data/grid/SupersystemDensityOnGridController.cpp:95-193): the densities (and
// gradients) of several subsystems, each in its own basis, summed on the common grid - the stage NAddFuncPotential uses for
// rho_tot.  Subsystems are added in the order given, starting from zero, as the reference does; host outputs [N].
int sxc_supersystem_density_on_grid(sxc_ctx* ctx, int grid, int ndens, const int* basis, const double* const* P, double* rho,
                                    double* gx, double* gy, double* gz) {
  if (!ctx || ndens < 1 || !basis || !P || !rho) return fail(ctx, SXC_ERR_INVALID, "sxc_supersystem_density_on_grid: bad arguments");
  CU(cudaSetDevice(ctx->device));
  Grid* gp = get_grid(ctx, grid);
  if (!gp) return fail(ctx, SXC_ERR_INVALID, "invalid grid handle %d", grid);
  Grid& g = *gp;
  for (int i = 0; i < ndens; ++i) {
    Plan* pp = nullptr;
    if (!get_basis(ctx, basis[i]) || !P[i]) return fail(ctx, SXC_ERR_INVALID, "invalid basis handle or matrix of density %d", i);
    TRY(get_plan(ctx, grid, basis[i], &pp));
  }
  TRY(ensure_point_arrays(ctx, g, true, 1));
  const long N = g.npts;
  g.env_valid = false;  // the accumulator is the frozen-environment cache of sxc_build_nadd
  CU(cudaMemsetAsync(g.envsum.p, 0, (size_t)4 * N * sizeof(double), ctx->stream));
  for (int i = 0; i < ndens; ++i) {
    Basis& b = *get_basis(ctx, basis[i]);
    Plan* pp = nullptr;
    TRY(get_plan(ctx, grid, basis[i], &pp));
    const size_t nb2 = (size_t)b.nbf * b.nbf;
    CU(ctx->dP.ensure(nb2 * sizeof(double) + P_SLACK_BYTES));
    CU(cudaMemcpyAsync(ctx->dP.p, P[i], nb2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // (dP is reused by the next subsystem; a diagnostic entry point)
    for (const Chunk& c : pp->chunks) {
      if (c.nslots == 0) continue;
      TRY(phase_basis(ctx, g, b, *pp, c));
      TRY(phase_density(ctx, g, b, *pp, c, ctx->dP.as<double>(), g.dens.as<double>(), gx != nullptr, nullptr));
      k_add4<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, gx ? 4 : 1, pp->block_id.as<int>() + c.slot0, g.envsum.as<double>(),
                                                g.dens.as<double>(), g.envsum.as<double>());
      LAUNCH_CHECK();
    }
  }
  const double* d = g.envsum.as<double>();
  CU(cudaMemcpyAsync(rho, d, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (gx) {
    CU(cudaMemcpyAsync(gx, d + N, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(gy, d + 2 * N, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(gz, d + 3 * N, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return SXC_OK;
}

// derivative level 2 of getBlockOnGridData (BasisFunctionOnGridController.cpp:302-304 radial second derivative, :381-440 /
// :1081-1095 finalisation): the six second derivatives of every basis function on one block, n x nbf column-major like
// sxc_basis_on_grid.  k_hessq contracts the Hessian with a vector; with the three unit vectors it returns the Hessian itself.
int sxc_basis_hessian_on_grid(sxc_ctx* ctx, int grid, int basis, int block, double* hxx, double* hxy, double* hxz, double* hyy,
                              double* hyz, double* hzz, int* n_out) {
  if (!ctx || !hxx || !hxy || !hxz || !hyy || !hyz || !hzz) return fail(ctx, SXC_ERR_INVALID, "sxc_basis_hessian_on_grid: bad arguments");
  CU(cudaSetDevice(ctx->device));
  Plan* pp = nullptr;
  TRY(get_plan(ctx, grid, basis, &pp, GRAD_TILE_COMPS));
  Plan& p = *pp;
  Grid& g = *get_grid(ctx, grid);
  Basis& b = *get_basis(ctx, basis);
  const int q = block - g.own_first;
  if (q < 0 || q >= p.nown) return fail(ctx, SXC_ERR_INVALID, "block %d is not owned by this context", block);
  const Chunk* ch = nullptr;
  for (const Chunk& c : p.chunks)
    if (q >= c.slot0 && q < c.slot0 + c.nslots) ch = &c;
  ctx->phi_owner = 0;
  CU(ctx->phi.ensure(ch->doubles * sizeof(double)));
  const int s = p.h_s[q], sp = p.h_s_pad[q];
  std::vector<long long> off(1);
  std::vector<int> sig(std::max(s, 1));
  CU(cudaMemcpyAsync(off.data(), p.phi_off.as<long long>() + q, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  if (s) CU(cudaMemcpyAsync(sig.data(), p.sig_bf.as<int>() + (size_t)q * p.nbf_pad, s * sizeof(int), cudaMemcpyDeviceToHost,
                            ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const long first = (long)block * g.blocksize;
  const int n = (int)std::min<long>(g.blocksize, g.npts - first);
  if (n_out) *n_out = n;
  // unit vector d -> slots 5, 6, 7 hold d_d d_x, d_d d_y, d_d d_z
  double* outs[3][3] = {{hxx, hxy, hxz}, {nullptr, hyy, hyz}, {nullptr, nullptr, hzz}};
  std::vector<double> tile((size_t)3 * sp * BP);
  for (int d = 0; d < 3; ++d) {
    k_hessq<<<1, BASIS_GROUPS * BP, 0, ctx->stream>>>(g.view(), b.view(), p.view(), q, nullptr, nullptr, nullptr, nullptr,
                                                      ctx->phi.as<double>(), d);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(tile.data(), ctx->phi.as<double>() + off[0] + (size_t)5 * sp * BP, tile.size() * sizeof(double),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int c = d; c < 3; ++c) {
      double* out = outs[d][c];
      std::fill(out, out + (size_t)n * b.nbf, 0.0);
      for (int k = 0; k < s; ++k) {
        const double* src = tile.data() + ((size_t)c * sp + k) * BP;
        std::copy(src, src + n, out + (size_t)sig[k] * n);
      }
    }
  }
  return SXC_OK;
}

// second derivatives of the density on the grid (MatrixOperatorToGridTransformer.cpp:166-188):
//   d_c d_d rho = 2 sum_mu,nu P_mu,nu ( phi_mu d_c d_d phi_nu + d_c phi_mu d_d phi_nu )
// as six runs of the density contraction on other tile slots of the 8-slot gradient plan: for every axis d, k_hessq (unit vector d)
// leaves d_d d_c phi in slots 5-7 and k_density(A = phi, epilogue slots 4-7) returns 2 sum (phi P) d_d d_c phi; k_density(A =
// d_d phi, epilogue slots 0-3) returns 2 sum (d_d phi P) d_c phi.  Host outputs [N] each.
int sxc_density_hessian_on_grid(sxc_ctx* ctx, int grid, int basis, const double* P, double* hxx, double* hxy, double* hxz,
                                double* hyy, double* hyz, double* hzz) {
  if (!ctx || !P || !hxx || !hxy || !hxz || !hyy || !hyz || !hzz)
    return fail(ctx, SXC_ERR_INVALID, "sxc_density_hessian_on_grid: bad arguments");
  CU(cudaSetDevice(ctx->device));
  Plan* pp = nullptr;
  TRY(get_plan(ctx, grid, basis, &pp, GRAD_TILE_COMPS));
  Plan& p = *pp;
  Grid& g = *get_grid(ctx, grid);
  Basis& b = *get_basis(ctx, basis);
  TRY(ensure_point_arrays(ctx, g, false, 2));  // two [4][N] outputs
  const size_t nb2 = (size_t)b.nbf * b.nbf;
  const long N = g.npts;
  CU(ctx->dP.ensure(nb2 * sizeof(double) + P_SLACK_BYTES));
  CU(cudaMemcpyAsync(ctx->dP.p, P, nb2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  double* t1 = g.dens.as<double>();          // [4][N]: (unused), 2 sum (phi P) d_d d_c phi for c = x, y, z
  double* t2 = t1 + (size_t)4 * N;           // [4][N]: (unused), 2 sum (d_d phi P) d_c phi
  double* outs[3][3] = {{hxx, hxy, hxz}, {nullptr, hyy, hyz}, {nullptr, nullptr, hzz}};
  std::vector<double> h1((size_t)3 * N), h2((size_t)3 * N);
  for (double* o : {hxx, hxy, hxz, hyy, hyz, hzz}) std::fill(o, o + N, 0.0);
  for (int d = 0; d < 3; ++d) {
    CU(cudaMemsetAsync(t1, 0, (size_t)8 * N * sizeof(double), ctx->stream));
    for (const Chunk& c : p.chunks) {
      if (c.nslots == 0) continue;
      TRY(phase_basis(ctx, g, b, p, c));
      k_hessq<<<c.nslots, BASIS_GROUPS * BP, 0, ctx->stream>>>(g.view(), b.view(), p.view(), c.slot0, p.order.as<int>() + c.order_off,
                                                               nullptr, nullptr, nullptr, ctx->phi.as<double>(), d);
      LAUNCH_CHECK();
      // (slot 4, the scatter's G, is not written on this path; it only feeds the first output row, which is not used)
      TRY(phase_density(ctx, g, b, p, c, ctx->dP.as<double>(), t1, true, nullptr, 0, 4));
      TRY(phase_density(ctx, g, b, p, c, ctx->dP.as<double>(), t2, true, nullptr, 1 + d, 0));
    }
    CU(cudaMemcpyAsync(h1.data(), t1 + N, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(h2.data(), t2 + N, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int c = d; c < 3; ++c)
      for (long i = 0; i < N; ++i) outs[d][c][i] = h1[(size_t)c * N + i] + h2[(size_t)c * N + i];
  }
  return SXC_OK;
}

int sxc_functional_on_grid(sxc_ctx* ctx, int func, int64_t npts, const double* w, const double* rho, const double* gx,
                           const double* gy, const double* gz, double* epuv, double* dFdRho, double* dFdGx, double* dFdGy,
                           double* dFdGz, double* energy) {
  if (!ctx || npts <= 0 || !w || !rho || !epuv || !dFdRho)
    return fail(ctx, SXC_ERR_INVALID, "sxc_functional_on_grid: bad arguments");
  if (func < 0 || func >= (int)ctx->funcs.size()) return fail(ctx, SXC_ERR_INVALID, "invalid functional handle %d", func);
  CU(cudaSetDevice(ctx->device));
  FuncView f = ctx->funcs[func];
  if (f.gga && !(gx && gy && gz)) return fail(ctx, SXC_ERR_INVALID, "GGA functional needs the density gradient");
  const bool want_g = f.gga && dFdGx && dFdGy && dFdGz;
  const int nlit = (int)((npts + FUNC_BLOCK - 1) / FUNC_BLOCK);
  DevMem buf;
  const size_t N = (size_t)npts;
  CU(buf.ensure((10 * N + nlit + 1) * sizeof(double)));
  double* d = buf.as<double>();
  double *d_w = d, *d_rho = d + N, *d_gx = d + 2 * N, *d_gy = d + 3 * N, *d_gz = d + 4 * N, *d_ep = d + 5 * N,
         *d_v = d + 6 * N, *d_part = d + 10 * N;
  CU(cudaMemcpyAsync(d_w, w, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_rho, rho, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (f.gga) {
    CU(cudaMemcpyAsync(d_gx, gx, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_gy, gy, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_gz, gz, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  k_functional<<<nlit, FUNC_BLOCK, 0, ctx->stream>>>(f, npts, nullptr, d_w, d_rho, d_gx, d_gy, d_gz, 1.0, 0, d_ep, d_v,
                                                     f.gga ? d_v + N : nullptr, f.gga ? d_v + 2 * N : nullptr,
                                                     f.gga ? d_v + 3 * N : nullptr, d_part, nullptr);
  LAUNCH_CHECK();
  k_reduce_partials<<<1, 256, 0, ctx->stream>>>(d_part, nlit, 1.0, 0, d_part + nlit);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(epuv, d_ep, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(dFdRho, d_v, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (want_g) {
    CU(cudaMemcpyAsync(dFdGx, d_v + N, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(dFdGy, d_v + 2 * N, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(dFdGz, d_v + 3 * N, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  double e = 0.0;
  CU(cudaMemcpyAsync(&e, d_part + nlit, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (energy) *energy = e;
  return SXC_OK;
}

int sxc_functional_on_grid_u(sxc_ctx* ctx, int func, int64_t npts, const double* w, const double* dens8, int has_grad,
                             double* epuv, double* out8, double* energy) {
  if (!ctx || npts <= 0 || !w || !dens8 || !epuv || !out8)
    return fail(ctx, SXC_ERR_INVALID, "sxc_functional_on_grid_u: bad arguments");
  if (func < 0 || func >= (int)ctx->funcs.size()) return fail(ctx, SXC_ERR_INVALID, "invalid functional handle %d", func);
  CU(cudaSetDevice(ctx->device));
  FuncView f = ctx->funcs[func];
  if (f.gga && !has_grad) return fail(ctx, SXC_ERR_INVALID, "GGA functional needs the density gradients");
  const int nlit = (int)((npts + FUNC_BLOCK - 1) / FUNC_BLOCK);
  DevMem buf;
  const size_t N = (size_t)npts;
  CU(buf.ensure((18 * N + nlit + 1) * sizeof(double)));
  double* d = buf.as<double>();
  double *d_w = d, *d_in = d + N, *d_ep = d + 9 * N, *d_out = d + 10 * N, *d_part = d + 18 * N;
  CU(cudaMemcpyAsync(d_w, w, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_in, dens8, 8 * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(d_out, 0, 8 * N * sizeof(double), ctx->stream));
  k_functional_u<<<nlit, FUNC_BLOCK, 0, ctx->stream>>>(f, npts, nullptr, d_w, d_in, 1.0, 0, d_ep, d_out, d_part, nullptr);
  LAUNCH_CHECK();
  k_reduce_partials<<<1, 256, 0, ctx->stream>>>(d_part, nlit, 1.0, 0, d_part + nlit);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(epuv, d_ep, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(out8, d_out, 8 * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  double e = 0.0;
  CU(cudaMemcpyAsync(&e, d_part + nlit, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (energy) *energy = e;
  return SXC_OK;
}

int sxc_scalar_to_matrix(sxc_ctx* ctx, int grid, int basis, double thr, const double* v, const double* gx,
                         const double* gy, const double* gz, double* V) {
  if (!ctx || !v || !V) return fail(ctx, SXC_ERR_INVALID, "sxc_scalar_to_matrix: bad arguments");
  CU(cudaSetDevice(ctx->device));
  Plan* pp = nullptr;
  TRY(get_plan(ctx, grid, basis, &pp));
  Grid& g = *get_grid(ctx, grid);
  Basis& b = *get_basis(ctx, basis);
  TRY(ensure_point_arrays(ctx, g, false, 1));
  const size_t nb2 = (size_t)b.nbf * b.nbf;
  const long N = g.npts;
  const bool gga = gx != nullptr;
  if (gga && !(gy && gz)) return fail(ctx, SXC_ERR_INVALID, "gradient operator needs all three components");
  CU(ctx->dOut.ensure((nb2 + 2) * sizeof(double)));
  double* pot = g.pot.as<double>();
  CU(cudaMemcpyAsync(pot, v, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (gga) {
    CU(cudaMemcpyAsync(pot + N, gx, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(pot + 2 * N, gy, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(pot + 3 * N, gz, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaMemsetAsync(ctx->dOut.p, 0, nb2 * sizeof(double), ctx->stream));
  for (const Chunk& c : pp->chunks) {
    TRY(phase_basis(ctx, g, b, *pp, c));
    TRY(phase_scatter(ctx, g, b, *pp, c, gga, thr, pot, ctx->dOut.as<double>()));
  }
  TRY(finish_matrix(ctx, b.nbf, ctx->dOut.as<double>()));
  std::vector<double> h(nb2);
  CU(cudaMemcpyAsync(h.data(), ctx->dOut.p, nb2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < nb2; ++i) V[i] += h[i];  // the reference adds into the caller's matrix
  return SXC_OK;
}

int sxc_scalar_to_matrix_ab(sxc_ctx* ctx, int grid, int basis_a, int basis_b, double thr, const double* v, const double* gx,
                            const double* gy, const double* gz, double* V) {
  if (!ctx || !v || !V) return fail(ctx, SXC_ERR_INVALID, "sxc_scalar_to_matrix_ab: bad arguments");
  CU(cudaSetDevice(ctx->device));
  Plan *pa = nullptr, *pb = nullptr;
  TRY(ab_plans(ctx, grid, basis_a, basis_b, &pa, &pb));
  Grid& g = *get_grid(ctx, grid);
  Basis& ba = *get_basis(ctx, basis_a);
  Basis& bb = *get_basis(ctx, basis_b);
  TRY(ensure_point_arrays(ctx, g, false, 1));
  const size_t nab = (size_t)ba.nbf * bb.nbf;
  const long N = g.npts;
  const bool gga = gx != nullptr;
  if (gga && !(gy && gz)) return fail(ctx, SXC_ERR_INVALID, "gradient operator needs all three components");
  CU(ctx->dOut.ensure((nab + 2) * sizeof(double)));
  double* pot = g.pot.as<double>();
  CU(cudaMemcpyAsync(pot, v, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (gga) {
    CU(cudaMemcpyAsync(pot + N, gx, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(pot + 2 * N, gy, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(pot + 3 * N, gz, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaMemsetAsync(ctx->dOut.p, 0, nab * sizeof(double), ctx->stream));
  if (!pa->chunks.empty()) {
    TRY(phase_basis(ctx, g, ba, *pa, pa->chunks[0]));
    TRY(phase_basis(ctx, g, bb, *pb, pb->chunks[0], &ctx->phi2));
    TRY(phase_scatter_ab(ctx, g, ba, *pa, *pb, gga, thr, pot, ctx->dOut.as<double>()));
  }
  std::vector<double> h(nab);
  CU(cudaMemcpyAsync(h.data(), ctx->dOut.p, nab * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < nab; ++i) V[i] += h[i];  // the reference adds into the caller's matrix
  return SXC_OK;
}

int sxc_build_ab(sxc_ctx* ctx, int grid, int func, int nspin, int basis_a, int basis_b, int ndens, const int* basis_c,
                 const double* const* P_c, double thr, double* V_ab, double* E) {
  if (!ctx || !V_ab || !E || ndens <= 0 || !basis_c || !P_c) return fail(ctx, SXC_ERR_INVALID, "sxc_build_ab: bad arguments");
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 or 2");
  Basis* ba = get_basis(ctx, basis_a);
  Basis* bb = get_basis(ctx, basis_b);
  if (!ba || !bb) return fail(ctx, SXC_ERR_INVALID, "invalid basis handle (%d, %d)", basis_a, basis_b);
  CU(cudaSetDevice(ctx->device));
  const size_t nab = (size_t)nspin * ba->nbf * bb->nbf;
  size_t total = 0;
  std::vector<size_t> offs(ndens);
  for (int i = 0; i < ndens; ++i) {
    Basis* bc = get_basis(ctx, basis_c[i]);
    if (!bc || !P_c[i]) return fail(ctx, SXC_ERR_INVALID, "invalid density basis handle %d", basis_c[i]);
    offs[i] = total;
    total += (size_t)nspin * bc->nbf * bc->nbf;
  }
  CU(ctx->dP.ensure(total * sizeof(double) + P_SLACK_BYTES));
  CU(ctx->dOut.ensure((nab + 2) * sizeof(double)));
  std::vector<const double*> dpc(ndens);
  HostCall host_guard{ctx};
  for (int i = 0; i < ndens; ++i) {
    Basis* bc = get_basis(ctx, basis_c[i]);
    dpc[i] = ctx->dP.as<double>() + offs[i];
    TRY(upload_async(ctx, ctx->dP.as<double>() + offs[i], P_c[i], (size_t)nspin * bc->nbf * bc->nbf * sizeof(double)));
  }
  TRY(upload_done(ctx));
  int rc = build_ab_device(ctx, grid, func, nspin, basis_a, basis_b, ndens, basis_c, dpc.data(), thr, ctx->dOut.as<double>());
  ctx->timing = 0;
  if (rc != SXC_OK) return abort_build(ctx, rc);
  CU(cudaMemcpyAsync(V_ab, ctx->dOut.p, nab * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(E, ctx->dOut.as<double>() + nab, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  collect_timers(ctx);
  return SXC_OK;
}

int sxc_build_ab_nadd(sxc_ctx* ctx, int grid, int func, int nspin, int basis_a, int basis_b, int basis_act, const double* P_act,
                      int nenv, const int* basis_env, const double* const* P_env, double thr, double* V_ab) {
  if (!ctx || !V_ab || !P_act || nenv < 0 || (nenv > 0 && (!basis_env || !P_env)))
    return fail(ctx, SXC_ERR_INVALID, "sxc_build_ab_nadd: bad arguments");
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 or 2");
  Basis* ba = get_basis(ctx, basis_a);
  Basis* bb = get_basis(ctx, basis_b);
  Basis* bact = get_basis(ctx, basis_act);
  if (!ba || !bb || !bact) return fail(ctx, SXC_ERR_INVALID, "invalid basis handle (%d, %d, %d)", basis_a, basis_b, basis_act);
  CU(cudaSetDevice(ctx->device));
  const size_t nab = (size_t)nspin * ba->nbf * bb->nbf;
  size_t total = (size_t)nspin * bact->nbf * bact->nbf;
  std::vector<size_t> offs(nenv);
  for (int i = 0; i < nenv; ++i) {
    Basis* be = get_basis(ctx, basis_env[i]);
    if (!be || !P_env[i]) return fail(ctx, SXC_ERR_INVALID, "invalid environment basis handle %d", basis_env[i]);
    offs[i] = total;
    total += (size_t)nspin * be->nbf * be->nbf;
  }
  CU(ctx->dP.ensure(total * sizeof(double) + P_SLACK_BYTES));
  CU(ctx->dOut.ensure((nab + 2) * sizeof(double)));
  HostCall host_guard{ctx};
  TRY(upload_async(ctx, ctx->dP.p, P_act, (size_t)nspin * bact->nbf * bact->nbf * sizeof(double)));
  std::vector<const double*> dpe(nenv);
  for (int i = 0; i < nenv; ++i) {
    Basis* be = get_basis(ctx, basis_env[i]);
    dpe[i] = ctx->dP.as<double>() + offs[i];
    TRY(upload_async(ctx, ctx->dP.as<double>() + offs[i], P_env[i], (size_t)nspin * be->nbf * be->nbf * sizeof(double)));
  }
  TRY(upload_done(ctx));
  int rc = build_ab_nadd_device(ctx, grid, func, nspin, basis_a, basis_b, basis_act, ctx->dP.as<double>(), nenv, basis_env,
                                dpe.data(), thr, ctx->dOut.as<double>());
  ctx->timing = 0;
  if (rc != SXC_OK) return abort_build(ctx, rc);
  CU(cudaMemcpyAsync(V_ab, ctx->dOut.p, nab * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  collect_timers(ctx);
  return SXC_OK;
}

int sxc_set_p_ready_event(sxc_ctx* ctx, void* cuda_event) {
  if (!ctx) return SXC_ERR_INVALID;
  ctx->p_ready = static_cast<cudaEvent_t>(cuda_event);
  return SXC_OK;
}

int sxc_set_tile_cache(sxc_ctx* ctx, int on) {
  if (!ctx) return SXC_ERR_INVALID;
  ctx->tile_cache = on != 0;
  if (!on) ctx->phi_owner = 0;
  return SXC_OK;
}

int sxc_set_output_slice(sxc_ctx* ctx, int part, int parts) {
  if (!ctx || parts < 1 || part < 0 || part >= parts) return SXC_ERR_INVALID;
  ctx->out_part = part;
  ctx->out_parts = parts;
  return SXC_OK;
}

int sxc_set_timing(sxc_ctx* ctx, int on) {
  if (!ctx) return SXC_ERR_INVALID;
  ctx->timing_device = on != 0;
  return SXC_OK;
}

int sxc_xc_gradient(sxc_ctx* ctx, int grid, int basis, int func, int nspin, const double* P, int natoms,
                    const int* atom_of_bf, double* grad) {
  if (!ctx || !P || !grad || !atom_of_bf || natoms <= 0) return fail(ctx, SXC_ERR_INVALID, "sxc_xc_gradient: bad arguments");
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 or 2");
  Basis* b = get_basis(ctx, basis);
  if (!b) return fail(ctx, SXC_ERR_INVALID, "invalid basis handle %d", basis);
  for (int i = 0; i < b->nbf; ++i)
    if (atom_of_bf[i] < 0 || atom_of_bf[i] >= natoms) return fail(ctx, SXC_ERR_INVALID, "atom_of_bf[%d] out of range", i);
  CU(cudaSetDevice(ctx->device));
  const size_t nv = (size_t)nspin * b->nbf * b->nbf;
  CU(ctx->dP.ensure(nv * sizeof(double) + P_SLACK_BYTES));
  CU(ctx->dOut.ensure((size_t)b->nbf * 3 * sizeof(double)));
  HostCall host_guard{ctx};
  TRY(upload_async(ctx, ctx->dP.p, P, nv * sizeof(double)));
  TRY(upload_done(ctx));
  TRY(build_gradient_device(ctx, grid, basis, func, nspin, ctx->dP.as<double>(), ctx->dOut.as<double>()));
  std::vector<double> t((size_t)b->nbf * 3);
  CU(cudaMemcpyAsync(t.data(), ctx->dOut.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  collect_timers(ctx);
  std::fill(grad, grad + (size_t)natoms * 3, 0.0);
  for (int nu = 0; nu < b->nbf; ++nu)  // g[atom(nu), c] = -2 t[nu, c], Eigen nAtoms x 3 column-major
    for (int c = 0; c < 3; ++c) grad[atom_of_bf[nu] + (size_t)c * natoms] -= 2.0 * t[(size_t)nu * 3 + c];
  return SXC_OK;
}

int sxc_partition_weights(sxc_ctx* ctx, int flavour, int becke_smoothing, int natoms, const double* coords,
                          const double* aij, int64_t npts, const double* xyz, const int* parent, double* w) {
  if (!ctx || natoms <= 0 || npts < 0 || !coords || (npts > 0 && (!xyz || !parent || !w)))
    return fail(ctx, SXC_ERR_INVALID, "sxc_partition_weights: bad arguments");
  if (flavour < 0 || flavour > 2)
    return fail(ctx, SXC_ERR_UNSUPPORTED, "grid flavour %d (0 = BECKE, 1 = SSF, 2 = VORONOI)", flavour);
  if ((size_t)PW_WARPS * natoms * sizeof(double) > 200 * 1024)
    return fail(ctx, SXC_ERR_UNSUPPORTED, "more than %d atoms", (int)(200 * 1024 / (PW_WARPS * sizeof(double))));
  if (npts == 0) return SXC_OK;
  for (int64_t p = 0; p < npts; ++p)
    if (parent[p] < 0 || parent[p] >= natoms) return fail(ctx, SXC_ERR_INVALID, "parent[%lld] out of range", (long long)p);
  CU(cudaSetDevice(ctx->device));
  // atom-atom distances and the nearest neighbour of every atom (GridFactory.cpp:76-92, :152-159)
  std::vector<double> adist((size_t)natoms * natoms), mind(natoms, 999999999.9);
  for (int i = 0; i < natoms; ++i)
    for (int j = 0; j < natoms; ++j) {
      const double dx = coords[3 * i] - coords[3 * j], dy = coords[3 * i + 1] - coords[3 * j + 1],
                   dz = coords[3 * i + 2] - coords[3 * j + 2];
      const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
      adist[i + (size_t)natoms * j] = d;
      if (i != j) mind[j] = std::min(mind[j], d);
    }
  DevMem dc, da, daij, dm, dx, dp, dw;
  const size_t n2 = (size_t)natoms * natoms * sizeof(double);
  CU(dc.ensure((size_t)3 * natoms * sizeof(double)));
  CU(da.ensure(n2));
  CU(dm.ensure(natoms * sizeof(double)));
  CU(dx.ensure((size_t)3 * npts * sizeof(double)));
  CU(dp.ensure((size_t)npts * sizeof(int)));
  CU(dw.ensure((size_t)npts * sizeof(double)));
  if (aij) {
    CU(daij.ensure(n2));
    CU(cudaMemcpyAsync(daij.p, aij, n2, cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaMemcpyAsync(dc.p, coords, (size_t)3 * natoms * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(da.p, adist.data(), n2, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dm.p, mind.data(), natoms * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dx.p, xyz, (size_t)3 * npts * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dp.p, parent, (size_t)npts * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dw.p, w, (size_t)npts * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const size_t smem = (size_t)PW_WARPS * natoms * sizeof(double);
  if (smem > 48 * 1024)
    CU(cudaFuncSetAttribute(k_partition_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t want = (npts + PW_WARPS - 1) / PW_WARPS;
  const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->num_sms * 8);  // grid-stride over the points
  struct PooledEvents {  // returned to the context's pool on every exit path
    sxc_ctx* c;
    cudaEvent_t a, b;
    ~PooledEvents() {
      c->event_pool.push_back(a);
      c->event_pool.push_back(b);
    }
  } ev{ctx, take_event(ctx), take_event(ctx)};
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  CU(cudaEventRecord(e0, ctx->stream));
  k_partition_weights<<<grid, PW_WARPS * 32, smem, ctx->stream>>>(flavour, std::max(1, becke_smoothing), natoms, dc.as<double>(), da.as<double>(),
                                                                   aij ? daij.as<double>() : nullptr, dm.as<double>(),
                                                                   (long)npts, dx.as<double>(), dp.as<int>(), dw.as<double>());
  CU(cudaEventRecord(e1, ctx->stream));
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(w, dw.p, (size_t)npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx->last_partition_ms = ms;
  ctx->stats.kernel_launches += 1;
  return SXC_OK;
}

double sxc_last_partition_ms(sxc_ctx* ctx) { return ctx ? (double)ctx->last_partition_ms : -1.0; }

int sxc_nadd_gradient(sxc_ctx* ctx, int grid, int func, int nspin, int basis_act, const double* P_act, int nenv,
                      const int* basis_env, const double* const* P_env, int natoms, const int* atom_of_bf, double* grad) {
  if (!ctx || !P_act || !grad || !atom_of_bf || natoms <= 0 || nenv < 0 || (nenv > 0 && (!basis_env || !P_env)))
    return fail(ctx, SXC_ERR_INVALID, "sxc_nadd_gradient: bad arguments");
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 or 2");
  Basis* b = get_basis(ctx, basis_act);
  if (!b) return fail(ctx, SXC_ERR_INVALID, "invalid active basis handle %d", basis_act);
  for (int i = 0; i < b->nbf; ++i)
    if (atom_of_bf[i] < 0 || atom_of_bf[i] >= natoms) return fail(ctx, SXC_ERR_INVALID, "atom_of_bf[%d] out of range", i);
  CU(cudaSetDevice(ctx->device));
  const size_t nvA = (size_t)nspin * b->nbf * b->nbf;
  size_t total = nvA;
  std::vector<size_t> offs(nenv);
  for (int i = 0; i < nenv; ++i) {
    Basis* be = get_basis(ctx, basis_env[i]);
    if (!be || !P_env[i]) return fail(ctx, SXC_ERR_INVALID, "invalid environment basis handle %d", basis_env[i]);
    offs[i] = total;
    total += (size_t)nspin * be->nbf * be->nbf;
  }
  CU(ctx->dP.ensure(total * sizeof(double) + P_SLACK_BYTES));
  CU(ctx->dOut.ensure((size_t)b->nbf * 3 * sizeof(double)));
  HostCall host_guard{ctx};
  TRY(upload_async(ctx, ctx->dP.p, P_act, nvA * sizeof(double)));
  std::vector<const double*> dpe(nenv);
  for (int i = 0; i < nenv; ++i) {
    Basis* be = get_basis(ctx, basis_env[i]);
    dpe[i] = ctx->dP.as<double>() + offs[i];
    TRY(upload_async(ctx, ctx->dP.as<double>() + offs[i], P_env[i], (size_t)nspin * be->nbf * be->nbf * sizeof(double)));
  }
  TRY(upload_done(ctx));
  TRY(build_gradient_device(ctx, grid, basis_act, func, nspin, ctx->dP.as<double>(), ctx->dOut.as<double>(), nenv, basis_env,
                            dpe.data()));
  std::vector<double> t((size_t)b->nbf * 3);
  CU(cudaMemcpyAsync(t.data(), ctx->dOut.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  collect_timers(ctx);
  std::fill(grad, grad + (size_t)natoms * 3, 0.0);
  for (int nu = 0; nu < b->nbf; ++nu)
    for (int c = 0; c < 3; ++c) grad[atom_of_bf[nu] + (size_t)c * natoms] -= 2.0 * t[(size_t)nu * 3 + c];
  return SXC_OK;
}

// ---- row f-4: LR-TDDFT / subsystem-TDDFT kernel ------------------------------------------------------------------
int sxc_kernel_create(sxc_ctx* ctx, int grid, int nspin, int gga, int* kernel) {
  if (!ctx || !kernel) return fail(ctx, SXC_ERR_INVALID, "sxc_kernel_create: bad arguments");
  if (nspin != 1 && nspin != 2) return fail(ctx, SXC_ERR_INVALID, "nspin must be 1 or 2");
  Grid* g = get_grid(ctx, grid);
  if (!g) return fail(ctx, SXC_ERR_INVALID, "invalid grid handle %d", grid);
  CU(cudaSetDevice(ctx->device));
  auto ks = std::make_unique<KernelStore>();
  ks->grid = grid;
  ks->nspin = nspin;
  ks->gga = gga != 0;
  ks->narr = nspin == 1 ? (ks->gga ? KR_ARRAYS : 1) : (ks->gga ? KU_ARRAYS : 3);
  const size_t bytes = (size_t)ks->narr * std::max<long>(g->npts, 1) * sizeof(double);
  CU(ks->data.ensure(bytes));
  CU(cudaMemsetAsync(ks->data.p, 0, bytes, ctx->stream));
  // reuse a released slot
  for (size_t i = 0; i < ctx->kstores.size(); ++i)
    if (!ctx->kstores[i]) {
      ctx->kstores[i] = std::move(ks);
      *kernel = (int)i;
      return SXC_OK;
    }
  ctx->kstores.push_back(std::move(ks));
  *kernel = (int)ctx->kstores.size() - 1;
  return SXC_OK;
}

int sxc_kernel_destroy(sxc_ctx* ctx, int kernel) {
  if (!ctx || kernel < 0 || kernel >= (int)ctx->kstores.size() || !ctx->kstores[kernel])
    return fail(ctx, SXC_ERR_INVALID, "invalid kernel handle %d", kernel);
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->kstores[kernel].reset();
  return SXC_OK;
}

int sxc_kernel_add(sxc_ctx* ctx, int kernel, int func, double sign, int ndens, const int* basis_c, const double* const* P_c) {
  if (!ctx || ndens <= 0 || !basis_c || !P_c) return fail(ctx, SXC_ERR_INVALID, "sxc_kernel_add: bad arguments");
  if (kernel < 0 || kernel >= (int)ctx->kstores.size() || !ctx->kstores[kernel])
    return fail(ctx, SXC_ERR_INVALID, "invalid kernel handle %d", kernel);
  if (func < 0 || func >= (int)ctx->funcs.size()) return fail(ctx, SXC_ERR_INVALID, "invalid functional handle %d", func);
  CU(cudaSetDevice(ctx->device));
  KernelStore& ks = *ctx->kstores[kernel];
  const int gh = ks.grid, nspin = ks.nspin;
  Grid& g = *get_grid(ctx, gh);
  const FuncView f = ctx->funcs[func];
  if (f.ncomp == 0) return SXC_OK;  // CompositeFunctionals::CLASSES::NONE adds nothing (Kernel.cpp:716, :722)
  if (f.gga && !ks.gga) return fail(ctx, SXC_ERR_INVALID, "a GGA functional needs a kernel store created with gga = 1");
  size_t total = 0;
  std::vector<size_t> offs(ndens);
  for (int i = 0; i < ndens; ++i) {
    Basis* bc = get_basis(ctx, basis_c[i]);
    if (!bc || !P_c[i]) return fail(ctx, SXC_ERR_INVALID, "invalid density basis handle %d", basis_c[i]);
    Plan* pc = nullptr;
    TRY(get_plan(ctx, gh, basis_c[i], &pc));
    offs[i] = total;
    total += (size_t)nspin * bc->nbf * bc->nbf;
  }
  CU(ctx->dP.ensure(total * sizeof(double) + P_SLACK_BYTES));
  HostCall host_guard{ctx};
  for (int i = 0; i < ndens; ++i) {
    Basis* bc = get_basis(ctx, basis_c[i]);
    TRY(upload_async(ctx, ctx->dP.as<double>() + offs[i], P_c[i], (size_t)nspin * bc->nbf * bc->nbf * sizeof(double)));
  }
  TRY(upload_done(ctx));
  TRY(ensure_point_arrays(ctx, g, true, nspin));
  const long N = g.npts;
  const int ncomp = 4 * nspin;
  double* dens = g.dens.as<double>();
  double* tot = g.tot.as<double>();
  begin_timing(ctx, timing_mode(ctx, true));
  const int launches0 = ctx->launches;
  Plan* p0 = nullptr;
  TRY(get_plan(ctx, gh, basis_c[0], &p0));
  ctx->stats = p0->stats;
  {
    PhaseTimer t_all(ctx, T_TOTAL);
    CU(cudaMemsetAsync(tot, 0, (size_t)ncomp * N * sizeof(double), ctx->stream));
    TRY(wait_p_ready(ctx));
    for (int i = 0; i < ndens; ++i) {  // total density of the listed systems (Kernel.cpp:693-712)
      Basis& bc = *get_basis(ctx, basis_c[i]);
      const size_t nc2 = (size_t)bc.nbf * bc.nbf;
      Plan* pc = nullptr;
      TRY(get_plan(ctx, gh, basis_c[i], &pc));
      for (const Chunk& c : pc->chunks) {
        TRY(phase_basis(ctx, g, bc, *pc, c));
        for (int sp = 0; sp < nspin; ++sp)
          TRY(phase_density(ctx, g, bc, *pc, c, ctx->dP.as<double>() + offs[i] + sp * nc2, dens + (size_t)4 * sp * N, true,
                            nullptr));
        if (pc->nown) {
          PhaseTimer t(ctx, SXC_T_DENSITY);
          k_add4<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, ncomp, pc->block_id.as<int>() + c.slot0, tot, dens, tot);
          LAUNCH_CHECK();
        }
      }
    }
    // storeDerivatives on the literal blocks this context owns
    const bool lit_is_block = g.blocksize == FUNC_BLOCK;
    for (const Chunk& c : p0->chunks) {
      const int nb = lit_is_block ? c.nslots : g.nlit;
      if (nb == 0) continue;
      const int* list = lit_is_block ? p0->block_id.as<int>() + c.slot0 : nullptr;
      PhaseTimer t(ctx, SXC_T_FUNCTIONAL);
      if (nspin == 2)
        k_kernel2_u<<<nb, FUNC_BLOCK, 0, ctx->stream>>>(f, N, list, tot, sign, ks.gga, ks.data.as<double>());
      else
        k_kernel2_r<<<nb, FUNC_BLOCK, 0, ctx->stream>>>(f, N, list, tot, sign, ks.gga, ks.data.as<double>());
      LAUNCH_CHECK();
      if (!lit_is_block) break;
    }
  }
  ctx->timing = 0;
  ctx->stats.kernel_launches = ctx->launches - launches0;
  CU(cudaStreamSynchronize(ctx->stream));
  collect_timers(ctx);
  return SXC_OK;
}

int sxc_kernel_get(sxc_ctx* ctx, int kernel, double* out) {
  if (!ctx || !out || kernel < 0 || kernel >= (int)ctx->kstores.size() || !ctx->kstores[kernel])
    return fail(ctx, SXC_ERR_INVALID, "invalid kernel handle %d", kernel);
  CU(cudaSetDevice(ctx->device));
  KernelStore& ks = *ctx->kstores[kernel];
  Grid& g = *get_grid(ctx, ks.grid);
  CU(cudaMemcpyAsync(out, ks.data.p, (size_t)ks.narr * g.npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SXC_OK;
}

int sxc_kernel_num_arrays(sxc_ctx* ctx, int kernel) {
  if (!ctx || kernel < 0 || kernel >= (int)ctx->kstores.size() || !ctx->kstores[kernel])
    return fail(ctx, SXC_ERR_INVALID, "invalid kernel handle %d", kernel);
  return ctx->kstores[kernel]->narr;
}

namespace {
__global__ void k_symmetrise(int n, double* __restrict__ D) {  // D += D^T (KernelSigmavector.cpp:201-208)
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < n && j <= i) {
    const double s = D[i + (size_t)j * n] + D[j + (size_t)i * n];
    D[i + (size_t)j * n] = s;
    D[j + (size_t)i * n] = s;
  }
}
}  // namespace

namespace {
// D_host != nullptr: uploaded on the side stream; else D_dev (device) is copied into the staging buffer on the build stream
int kernel_contract_impl(sxc_ctx* ctx, int grid, int basis_j, int nkern, const int* kernels, int mode, int nvec,
                         const double* D_host, const double* D_dev, int accumulate, bool sync) {
  const double* D = D_host ? D_host : D_dev;
  if (!ctx || !kernels || !D || nvec <= 0 || nkern < 1 || nkern > 3)
    return fail(ctx, SXC_ERR_INVALID, "sxc_kernel_contract: bad arguments (1 to 3 kernel stores, nvec > 0)");
  if (mode < 0 || mode > 2) return fail(ctx, SXC_ERR_INVALID, "mode must be 0 (singlet), 1 (triplet) or 2 (UNRESTRICTED)");
  Grid* gp = get_grid(ctx, grid);
  Basis* bp = get_basis(ctx, basis_j);
  if (!gp || !bp) return fail(ctx, SXC_ERR_INVALID, "invalid grid (%d) or basis (%d) handle", grid, basis_j);
  const int store_nspin = mode == 0 ? 1 : 2, nspin = mode == 2 ? 2 : 1;
  const double* st[3] = {nullptr, nullptr, nullptr};
  int gga = -1;
  for (int k = 0; k < nkern; ++k) {
    const int h = kernels[k];
    if (h < 0 || h >= (int)ctx->kstores.size() || !ctx->kstores[h]) return fail(ctx, SXC_ERR_INVALID, "invalid kernel handle %d", h);
    KernelStore& ks = *ctx->kstores[h];
    if (ks.grid != grid || ks.nspin != store_nspin || (gga >= 0 && ks.gga != gga))
      return fail(ctx, SXC_ERR_INVALID, "kernel store %d does not match (grid, spin mode, gga) of this contraction", h);
    gga = ks.gga;
    st[k] = ks.data.as<double>();
  }
  CU(cudaSetDevice(ctx->device));
  Grid& g = *gp;
  Basis& b = *bp;
  Plan* pp = nullptr;
  TRY(get_plan(ctx, grid, basis_j, &pp));
  Plan& p = *pp;
  TRY(ensure_point_arrays(ctx, g, false, nspin));
  const long N = g.npts;
  const size_t nb2 = (size_t)b.nbf * b.nbf;
  const size_t rows = (size_t)4 * nspin;
  const size_t resp_bytes = (size_t)nvec * rows * std::max<long>(N, 1) * sizeof(double);
  const bool fresh = !accumulate || g.resp_nvec != nvec || g.resp_nspin != nspin || g.resp_gga != gga;
  if (accumulate && fresh && g.resp_nvec != 0)
    return fail(ctx, SXC_ERR_INVALID, "accumulate: the stored response (nvec %d, nspin %d) does not match", g.resp_nvec, g.resp_nspin);
  if (fresh) {
    CU(g.resp.ensure(resp_bytes));
    CU(cudaMemsetAsync(g.resp.p, 0, resp_bytes, ctx->stream));
    g.resp_nvec = nvec;
    g.resp_nspin = nspin;
    g.resp_gga = gga;
  }
  DevMem& stage = D_host ? ctx->dP : ctx->dD;
  CU(stage.ensure((size_t)nvec * nspin * nb2 * sizeof(double)));
  double* dDs = stage.as<double>();
  HostCall host_guard{ctx};
  if (D_host) {
    TRY(upload_async(ctx, dDs, D_host, (size_t)nvec * nspin * nb2 * sizeof(double)));
    TRY(upload_done(ctx));
  } else {  // the caller's matrices are left untouched: D += D^T works on the staged copy
    CU(cudaMemcpyAsync(dDs, D_dev, (size_t)nvec * nspin * nb2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  begin_timing(ctx, timing_mode(ctx, sync));
  ctx->stats = p.stats;
  const int launches0 = ctx->launches;
  {
    PhaseTimer t_all(ctx, T_TOTAL);
    TRY(wait_p_ready(ctx));
    {
      dim3 blk(32, 8), grd((b.nbf + 31) / 32, (b.nbf + 7) / 8);
      for (int m = 0; m < nvec * nspin; ++m) {
        k_symmetrise<<<grd, blk, 0, ctx->stream>>>(b.nbf, dDs + (size_t)m * nb2);
        LAUNCH_CHECK();
      }
    }
    double* dens = g.dens.as<double>();
    for (const Chunk& c : p.chunks) {
      if (c.nslots == 0) continue;
      TRY(phase_basis(ctx, g, b, p, c));
      for (int v = 0; v < nvec; ++v) {
        for (int sp = 0; sp < nspin; ++sp)
          TRY(phase_density(ctx, g, b, p, c, dDs + ((size_t)v * nspin + sp) * nb2, dens + (size_t)4 * sp * N,
                            gga != 0, nullptr));
        PhaseTimer t(ctx, SXC_T_FUNCTIONAL);
        k_kernel_apply<<<c.nslots, 128, 0, ctx->stream>>>(N, g.blocksize, p.block_id.as<int>() + c.slot0, mode, gga, st[0], st[1],
                                                          st[2], dens, 1, g.resp.as<double>() + (size_t)v * rows * N);
        LAUNCH_CHECK();
      }
    }
  }
  ctx->timing = 0;
  ctx->stats.kernel_launches = ctx->launches - launches0;
  if (sync) {
    CU(cudaStreamSynchronize(ctx->stream));
    collect_timers(ctx);
  }
  return SXC_OK;
}
}  // namespace

int sxc_kernel_contract(sxc_ctx* ctx, int grid, int basis_j, int nkern, const int* kernels, int mode, int nvec, const double* D,
                        int accumulate) {
  return kernel_contract_impl(ctx, grid, basis_j, nkern, kernels, mode, nvec, D, nullptr, accumulate, true);
}
int sxc_kernel_contract_device(sxc_ctx* ctx, int grid, int basis_j, int nkern, const int* kernels, int mode, int nvec,
                               const double* d_D, int accumulate) {
  return kernel_contract_impl(ctx, grid, basis_j, nkern, kernels, mode, nvec, nullptr, d_D, accumulate, false);
}

int sxc_kernel_response_copy(sxc_ctx* ctx, int grid, int save) {
  Grid* gp = ctx ? get_grid(ctx, grid) : nullptr;
  if (!gp) return fail(ctx, SXC_ERR_INVALID, "invalid grid handle %d", grid);
  Grid& g = *gp;
  CU(cudaSetDevice(ctx->device));
  if (save) {
    if (g.resp_nvec == 0) return fail(ctx, SXC_ERR_INVALID, "no contracted response on grid %d to save", grid);
    const size_t bytes = (size_t)g.resp_nvec * 4 * g.resp_nspin * std::max<long>(g.npts, 1) * sizeof(double);
    CU(g.resp_saved.ensure(bytes));
    CU(cudaMemcpyAsync(g.resp_saved.p, g.resp.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    g.saved_nvec = g.resp_nvec;
    g.saved_nspin = g.resp_nspin;
    g.saved_gga = g.resp_gga;
  } else {
    if (g.saved_nvec == 0) return fail(ctx, SXC_ERR_INVALID, "no saved response on grid %d", grid);
    const size_t bytes = (size_t)g.saved_nvec * 4 * g.saved_nspin * std::max<long>(g.npts, 1) * sizeof(double);
    CU(g.resp.ensure(bytes));
    CU(cudaMemcpyAsync(g.resp.p, g.resp_saved.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    g.resp_nvec = g.saved_nvec;
    g.resp_nspin = g.saved_nspin;
    g.resp_gga = g.saved_gga;
  }
  return SXC_OK;
}

namespace {
// F_host: result staged in ctx->dOut and copied back; else written to d_F (device), asynchronously
int kernel_integrate_impl(sxc_ctx* ctx, int grid, int basis_i, double* F, double* d_F) {
  if (!ctx || (!F && !d_F)) return fail(ctx, SXC_ERR_INVALID, "sxc_kernel_integrate: bad arguments");
  Grid* gp = get_grid(ctx, grid);
  Basis* bp = get_basis(ctx, basis_i);
  if (!gp || !bp) return fail(ctx, SXC_ERR_INVALID, "invalid grid (%d) or basis (%d) handle", grid, basis_i);
  Grid& g = *gp;
  Basis& b = *bp;
  if (g.resp_nvec == 0) return fail(ctx, SXC_ERR_INVALID, "no contracted response on grid %d (call sxc_kernel_contract first)", grid);
  CU(cudaSetDevice(ctx->device));
  Plan* pp = nullptr;
  TRY(get_plan(ctx, grid, basis_i, &pp));
  Plan& p = *pp;
  const int nvec = g.resp_nvec, nspin = g.resp_nspin;
  const bool gga = g.resp_gga != 0;
  const long N = g.npts;
  const size_t nb2 = (size_t)b.nbf * b.nbf, nmat = (size_t)nvec * nspin;
  const size_t rows = (size_t)4 * nspin;
  double* dF = d_F;
  if (!dF) {
    CU(ctx->dOut.ensure((nmat * nb2 + 2) * sizeof(double)));
    dF = ctx->dOut.as<double>();
  }
  begin_timing(ctx, timing_mode(ctx, F != nullptr));
  ctx->stats = p.stats;
  const int launches0 = ctx->launches;
  {
    PhaseTimer t_all(ctx, T_TOTAL);
    CU(cudaMemsetAsync(dF, 0, nmat * nb2 * sizeof(double), ctx->stream));
    for (const Chunk& c : p.chunks) {
      if (c.nslots == 0) continue;
      TRY(phase_basis(ctx, g, b, p, c));
      // F + F^T = sum_p scal phi_i phi_j + grad . (phi_i grad phi_j + grad phi_i phi_j): the XC scatter with the
      // weights already inside scal / grad ... which the contraction left out, so the scatter's w restores them
      for (size_t m = 0; m < nmat; ++m)
        TRY(phase_scatter(ctx, g, b, p, c, gga, 0.0, g.resp.as<double>() + (m / nspin) * rows * N + (m % nspin) * 4 * N,
                          dF + m * nb2));
    }
    for (size_t m = 0; m < nmat; ++m) TRY(finish_matrix(ctx, b.nbf, dF + m * nb2));
    TRY(allreduce_result(ctx, g, dF, nmat * nb2));
  }
  ctx->timing = 0;
  ctx->stats.kernel_launches = ctx->launches - launches0;
  if (F) {
    CU(cudaMemcpyAsync(F, dF, nmat * nb2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    collect_timers(ctx);
  }
  return SXC_OK;
}
}  // namespace

int sxc_kernel_integrate(sxc_ctx* ctx, int grid, int basis_i, double* F) {
  if (!F) return fail(ctx, SXC_ERR_INVALID, "sxc_kernel_integrate: bad arguments");
  return kernel_integrate_impl(ctx, grid, basis_i, F, nullptr);
}
int sxc_kernel_integrate_device(sxc_ctx* ctx, int grid, int basis_i, double* d_F) {
  if (!d_F) return fail(ctx, SXC_ERR_INVALID, "sxc_kernel_integrate_device: bad arguments");
  return kernel_integrate_impl(ctx, grid, basis_i, nullptr, d_F);
}

int sxc_kernel_sigma(sxc_ctx* ctx, int grid, int basis, int nkern, const int* kernels, int mode, int nvec, const double* D,
                     double* F) {
  TRY(sxc_kernel_contract(ctx, grid, basis, nkern, kernels, mode, nvec, D, 0));
  return sxc_kernel_integrate(ctx, grid, basis, F);
}

int sxc_get_stats(sxc_ctx* ctx, sxc_stats* out) {
  if (!ctx || !out) return SXC_ERR_INVALID;
  collect_timers(ctx);
  *out = ctx->stats;
  return SXC_OK;
}

}  // extern "C"
