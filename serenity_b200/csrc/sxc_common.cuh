// sxc_common.cuh - shared device helpers (sm_100a): FP64 tensor-core MMA, cp.async, reductions.
#pragma once

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

namespace sxc {

constexpr int BP = 128;       // points per tile (blocks are padded to this; reference default blocksize)
constexpr int SPAD = 32;      // significant-function count is padded to a multiple of this
constexpr int LMAX = 6;       // AM_MAX, src/parameters/Constants.h:31
constexpr int TILE_COMPS = 5;  // tile = [phi, d/dx, d/dy, d/dz, G][s_pad][128]; G is written by the scatter phase
constexpr int FUNC_BLOCK = 128;  // the literal block size of the functional evaluation (FuncPotential.cpp:85)

// D(8x8) += A(8x4, row) * B(4x8, col) in FP64 on the tensor cores (SASS: DMMA.8x8x4).
// lane holds A[lane/4][lane%4], B[lane%4][lane/4], D[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// fire-and-forget FP64 accumulation into global memory (SASS: REDG.E.ADD.F64; written as PTX so that no code path turns it
// into an ATOMG that waits for the old value)
__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(addr), "d"(v) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void prefetch_l2(const void* gmem) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(gmem)); }
// `bytes` (multiple of 16) of contiguous global memory into L2, no destination, no completion
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- mbarrier (shared::cta) helpers: full / empty barriers of the producer-warp pipeline (k_vmat) ------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
// arrive-on that fires once all cp.async issued so far by this thread have landed (does not add to the pending count)
__device__ __forceinline__ void mbar_arrive_cp_async(uint64_t* bar) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile(
      "{\n .reg .pred p;\n"
      "WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n"
      " bra WAIT_%=;\n"
      "DONE_%=:\n}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}

// ---- TMA (cp.async.bulk.tensor) helpers: the producers of k_density_tma / k_vmat_tma ---------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// one box of a 2-D tensor map -> shared memory; completion is signalled to `bar` as transferred bytes
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(s),
      "l"(map), "r"(b), "r"(c0), "r"(c1)
      : "memory");
}
// the same box into L2 only (no shared memory, no completion): hides the HBM latency of a later tma_load_2d of the box
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (all threads must call); result valid in all threads
__device__ __forceinline__ double block_sum(double v, double* scratch /* >= 32 doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; ++i) r += scratch[i];  // fixed order: deterministic
  return r;
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double r = scratch[0];
  for (int i = 1; i < nw; ++i) r = fmax(r, scratch[i]);
  return r;
}

// Work item of the two DMMA kernels: a block (plan slot q) or, when a shard leaves too few blocks to fill the GPU,
// a segment [begin, end) of its j-tiles (k_density) / rounds (k_vmat).
struct WorkItem {
  int q;
  short begin, end;
  short seg, nseg;  // this item is piece `seg` of the `nseg` pieces the block was cut into (0 of 1: the whole block)
};

// Device view of a shell table (structure of arrays), see sxc_add_basis.
struct ShellView {
  int nshell;
  int nbf;
  const int* l;
  const int* pure;
  const int* nprim;
  const int* prim_off;
  const int* first_bf;
  const int* nfunc;
  const double* centre;  // [3*nshell]
  const double* alpha;
  const double* coeff;
  const double* normfac;  // [nbf]
  double radial_thr;
  double exp_thr;  // -log(radial_thr)
};

// Device view of a grid (SoA) and of the blocks this context owns.
struct GridView {
  long npts;
  int blocksize;
  int nblocks;  // all blocks of the grid
  const double* x;
  const double* y;
  const double* z;
  const double* w;
};

// Per (grid, basis) screening plan, slot q = position in the owned-block list.
struct PlanView {
  int nown;                // owned blocks
  const int* block_id;     // [nown] grid block index of slot q
  int* nsig_shell;         // [nown]
  int* s;                  // [nown] significant functions
  int* sig_shell;          // [nown * nshell]
  int* sig_c0;             // [nown * nshell] first compact index of each significant shell
  int* sig_bf;             // [nown * nbf_pad] compact index -> basis function
  int nbf_pad;             // row stride of sig_bf
  const int* s_pad;        // [nown] padded s (multiple of SPAD), valid after plan creation
  const long long* phi_off;  // [nown] offset (doubles) of the block's tile inside the current chunk buffer
};

}  // namespace sxc
