"""Multi-GPU host layer: grid blocks sharded over the ranks of one node, ONE all-reduce of [V | E | N] per build.

SURVEY.md section 8e: the reference parallelises the path with `#pragma omp parallel for` over 128-point blocks
(MatrixOperatorToGridTransformer.cpp:103, XCFun.cpp:129, ScalarOperatorToMatrixAdder.cpp:66,101) and sums per-thread
nb x nb accumulators serially (ScalarOperatorToMatrixAdder.cpp:73-75).  Here: one process per GPU (torchrun), every rank
holds the grid / shell table / P, evaluates a contiguous cost-balanced range of blocks (sxc_set_grid_shard) and the
partial sums are combined by a single `all_reduce(SUM)` of nspin*nb*nb + 2 doubles over NCCL (NVLink 5 / NVSwitch).
torch is plumbing only (device buffers, streams, the process group); all numbers come from the CUDA library.

The local builder is injectable so that the reduce / layout logic is testable with the gloo backend on CPU
(tests/test_sharded_gloo.py drives it with the oracle restricted to the rank's block range).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(costs, world: int):
    """Contiguous block ranges of nearly equal summed cost: the host-only helper behind sxc_set_grid_shard."""
    import ctypes as C
    from . import _lib
    costs = np.ascontiguousarray(costs, dtype=np.float64)
    bounds = np.zeros(world + 1, dtype=np.int32)
    rc = _lib.load().sxc_balance_ranges(len(costs), costs.ctypes.data_as(C.c_void_p), world,
                                        bounds.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise _lib.SerenityError("sxc_balance_ranges failed (%d)" % rc)
    return bounds


class ShardedBuild:
    """[V | E | N] of one XC build summed over the ranks of `group`.

    local_build(d_P, d_VEN, p_ready_event_or_None) must enqueue this rank's partial build on the current stream / device of d_VEN
    (CUDA: XCContext.build_xc_device with the context's stream set to torch's current stream).
    """

    def __init__(self, nbf: int, local_build: Callable[[torch.Tensor, torch.Tensor], None], device,
                 group: Optional[dist.ProcessGroup] = None, ntail: int = 2, nspin: int = 1):
        if nspin not in (1, 2):
            raise ValueError("nspin must be 1 (RESTRICTED) or 2 (UNRESTRICTED)")
        self.nbf, self.ntail, self.nspin = nbf, ntail, nspin
        self._local = local_build
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.device = torch.device(device)
        n = nspin * nbf * nbf  # UNRESTRICTED: {alpha, beta} matrices back to back, in P and in V
        self._n = n
        # P travels in `world` equal slices (one per rank), so its device buffer is padded to a multiple of world
        self._slice = (n + self.world - 1) // self.world
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._d_P_all = torch.zeros(self._slice * self.world, dtype=torch.float64, device=self.device)
        self.d_P = self._d_P_all[:n]
        self.d_VEN = torch.zeros(n + ntail, dtype=torch.float64, device=self.device)
        cuda = self.device.type == "cuda"
        self.h_P = torch.zeros(n, dtype=torch.float64, pin_memory=cuda)      # pinned staging buffers of the host API
        self.h_VEN = torch.zeros(n + ntail, dtype=torch.float64, pin_memory=cuda)
        self._copy_stream = torch.cuda.Stream(self.device) if cuda else None
        self._p_event = torch.cuda.Event() if cuda else None

    # device-resident: P already in d_P, result stays in d_VEN (asynchronous on CUDA)
    def build_device(self, p_ready=None) -> torch.Tensor:
        self._local(self.d_P, self.d_VEN, p_ready)
        if self.world > 1:
            dist.all_reduce(self.d_VEN, op=dist.ReduceOp.SUM, group=self.group)
        return self.d_VEN

    def _upload(self):
        """H2D of P: every rank copies ONE slice of the (identical) host matrix over its own PCIe link and the slices
        are all-gathered over NVLink - nb^2 * 8 bytes cross PCIe once per build instead of once per rank."""
        n = self._n
        if self.world == 1:
            self.d_P.copy_(self.h_P, non_blocking=True)
            return
        lo = self.rank * self._slice
        hi = min(lo + self._slice, n)
        mine = self._d_P_all[lo: lo + self._slice]
        if hi > lo:
            mine[: hi - lo].copy_(self.h_P[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(self._d_P_all, mine, group=self.group)  # in place: slice r of the output is rank r's input

    def build_pinned(self):
        """P has been written into the pinned buffer h_P (column-major); returns views into the pinned result buffer
        h_VEN: (V [nb, nb], E, N) - V only on rank 0 (the SCF driver's process), E and N on
        every rank.  The upload runs on a side stream and is awaited by the library only before the density kernel,
        i.e. it overlaps with the screening and basis kernels.  nspin = 2: V is [2, nb, nb] (alpha, beta)."""
        n = self._n
        if self._copy_stream is not None:
            cur = torch.cuda.current_stream(self.device)
            self._copy_stream.wait_stream(cur)  # the previous build no longer reads d_P
            with torch.cuda.stream(self._copy_stream):
                self._upload()
                self._p_event.record(self._copy_stream)
            self.build_device(self._p_event)
            if self.rank == 0:
                self.h_VEN.copy_(self.d_VEN, non_blocking=True)
            else:
                self.h_VEN[n:].copy_(self.d_VEN[n:], non_blocking=True)
            cur.synchronize()
        else:
            self._upload()
            self.build_device()
            self.h_VEN.copy_(self.d_VEN)
        out = self.h_VEN.numpy()
        nb = self.nbf
        if self.nspin == 1:
            V = out[:n].reshape(nb, nb, order="F")
        else:
            V = np.stack([out[k * nb * nb:(k + 1) * nb * nb].reshape(nb, nb, order="F") for k in range(2)])
        return V, float(out[n]), float(out[n + 1])

    # host buffers in, host buffers out: what FuncPotential::getMatrix/getEnergy hand to the SCF driver
    def build(self, P: np.ndarray):
        P = np.asarray(P, dtype=np.float64)
        if self.nspin == 1:
            self.h_P.numpy()[:] = P.reshape(-1, order="F")
        else:
            self.h_P.numpy()[:] = np.concatenate([P[k].reshape(-1, order="F") for k in range(2)])
        V, E, N = self.build_pinned()
        return (V.copy(order="F") if self.nspin == 1 else V.copy()), E, N

    @property
    def h2d_bytes(self) -> int:
        """bytes uploaded per build, all ranks together"""
        return self.h_P.numel() * 8

    @property
    def d2h_bytes(self) -> int:
        """bytes downloaded per build, all ranks together: [V|E|N] on rank 0, [E|N] elsewhere"""
        return self.h_VEN.numel() * 8 + (self.world - 1) * self.ntail * 8


def cuda_local_build(ctx, grid: int, basis: int, func: int, block_ave_threshold: float = 1e-11, nspin: int = 1):
    """local_build for ShardedBuild(..., nspin=nspin) on a CUDA rank: sxc_build_xc_device on torch's current stream."""

    def run(d_P: torch.Tensor, d_VEN: torch.Tensor, p_ready=None):
        if p_ready is not None:
            ctx.set_p_ready_event(p_ready.cuda_event)
        # torch's default stream is the legacy NULL stream (handle 0); the C ABI reads NULL as "the context's own
        # stream", so name the legacy stream explicitly (cudaStreamLegacy == (cudaStream_t)0x1)
        ctx.set_stream(torch.cuda.current_stream(d_P.device).cuda_stream or 1)
        ctx.build_xc_device(grid, basis, func, d_P.data_ptr(), d_VEN.data_ptr(), block_ave_threshold, nspin)

    return run


def cuda_local_nadd_build(ctx, grid: int, func: int, basis_act: int, basis_env, d_P_env, env_frozen: bool = True,
                          block_ave_threshold: float = 1e-11):
    """local_build for ShardedBuild(nbf_act, ..., ntail = 2 + len(basis_env)) on a CUDA rank: one NAddFuncPotential build
    (sxc_build_nadd_device) of the active system against device-resident environment density matrices `d_P_env` (torch
    tensors).  The all-reduced buffer is [V_nadd | E[rho_tot] | E[rho_act] | E[rho_env_i] ...]: every term is a sum over
    grid blocks, so the freeze-and-thaw potentials shard like the KS potential (ShardedBuild.build_pinned returns the first
    two energies; the environment terms stay in h_VEN[nb * nb + 2:])."""
    ptrs = [t.data_ptr() for t in d_P_env]

    def run(d_P: torch.Tensor, d_VE: torch.Tensor, p_ready=None):
        if p_ready is not None:
            ctx.set_p_ready_event(p_ready.cuda_event)
        ctx.set_stream(torch.cuda.current_stream(d_P.device).cuda_stream or 1)
        ctx.build_nadd_device(grid, func, basis_act, d_P.data_ptr(), list(basis_env), ptrs, d_VE.data_ptr(), env_frozen,
                              block_ave_threshold)

    return run


class ShardedSigma:
    """LR-TDDFT kernel sigma build (row f-4) over the ranks of `group`: every rank holds the kernel store of its own grid
    blocks, contracts / integrates all nvec trial vectors on them and the partial Fock-like matrices are summed by ONE
    all_reduce of nvec * nspin * nb * nb doubles (KernelSigmavector.cpp:236-249 sums its per-thread matrices the same way).

    local_sigma(d_D, d_F) must enqueue this rank's partial contraction + integration (CUDA: XCContext.kernel_sigma_device).
    """

    def __init__(self, nbf: int, nvec: int, local_sigma: Callable[[torch.Tensor, torch.Tensor], None], device,
                 group: Optional[dist.ProcessGroup] = None, nspin: int = 1):
        self.nbf, self.nvec, self.nspin = nbf, nvec, nspin
        self._local = local_sigma
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.device = torch.device(device)
        n = nvec * nspin * nbf * nbf
        cuda = self.device.type == "cuda"
        self.d_D = torch.zeros(n, dtype=torch.float64, device=self.device)
        self.d_F = torch.zeros(n, dtype=torch.float64, device=self.device)
        self.h_D = torch.zeros(n, dtype=torch.float64, pin_memory=cuda)
        self.h_F = torch.zeros(n, dtype=torch.float64, pin_memory=cuda)

    def sigma_device(self) -> torch.Tensor:
        self._local(self.d_D, self.d_F)
        if self.world > 1:
            dist.all_reduce(self.d_F, op=dist.ReduceOp.SUM, group=self.group)
        return self.d_F

    def sigma(self, D):
        """D: nvec (x nspin) matrices [nb, nb] -> list of nvec * nspin Fock-like matrices (host buffers in and out)."""
        flat = np.concatenate([np.asarray(m, dtype=np.float64).reshape(-1, order="F") for m in D])
        assert flat.size == self.h_D.numel()
        self.h_D.numpy()[:] = flat
        self.d_D.copy_(self.h_D, non_blocking=True)
        self.sigma_device()
        self.h_F.copy_(self.d_F, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        nb = self.nbf
        out = self.h_F.numpy().reshape(self.nvec * self.nspin, nb * nb)
        return [out[m].reshape(nb, nb, order="F").copy(order="F") for m in range(out.shape[0])]


def cuda_local_sigma(ctx, grid: int, basis: int, kernels, nvec: int, mode: int = 0):
    """local_sigma for ShardedSigma on a CUDA rank: sxc_kernel_contract_device + sxc_kernel_integrate_device on torch's
    current stream."""

    def run(d_D: torch.Tensor, d_F: torch.Tensor):
        ctx.set_stream(torch.cuda.current_stream(d_D.device).cuda_stream or 1)
        ctx.kernel_sigma_device(grid, basis, kernels, d_D.data_ptr(), d_F.data_ptr(), nvec, mode)

    return run
