#!/usr/bin/env python
"""bench.py - XC / embedding-potential build throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tetracene] [--impl b200|reference]
                    [--workloads h2o,water64,fde_water64,peptide | none]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one FuncPotential::getMatrix()-equivalent XC build: given grid, shell table and P,
produce V_xc (nb x nb) and E_xc.  Headline workload at every N: BASELINE.json configs[1], tetracene C18H12 B3LYP/def2-TZVP
on the accuracy-6 grid (synthetic: ideal geometry, random PSD density matrix scaled to N_el; serenity_b200/inputs).  For
N > 1 the SAME grid is sharded over the ranks (contiguous cost-balanced block ranges) and the partial [V | E | N] are summed
by ONE ncclAllReduce issued INSIDE the library (sxc_comm_init_rank; no torch.distributed call in any timed region - torch
only bootstraps the 128-byte NCCL id, the barriers and the max over ranks): total work is fixed -> "scaling": "strong".

Printed by rank 0: ONE JSON line.
  value         grid-pts/s with P already resident in HBM and V left in HBM (kernels + all-reduce), K steps bracketed by
                barrier + synchronize, CUDA events on the launching stream, max over ranks
  e2e           same metric through the reference-facing C-ABI call sxc_build_xc with CALLER-OWNED (pageable) host buffers:
                P host -> H2D -> build -> all-reduce -> D2H of V (rank 0), E, N; `e2e.pinned` = same call with the caller's
                buffers in page-locked memory (sxc_host_alloc)
  roofline      dominant kernel: algorithmic FP64 flops per launch / live CUDA-event duration, vs a live cuBLAS DGEMM
  parity        |dE_xc|, max|dV_xc| of the timed build against the CPU oracle on the full grid (rank 0, every N), with the
                functionals whose oracle is pinned / unpinned to reference-produced numbers
  workloads     the other BASELINE configs (h2o = configs[0], water64 = [2], fde_water64 = [3], peptide = [4]) at the same N:
                ms_per_step, e2e, per-kernel ms, roofline, parity
  cpu_baseline  the CPU oracle (restatement of the reference's OpenMP path, oracle/) on this box's host cores (N = 1)
`--impl reference` times that CPU path alone (the reference itself cannot be compiled here).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "xc_potential_build_throughput"
UNIT = "grid-pts/s"
EXTRA_WORKLOADS = "h2o,water64,fde_water64,peptide"
TOL = {"dE_xc": 1e-9, "max_dV_xc": 1e-8}
# functionals whose CPU-oracle arithmetic is pinned to reference-produced numbers (H2/def2-TZVP V_xc matrices of
# FuncPotential_test.cpp, tests/test_reference_kats.py) and those that restate published formulas only
PINNED = {2: "slaterx", 45: "vwn5c", 80: "beckex", 193: "p86c"}
UNPINNED = {66: "tfk", 81: "beckecorrx", 135: "pbex", 184: "lypc", 197: "pbec", 283: "pw91k", 286: "llp91k"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="tetracene",
                    help="headline workload: tetracene | water64 | peptide | h2o | water8 (serenity_b200.inputs.make_config)")
    ap.add_argument("--workloads", default=EXTRA_WORKLOADS,
                    help="comma-separated further workloads reported under 'workloads' (or 'none')")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity legs (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--emulate-world", type=int, default=0,
                    help="development aid: time rank 0's shard of a W-rank run on one GPU (no collective)")
    return ap.parse_args()


def functional_of(name):
    from serenity_b200.inputs.configs import FUNCTIONALS
    return FUNCTIONALS[name]


def parity_lists(id_lists):
    ids = sorted({i for l in id_lists for i in l})
    return ([PINNED[i] for i in ids if i in PINNED], [UNPINNED.get(i, str(i)) for i in ids if i not in PINNED])


def config_dict(cfg, n_gpus):
    """Identical in both arms (the driver compares it): what the workload IS, nothing measured."""
    sub = cfg.subsystems[0]
    return {"workload": cfg.description, "name": cfg.name, "functional": cfg.functional, "grid_points": cfg.npts,
            "blocks": (cfg.npts + 127) // 128, "basis_functions": sub.basis.nbf, "shells": sub.basis.nshell,
            "blocksize": 128, "radial_threshold": 1e-9, "block_ave_threshold": 1e-11, "spin": "restricted",
            "sharding": "grid blocks over %d rank(s), one all-reduce of [V|E|N]" % n_gpus,
            "l2": "no flush: every build streams its phi / grad-phi tiles (GBs, far beyond the 126 MB L2) between two uses of "
                  "any input"}


# ------------------------------------------------------------------------------------------------- CPU legs (oracle)
def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def oracle_setup():
    """All host cores of this process, AVX-512 build when the CPU has it, dense products through OpenBLAS (stand-in for the
    reference's Eigen GEMM); returns (module, description of the arithmetic back end)."""
    from oracle import pyoracle as orc
    # (torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not inherit that)
    orc.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    gemm = orc.use_openblas(True)
    gemm["vector_isa"] = orc.variant()
    return orc, gemm


def time_oracle(cfg, ids, mix, steps, warmup, target_s=4.0, P=None):
    """Bounded sample of the workload: every k-th 128-point block of its grid (k chosen so that one oracle build takes
    about target_s on this box's cores; k = 1 is the full grid)."""
    import ctypes as C
    import numpy as np
    orc, gemm = oracle_setup()
    sub = cfg.subsystems[0]
    P = sub.P if P is None else P
    nblk = (cfg.npts + 127) // 128
    probe_k = 16 if nblk >= 64 else 1

    def pick(k):
        if k == 1:
            return cfg.xyz, cfg.w
        idx = np.concatenate([np.arange(b * 128, min((b + 1) * 128, cfg.npts)) for b in range(0, nblk, k)])
        return np.ascontiguousarray(cfg.xyz[idx]), np.ascontiguousarray(cfg.w[idx])

    ob, of = orc.Basis(sub.basis), orc.Functional(ids, mix)
    xyz, w = pick(probe_k)
    t0 = time.perf_counter()
    orc.build_xc(ob, orc.Grid(xyz, w, 128), of, P)
    t_probe = time.perf_counter() - t0
    k = 1
    while t_probe * probe_k / k > target_s and k < 64:
        k *= 2
    xyz, w = pick(k)
    og = orc.Grid(xyz, w, 128)
    for _ in range(warmup):
        orc.build_xc(ob, og, of, P)
    times, phases = [], None
    L = orc.lib()
    L.orc_probe_reset()
    for _ in range(steps):
        t0 = time.perf_counter()
        _, _, _, tm = orc.build_xc(ob, og, of, P)
        times.append(time.perf_counter() - t0)
        phases = tm
    pr = (C.c_double * 3)()
    L.orc_probe_get(pr)
    t = sum(times) / len(times)
    gemm["gflops_per_core"] = (pr[2] / pr[1] / 1e9) if pr[1] > 0 else None
    gemm["thread0_share"] = {"basis_function_evaluation_s": pr[0] / steps, "dense_products_s": pr[1] / steps}
    sample = ("%d of %d grid points (every %s128-point block of the workload), full XC build per step, "
              "%d steps" % (len(w), cfg.npts, "" if k == 1 else "%d-th " % k, steps))
    return {"value": len(w) / t, "unit": UNIT, "cores": orc.max_threads(), "kind": "port", "sample": sample,
            "cpu_model": cpu_model(), "omp_proc_bind": os.environ.get("OMP_PROC_BIND", "unset"),
            "omp_places": os.environ.get("OMP_PLACES", "unset"), "gemm": gemm,
            "s_per_build_sample": t,
            "phases_s": {"density_on_grid": phases.density_on_grid, "functional": phases.functional,
                         "grid_to_matrix": phases.grid_to_matrix}}


CPU_KEYS = ("value", "unit", "cores", "kind", "sample", "phases_s", "cpu_model", "omp_proc_bind", "omp_places", "gemm")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from serenity_b200.inputs import make_config
    cfg = make_config(args.workload)
    ids, mix = functional_of(cfg.functional)
    res = time_oracle(cfg, ids, mix, max(1, args.steps), max(0, args.warmup))
    t = res["s_per_build_sample"]
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, args.gpus),
            "cpu_baseline": {k: res[k] for k in CPU_KEYS},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU restatement of Serenity's OpenMP path (oracle/, kind=port): the reference needs Eigen3, libint2, "
                    "xcfun/libxc, HDF5 and cannot be compiled in this image"}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "no samples"}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return float("nan")
        return {"sm_mhz": statistics.median(num(r[0]) for r in rows), "sm_max_mhz": num(rows[0][1]),
                "power_w_max": max(num(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------- B200 arm
def dgemm_peak_tflops(torch, dev):
    """Live FP64 ceiling for the DMMA kernels: cuBLAS DGEMM 4096^3 through torch.matmul, best of 5 (burst)."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


class Env:
    """Process-wide plumbing of the B200 arm: torch device / stream, rank bookkeeping, barriers, event timing."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL writes its version / debug lines to stdout; keep stdout for the ONE JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed + K timed calls, barrier + synchronize on both sides, CUDA events on the launching stream (the library
        runs on torch's current stream), max over ranks.  Returns (ms per step, wall t0, wall t1, per-rank ms list)."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        t1 = time.perf_counter()
        mine = e0.elapsed_time(e1) / steps
        ranks = [mine]
        if self.world > 1:
            buf = torch.zeros(self.world, dtype=torch.float64, device=self.dev)
            buf[self.rank] = mine
            self.dist.all_reduce(buf)
            ranks = [float(x) for x in buf.tolist()]
        return max(ranks), t0, t1, ranks

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def make_context(env):
    """One library context per process / GPU; with N > 1 ranks its NCCL communicator is created from an id that rank 0 draws
    (sxc_comm_unique_id) and torch broadcasts - the only thing torch.distributed carries for the library."""
    from serenity_b200.xc import XCContext
    ctx = XCContext(env.local)
    # torch's default stream is the legacy NULL stream (handle 0); the C ABI reads NULL as "the context's own stream", so name
    # the legacy stream explicitly (cudaStreamLegacy == (cudaStream_t)0x1)
    ctx.set_stream(env.torch.cuda.current_stream(env.dev).cuda_stream or 1)
    if env.world > 1:
        box = [XCContext.comm_unique_id() if env.rank == 0 else None]
        env.dist.broadcast_object_list(box, src=0)
        ctx.comm_init_rank(env.rank, env.world, box[0])
    return ctx


def roofline_of(env, st, per_build_ms, n_launch, nbf, peak_tf, hbm_peak, hbm_src, ms_dev, cfg_name, top=None, top_launches=None):
    """Dominant kernel of rank 0's shard (or the kernel named by `top`) against its ceiling + the whole build against both
    ceilings.  top_launches: contraction launches behind that timer slot when the slot also times small helper kernels."""
    top = top or max(per_build_ms, key=per_build_ms.get)
    n_launch = dict(n_launch)
    if top_launches:
        n_launch[top] = top_launches
    n_top = max(1, n_launch.get(top, 1))
    ms_launch = per_build_ms[top] / n_top
    # each contraction is 2 n s^2 flops; a launch covers one workspace chunk of the shard's blocks (several
    # launches per build with one chunk = several full contractions: two spins, or one scatter per NAdd functional)
    flops_launch = 2.0 * st["sum_ns2"] / max(1, st.get("nchunks", 1))
    if top in ("k_density", "k_scatter"):
        achieved = flops_launch / (ms_launch * 1e-3) / 1e12
        roof = {"kernel": {"k_scatter": "k_vmat"}.get(top, top), "bound": "tensor", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": "FP64 cuBLAS DGEMM 4096^3 measured live (MEASURED_PEAKS.json has no FP64 figure; B200 nominal "
                               "FP64 = FP64-tensor = 40 TFLOP/s)",
                "algorithmic_flops_per_launch": flops_launch}
    else:
        bytes_build = 32.0 * st["npts"] + 16.0 * st["sum_s2"] + 16.0 * nbf * nbf
        achieved = bytes_build / (ms_launch * n_top * 1e-3) / 1e9
        roof = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "peak_source": hbm_src, "algorithmic_bytes_per_build": bytes_build}
    roof["ms_per_launch"] = ms_launch
    roof["launches_per_build"] = n_top
    # both contractions, each against the DGEMM ceiling (k_scatter = k_vmat_fg on a full GPU: the formation of G - an HBM phase
    # of 40 * 128 * s_pad bytes per block - runs inside it, so its time is not tensor time alone)
    roof["contractions"] = {
        {"k_scatter": "k_vmat"}.get(k, k): {"ms_per_launch": per_build_ms[k] / max(1, n_launch.get(k, 1)),
                                             "tflops": flops_launch / (per_build_ms[k] / max(1, n_launch.get(k, 1)) * 1e-3) / 1e12,
                                             "frac_of_dgemm": flops_launch / (per_build_ms[k] / max(1, n_launch.get(k, 1)) * 1e-3) /
                                             1e12 / peak_tf}
        for k in ("k_density", "k_scatter") if per_build_ms.get(k, 0.0) > 0.0}
    # dram__bytes of that kernel per launch from a committed `ncu --set full` capture of THIS workload at THIS rank count
    # (profiles/traffic.json: {workload: {"n<N>": {kernel: bytes}}}); null when no such capture exists
    roof["traffic"] = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tr = json.load(fh)
        roof["traffic"] = tr.get(cfg_name, {}).get("n%d" % env.world, {}).get(roof["kernel"])
    except (OSError, ValueError, AttributeError):
        pass
    t_build = ms_dev * 1e-3
    bytes_fused = 32.0 * st["npts"] + 16.0 * st["sum_s2"] + 16.0 * nbf * nbf
    tile_bytes = 4.0 * 8.0 * 128.0 * st["sum_s"]  # phi + grad phi tiles as materialised (written once)
    roof["build"] = {"gemm_tflops": 4.0 * st["sum_ns2"] / t_build / 1e12,
                     "gemm_frac_of_dgemm": 4.0 * st["sum_ns2"] / t_build / 1e12 / peak_tf,
                     "fused_bytes_gbs": bytes_fused / t_build / 1e9, "tile_bytes_written": tile_bytes,
                     "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
                     "note": "rank 0's shard over the all-rank build time"}
    return roof


def kernel_times(env, ctx, fn, reps=5):
    """Per-kernel CUDA-event times of one build (separate short pass: the events cost a few microseconds per kernel)."""
    ctx.set_timing(True)
    acc, nacc = {}, {}
    for _ in range(reps):
        fn()
        s = ctx.stats()
        for k, v in s["ms_kernel"].items():
            acc[k] = acc.get(k, 0.0) + v
            nacc[k] = nacc.get(k, 0) + s["n_kernel"][k]
    ctx.set_timing(False)
    return {k: v / reps for k, v in acc.items()}, {k: v // reps for k, v in nacc.items()}


def measure_ks(env, ctx, cfg, steps, warmup, peaks, headline):
    """One KS-DFT workload (FuncPotential): device-resident value, e2e through sxc_build_xc, kernels, roofline, parity."""
    import numpy as np
    torch, args = env.torch, env.args
    ids, mix = functional_of(cfg.functional)
    sub = cfg.subsystems[0]
    nbf = sub.basis.nbf
    n2 = nbf * nbf
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    if env.world == 1 and args.emulate_world > 1:
        ctx.set_grid_shard(g, 0, args.emulate_world)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)

    # caller-owned host buffers of the reference-facing call (what Eigen matrices are: ordinary pageable memory)
    P_host = np.asfortranarray(sub.P, dtype=np.float64)
    V_host = np.zeros((nbf, nbf), order="F") if env.rank == 0 else None
    # scale P so that the grid integrates to N_el; first build also creates the screening plan
    _, ne0 = ctx.build_xc_into(g, b, f, P_host, V_host)
    P_host = np.asfortranarray(sub.P * (sub.n_electrons / ne0))
    E0, ne0 = ctx.build_xc_into(g, b, f, P_host, V_host)
    V0 = V_host.copy(order="F") if V_host is not None else None
    st = ctx.stats()
    launches_per_build = st["kernel_launches"]

    d_P = torch.from_numpy(P_host.reshape(-1, order="F").copy()).to(env.dev)
    d_VEN = torch.zeros(n2 + 2, dtype=torch.float64, device=env.dev)

    def dev_build():
        ctx.build_xc_device(g, b, f, d_P.data_ptr(), d_VEN.data_ptr(), 1e-11)

    ms_dev, t0, t1, rank_ms = env.timed(dev_build, steps, warmup)
    out = {"name": cfg.name, "ms_per_step": ms_dev, "value": cfg.npts / (ms_dev * 1e-3), "unit": UNIT}
    if env.world > 1:
        out["rank_ms_per_step"] = rank_ms
    # per-kernel CUDA events right after the timed loop (same thermal / power state as `value`)
    per_build_ms, n_launch = kernel_times(env, ctx, dev_build)
    if not args.no_e2e:
        def host_build():
            ctx.build_xc_into(g, b, f, P_host, V_host)
        ms_e2e, _, t1, _ = env.timed(host_build, steps, warmup)
        Pp = ctx.pinned_array((nbf, nbf))
        Pp[...] = P_host
        Vp = ctx.pinned_array((nbf, nbf)) if env.rank == 0 else None

        def host_build_pinned():
            ctx.build_xc_into(g, b, f, Pp, Vp)
        ms_pin, _, t1, _ = env.timed(host_build_pinned, steps, warmup)
        out["e2e"] = {"value": cfg.npts / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                      "h2d_bytes_per_step": n2 * 8 * env.world, "d2h_bytes_per_step": n2 * 8 + 16 * env.world,
                      "api": "sxc_build_xc (C ABI) on caller-owned pageable host buffers: every rank uploads P over its own "
                             "PCIe link behind the launches of the screening / basis kernels, build, ncclAllReduce of [V|E|N] "
                             "inside the library, D2H of V on rank 0 and of E, N on every rank, stream synchronize",
                      "pinned": {"value": cfg.npts / (ms_pin * 1e-3), "ms_per_step": ms_pin,
                                 "note": "same call, caller's P / V in page-locked memory (sxc_host_alloc)"}}

    out["kernels_ms_per_build"] = per_build_ms
    if headline:
        # NOT the headline: the same build with sxc_set_tile_cache(1) - phi / grad phi tiles of the previous build kept in HBM
        # (they depend on grid and basis only), so k_screen and k_basis run in the first build of an SCF only.
        ctx.set_tile_cache(True)
        ms_cached, _, _, _ = env.timed(dev_build, max(3, steps // 2), 2)
        ctx.set_tile_cache(False)
        out["with_tile_cache"] = {"ms_per_step": ms_cached, "value": cfg.npts / (ms_cached * 1e-3), "unit": UNIT,
                                  "note": "optional sxc_set_tile_cache(1): basis-function tiles resident across SCF "
                                          "iterations; not used for value / e2e / roofline"}
    stats_all = env.gather({"st": st, "ms": per_build_ms})
    if env.world > 1:
        tot = [sum(x["ms"].values()) for x in stats_all]
        out["rank_balance"] = {"kernel_ms_sum_per_rank": tot, "max_over_mean": max(tot) / (sum(tot) / len(tot)),
                               "per_rank_kernels_ms": [x["ms"] for x in stats_all]}
    out["roofline"] = roofline_of(env, st, per_build_ms, n_launch, nbf, peaks["dgemm"], peaks["hbm"], peaks["hbm_src"], ms_dev,
                                  cfg.name)
    out["stats"] = {"sum_n_s2": st["sum_ns2"], "sum_n_s2_padded": st["sum_ns2_padded"], "s_max": st["s_max"],
                    "s_mean": st["sum_s"] / max(1, st["nblocks"]), "chunks": st["nchunks"],
                    "shard_points": [x["st"]["npts"] for x in stats_all], "launches_per_build": launches_per_build}
    out["result"] = {"E_xc": E0, "n_electrons_on_grid": ne0}
    out["_timing_window"] = (t0, t1)
    out["_launches"] = launches_per_build
    if env.rank == 0 and not args.no_parity:
        pin, unpin = parity_lists([ids])
        try:
            orc, gemm = oracle_setup()
            tt = time.perf_counter()
            Vo, Eo, No, _ = orc.build_xc(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix), P_host)
            tt = time.perf_counter() - tt
            out["parity"] = {"dE_xc": float(abs(E0 - Eo)), "max_dV_xc": float(np.abs(V0 - Vo).max()),
                             "dN_el": float(abs(ne0 - No)), "against": "CPU oracle, same P, full grid", "tolerance": TOL,
                             "within": bool(abs(E0 - Eo) <= TOL["dE_xc"] and np.abs(V0 - Vo).max() <= TOL["max_dV_xc"]),
                             "pinned": pin, "unpinned": unpin,
                             "oracle_s_full_grid": tt, "oracle_cores": orc.max_threads(), "oracle_gemm": gemm["gemm"]}
        except Exception as exc:  # the oracle is a checker; its absence must not hide the GPU number
            out["parity"] = {"failed": str(exc), "pinned": pin, "unpinned": unpin}
    del d_P, d_VEN
    ctx.release_grid(g)
    ctx.release_basis(b)
    return out, P_host


def measure_fde(env, ctx, cfg, steps, warmup, peaks):
    """BASELINE configs[3]: one freeze-and-thaw iteration's grid work for the active subsystem = the non-additive XC (PBE) and
    kinetic (PW91k) potentials on the supersystem grid with the environment density frozen (FDEPotentials.cpp:43-61)."""
    import numpy as np
    torch, args = env.torch, env.args
    act, envs = cfg.subsystems[0], cfg.subsystems[1:]
    names = [cfg.functional, cfg.nadd_kin]
    fdefs = [functional_of(n) for n in names]
    nA = act.basis.nbf
    n2 = nA * nA
    nenv = len(envs)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    ba = ctx.add_basis(act.basis, 1e-9)
    be = [ctx.add_basis(e.basis, 1e-9) for e in envs]
    fh = [ctx.set_functional(*d) for d in fdefs]
    PA = np.asfortranarray(act.P, dtype=np.float64)
    PE = [np.asfortranarray(e.P, dtype=np.float64) for e in envs]
    TAG = 7  # one environment state for the whole run
    # parity inputs: every object's own matrix and energies (also creates the plans and the frozen-environment cache)
    Vs, Es = ctx.build_nadd_multi(g, fh, ba, PA, be, PE, env_frozen=TAG, sum_matrices=False)
    Vsum, Esum = ctx.build_nadd_multi(g, fh, ba, PA, be, PE, env_frozen=TAG, sum_matrices=True)
    st = ctx.stats()
    d_PA = torch.from_numpy(PA.reshape(-1, order="F").copy()).to(env.dev)
    d_PE = [torch.from_numpy(p.reshape(-1, order="F").copy()).to(env.dev) for p in PE]
    pe_ptrs = [t.data_ptr() for t in d_PE]
    ne = 2 + nenv
    d_out = torch.zeros(2 * n2 + 2 * ne, dtype=torch.float64, device=env.dev)
    import ctypes as C
    lib, h = ctx._lib, ctx._h
    fh_arr = np.ascontiguousarray(fh, dtype=np.int32)
    be_arr = np.ascontiguousarray(be, dtype=np.int32)
    pp = (C.c_void_p * max(nenv, 1))(*pe_ptrs)

    def multi(sum_mode):
        ctx._check(lib.sxc_build_nadd_multi_device(h, g, 2, fh_arr.ctypes.data_as(C.c_void_p), 1, ba, C.c_void_p(d_PA.data_ptr()),
                                                   nenv, be_arr.ctypes.data_as(C.c_void_p), pp, TAG, 1e-11, sum_mode,
                                                   C.c_void_p(d_out.data_ptr())))

    def two_objects():  # what two independent NAddFuncPotential objects cost: everything twice
        for k in range(2):
            ctx.build_nadd_device(g, fh[k], ba, d_PA.data_ptr(), be, pe_ptrs, d_out.data_ptr(), TAG, 1e-11)

    ms_sep, t0, t1, rank_ms = env.timed(lambda: multi(0), steps, warmup)
    ms_sum, _, _, _ = env.timed(lambda: multi(1), steps, warmup)
    ms_two, _, t1, _ = env.timed(two_objects, steps, warmup)
    out = {"name": cfg.name, "ms_per_step": ms_sep, "value": cfg.npts / (ms_sep * 1e-3), "unit": UNIT,
           "step": "one freeze-and-thaw iteration of the active subsystem: NAdd XC (%s) + NAdd kinetic (%s), frozen environment, "
                   "one device pass (sxc_build_nadd_multi_device), every object's own matrix and energies" % tuple(names),
           "variants_ms": {"one_pass_separate_matrices": ms_sep, "one_pass_summed_matrix": ms_sum,
                           "two_independent_objects": ms_two}}
    if env.world > 1:
        out["rank_ms_per_step"] = rank_ms
    if not args.no_e2e:
        V_host = np.zeros(2 * n2) if env.rank == 0 else None
        E_host = np.zeros(2 * ne)
        pph = (C.c_void_p * max(nenv, 1))(*[p.ctypes.data for p in PE])

        def host_build():
            ctx._check(lib.sxc_build_nadd_multi(h, g, 2, fh_arr.ctypes.data_as(C.c_void_p), 1, ba, PA.ctypes.data_as(C.c_void_p), nenv,
                                                be_arr.ctypes.data_as(C.c_void_p), pph, TAG, 1e-11, 0,
                                                None if V_host is None else V_host.ctypes.data_as(C.c_void_p),
                                                E_host.ctypes.data_as(C.c_void_p)))
        ms_e2e, _, t1, _ = env.timed(host_build, steps, warmup)
        out["e2e"] = {"value": cfg.npts / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                      "h2d_bytes_per_step": n2 * 8 * env.world, "d2h_bytes_per_step": 2 * n2 * 8 + 2 * ne * 8 * env.world,
                      "api": "sxc_build_nadd_multi (C ABI) on caller-owned pageable host buffers; the frozen environment's "
                             "density matrices are not uploaded again"}
    per_build_ms, n_launch = kernel_times(env, ctx, lambda: multi(0))
    out["kernels_ms_per_build"] = per_build_ms
    # the one-pass iteration contracts rho_act once and scatters twice: 3 contractions of 2 n s^2 flops
    t_build = ms_sep * 1e-3
    # the density contraction runs over every block; the scatters skip the blocks whose weighted non-additive potential is below
    # blockAveThreshold (ScalarOperatorToMatrixAdder.cpp:262-268, most of the grid far from the subsystem interface), so their
    # executed flops are fewer than 2 n s^2 and an "achieved" rate from the algorithmic count would overstate them
    # (the k_density timer slot of this build holds ONE contraction - the active density; the frozen environment's is cached -
    # plus the k_add4 launches that sum the subsystem densities)
    out["roofline"] = roofline_of(env, st, per_build_ms, n_launch, nA, peaks["dgemm"], peaks["hbm"], peaks["hbm_src"], ms_sep, cfg.name,
                                  top="k_density", top_launches=1)
    out["roofline"]["contractions"].pop("k_vmat", None)
    out["roofline"]["build"]["gemm_tflops"] = 6.0 * st["sum_ns2"] / t_build / 1e12
    out["roofline"]["build"]["gemm_frac_of_dgemm"] = 6.0 * st["sum_ns2"] / t_build / 1e12 / peaks["dgemm"]
    out["roofline"]["build"]["note"] = ("3 contractions per iteration by the algorithmic count (one density, two scatters; the scatters skip the blocks below "
                                        "blockAveThreshold, so this is an upper bound of the executed work); rank 0's shard")
    out["stats"] = {"sum_n_s2": st["sum_ns2"], "s_max": st["s_max"], "s_mean": st["sum_s"] / max(1, st["nblocks"]),
                    "nbf_active": nA, "nbf_environment": [e.basis.nbf for e in envs]}
    out["_timing_window"] = (t0, t1)
    if env.rank == 0 and not args.no_parity:
        pin, unpin = parity_lists([d[0] for d in fdefs])
        try:
            orc, gemm = oracle_setup()
            og, oa = orc.Grid(cfg.xyz, cfg.w, 128), orc.Basis(act.basis)
            oe = [(orc.Basis(e.basis), p) for e, p in zip(envs, PE)]
            par = {"against": "CPU oracle (NAddFuncPotential restatement), same P, full grid", "tolerance": TOL,
                   "pinned": pin, "unpinned": unpin, "objects": {}}
            tt = time.perf_counter()
            Vref_sum = 0.0
            ok = True
            for k, nm in enumerate(names):
                V_ref, E_ref, _ = orc.build_nadd(oa, PA, oe, og, orc.Functional(*fdefs[k]))
                Vref_sum = Vref_sum + V_ref
                e = Es[k]
                dE = float(abs(e[0] - e[1] - e[2:].sum() - E_ref))  # E_nadd = E[tot] - E[act] - sum E[env]
                dV = float(np.abs(Vs[k] - V_ref).max())
                par["objects"][nm] = {"dE_nadd": dE, "max_dV_nadd": dV}
                ok = ok and dE <= TOL["dE_xc"] and dV <= TOL["max_dV_xc"]
            par["max_dV_summed_matrix"] = float(np.abs(Vsum - Vref_sum).max())
            par["dE_summed_vs_separate"] = float(np.abs(Esum - Es).max())
            par["within"] = bool(ok and par["max_dV_summed_matrix"] <= TOL["max_dV_xc"])
            par["oracle_s_full_grid"] = time.perf_counter() - tt
            par["oracle_cores"] = orc.max_threads()
            out["parity"] = par
        except Exception as exc:
            out["parity"] = {"failed": str(exc), "pinned": pin, "unpinned": unpin}
    del d_PA, d_PE, d_out
    ctx.release_grid(g)
    for hb in [ba] + be:
        ctx.release_basis(hb)
    return out


def run_b200(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:  # convenience: relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29511"] + sys.argv
            return subprocess.call(cmd, stdout=sys.stdout)  # (fd 1 of this process points at stderr, see main())
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the XC build has no CPU fallback (use --impl reference for the CPU arm)")
    from serenity_b200.inputs import make_config
    env = Env(args)
    ctx = make_context(env)
    rank = env.rank

    peaks = {"dgemm": dgemm_peak_tflops(torch, env.dev)}
    mp = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mp = json.load(fh)
    except OSError:
        pass
    peaks["hbm"], peaks["hbm_src"] = (mp["hbm_gbs"], "MEASURED_PEAKS.json") if "hbm_gbs" in mp else (6650.0, "fallback")

    cfg = make_config(args.workload)
    sampler = ClockSampler(env.local) if rank == 0 else None
    head, P_scaled = measure_ks(env, ctx, cfg, args.steps, args.warmup, peaks, headline=True)
    clocks = sampler.stop(*head.pop("_timing_window")) if sampler else None  # samples span the device-resident and e2e regions
    launches = head.pop("_launches")

    extra = []
    names = [] if args.workloads in ("", "none") else [n.strip() for n in args.workloads.split(",") if n.strip()]
    steps_x, warm_x = max(3, min(args.steps, 10)), max(3, min(args.warmup, 3))
    grids = {}
    for name in names:
        if name == cfg.name:
            continue
        try:
            base = name[4:] if name.startswith("fde_") else name
            wcfg = make_config(name, grid=grids.get(base))
            grids[base] = (wcfg.xyz, wcfg.w)
            if name.startswith("fde_"):
                r = measure_fde(env, ctx, wcfg, steps_x, warm_x, peaks)
            else:
                r, _ = measure_ks(env, ctx, wcfg, steps_x, warm_x, peaks, headline=False)
                r.pop("_launches", None)
            r.pop("_timing_window", None)
            r["config"] = config_dict(wcfg, env.world)
            r["steps"], r["warmup"] = steps_x, warm_x
        except Exception as exc:  # one failing side workload must not hide the headline
            r = {"name": name, "failed": "%s: %s" % (type(exc).__name__, exc)}
        extra.append(r)

    comm = ctx.comm_info()
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": env.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "s_per_build": head["ms_per_step"] * 1e-3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(cfg, env.world), "clocks": clocks, "e2e": head.get("e2e"),
                "gpu_launches": launches * args.steps * env.world,
                "collective": {"library": "NCCL %s bound by libserenity_xc_b200.so (sxc_comm_init_rank)" % comm["nccl_version"],
                               "ranks": comm["world"], "allreduces_issued_by_rank0": comm["collectives"],
                               "torch_distributed_in_timed_region": False} if env.world > 1 else None}
        for k in ("kernels_ms_per_build", "roofline", "with_tile_cache", "stats", "result", "parity", "rank_ms_per_step",
                  "rank_balance"):
            if k in head:
                line[k] = head[k]
        line["workloads"] = extra
    ctx.close()
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    if rank == 0:
        if env.world == 1 and not args.no_cpu_baseline:
            try:
                ids, mix = functional_of(cfg.functional)
                cb = time_oracle(cfg, ids, mix, steps=3, warmup=1, P=P_scaled)
                line["cpu_baseline"] = {k: cb[k] for k in CPU_KEYS}
            except Exception as exc:  # the oracle is a checker; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %s" % exc}
        print(json.dumps(line), flush=True)
    return 0


def main():
    args = parse_args()
    # stdout carries exactly ONE line, the JSON: libraries that print from C (NCCL's version banner under NCCL_DEBUG=VERSION
    # ignores NCCL_DEBUG_FILE on some builds) write to file descriptor 1, so fd 1 points at stderr while the run lasts and
    # Python's sys.stdout keeps the real one
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
