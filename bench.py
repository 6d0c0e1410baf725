#!/usr/bin/env python
"""bench.py - XC / embedding-potential build throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tetracene] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one FuncPotential::getMatrix()-equivalent XC build: given grid, shell table and P,
produce V_xc (nb x nb) and E_xc.  Workload at every N: BASELINE.json configs[1], tetracene C18H12 B3LYP/def2-TZVP on the
accuracy-6 grid (synthetic: ideal geometry, random PSD density matrix scaled to N_el; serenity_b200/inputs).  For N > 1
the SAME grid is sharded over the ranks (contiguous cost-balanced block ranges) and the partial [V | E | N] are summed
by one NCCL all-reduce: total work is fixed -> "scaling": "strong".

Printed by rank 0: ONE JSON line.
  value     grid-pts/s with P already resident in HBM and V left in HBM (kernels + all-reduce), K steps bracketed by
            barrier + synchronize, CUDA events on the launching stream, max over ranks
  e2e       same metric through the host-buffer API (pinned host P -> H2D -> build -> all-reduce -> D2H of V,E,N)
  roofline  dominant kernel: algorithmic FP64 flops per launch / live CUDA-event duration, vs a live cuBLAS DGEMM
  cpu_baseline  the CPU oracle (restatement of the reference's OpenMP path, oracle/) on this box's host cores
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="tetracene",
                    help="tetracene | water64 | peptide | h2o | water8 (serenity_b200.inputs.make_config)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--emulate-world", type=int, default=0,
                    help="development aid: time rank 0's shard of a W-rank run on one GPU (no collective)")
    return ap.parse_args()


def workload(name):
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.configs import FUNCTIONALS
    cfg = make_config(name)
    ids, mix = FUNCTIONALS[cfg.functional]
    return cfg, ids, mix


def config_dict(cfg, n_gpus, extra=None):
    sub = cfg.subsystems[0]
    d = {"workload": cfg.description, "name": cfg.name, "functional": cfg.functional, "grid_points": cfg.npts,
         "blocks": (cfg.npts + 127) // 128, "basis_functions": sub.basis.nbf, "shells": sub.basis.nshell,
         "blocksize": 128, "radial_threshold": 1e-9, "block_ave_threshold": 1e-11, "spin": "restricted",
         "sharding": "grid blocks over %d rank(s), one all-reduce of [V|E|N]" % n_gpus}
    if extra:
        d.update(extra)
    return d


# ------------------------------------------------------------------------------------------------- CPU legs (oracle)
def oracle_sample(cfg, target_s=4.0):
    """Bounded sample for the CPU legs: every k-th 128-point block of the workload's grid (k chosen so that one
    oracle build takes about target_s on this box's cores); k = 1 is the full grid."""
    import numpy as np
    from oracle import pyoracle as orc
    sub = cfg.subsystems[0]
    nblk = (cfg.npts + 127) // 128
    probe_k = 16 if nblk >= 64 else 1

    def pick(k):
        if k == 1:
            return cfg.xyz, cfg.w
        idx = np.concatenate([np.arange(b * 128, min((b + 1) * 128, cfg.npts)) for b in range(0, nblk, k)])
        return np.ascontiguousarray(cfg.xyz[idx]), np.ascontiguousarray(cfg.w[idx])

    return orc, sub, pick, probe_k


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def time_oracle(cfg, ids, mix, steps, warmup, target_s=4.0, P=None):
    """P: density matrix to use instead of the config's (timing does not depend on it); with a full-grid sample the last
    build's (V, E, N_el) come back under 'result' so that the caller can state parity."""
    orc, sub, pick, probe_k = oracle_sample(cfg)
    P = sub.P if P is None else P
    # all host cores of this process (torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not inherit that)
    orc.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
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
    times, phases, last = [], None, None
    for _ in range(steps):
        t0 = time.perf_counter()
        Vo, Eo, No, tm = orc.build_xc(ob, og, of, P)
        last = (Vo, Eo, No) if k == 1 else None
        times.append(time.perf_counter() - t0)
        phases = tm
    t = sum(times) / len(times)
    sample = ("%d of %d grid points (every %s128-point block of the workload), full XC build per step, "
              "%d steps" % (len(w), cfg.npts, "" if k == 1 else "%d-th " % k, steps))
    return {"value": len(w) / t, "unit": UNIT, "cores": orc.max_threads(), "kind": "port", "sample": sample,
            "cpu_model": cpu_model(), "omp_proc_bind": os.environ.get("OMP_PROC_BIND", "unset"),
            "omp_places": os.environ.get("OMP_PLACES", "unset"),
            "s_per_build_sample": t, "result": last,
            "phases_s": {"density_on_grid": phases.density_on_grid, "functional": phases.functional,
                         "grid_to_matrix": phases.grid_to_matrix}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg, ids, mix = workload(args.workload)
    res = time_oracle(cfg, ids, mix, max(1, args.steps), max(0, args.warmup))
    t = res["s_per_build_sample"]
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, args.gpus),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "phases_s", "cpu_model", "omp_proc_bind",
                                                     "omp_places")},
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


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from serenity_b200.sharded import ShardedBuild, cuda_local_build
    from serenity_b200.xc import XCContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:  # convenience: relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29511"] + sys.argv
            return subprocess.call(cmd, stdout=sys.stdout)  # (fd 1 of this process points at stderr, see main())
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the XC build has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its version / debug lines to stdout; keep stdout for the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p")
        dist.init_process_group("nccl", device_id=dev)

    cfg, ids, mix = workload(args.workload)
    sub = cfg.subsystems[0]
    nbf = sub.basis.nbf
    ctx = XCContext(local)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    if world > 1:
        ctx.set_grid_shard(g, rank, world)
    elif args.emulate_world > 1:
        ctx.set_grid_shard(g, 0, args.emulate_world)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    sb = ShardedBuild(nbf, cuda_local_build(ctx, g, b, f, 1e-11), dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # scale P so that the grid integrates to N_el; first build also creates the screening plan
    V0, E0, ne0 = sb.build(sub.P)
    P = np.asfortranarray(sub.P * (sub.n_electrons / ne0))
    V0, E0, ne0 = sb.build(P)
    sb.d_P.copy_(torch.from_numpy(P.reshape(-1, order="F").copy()))
    st = ctx.stats()
    launches_per_build = st["kernel_launches"]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, t0, t1

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, t0, t1 = timed(sb.build_device, args.steps, args.warmup)
    e2e = None
    if not args.no_e2e:
        sb.h_P.numpy()[:] = P.reshape(-1, order="F")  # the caller's P lives in the pinned staging buffer
        ms_e2e, _, t1 = timed(sb.build_pinned, args.steps, args.warmup)
        e2e = {"value": cfg.npts / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": sb.h2d_bytes, "d2h_bytes_per_step": sb.d2h_bytes,
               "api": "ShardedBuild.build_pinned: pinned host P -> H2D (one slice per rank + all-gather; side stream, awaited before k_density) -> sxc_build_xc_device -> all_reduce -> D2H [V|E|N] on rank 0, [E|N] elsewhere -> synchronize"}
    clocks = sampler.stop(t0, t1) if sampler else None  # samples span both timed regions (device-resident and e2e)

    # per-kernel CUDA-event times (separate short pass: the events cost a few microseconds per kernel)
    ctx.set_timing(True)
    acc, nacc, reps = {}, {}, 5
    for _ in range(reps):
        sb.build_device()
        s = ctx.stats()
        for k, v in s["ms_kernel"].items():
            acc[k] = acc.get(k, 0.0) + v
            nacc[k] = nacc.get(k, 0) + s["n_kernel"][k]
    ctx.set_timing(False)
    # NOT the headline: the same build with sxc_set_tile_cache(1) - phi / grad phi tiles of the previous build kept in HBM
    # (they depend on grid and basis only), so k_screen and k_basis run in the first build of an SCF only.  Reported next to
    # the default (the reference's per-build recomputation), never instead of it.
    ctx.set_tile_cache(True)
    ms_cached, _, _ = timed(sb.build_device, max(3, args.steps // 2), 2)
    ctx.set_tile_cache(False)
    per_build_ms = {k: v / reps for k, v in acc.items()}
    top = max(per_build_ms, key=per_build_ms.get)
    n_top = max(1, nacc[top] // reps)
    ms_launch = per_build_ms[top] / n_top

    line = None
    stats_all = [st]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, {"st": st, "ms": per_build_ms})
        stats_all = [x["st"] for x in gathered]
    if rank == 0:
        peak_tf = dgemm_peak_tflops(torch, dev)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json") if "hbm_gbs" in peaks else (6650.0, "fallback")
        # algorithmic work of rank 0's shard: each contraction is 2 n s^2 flops
        flops_contraction = 2.0 * st["sum_ns2"]
        if top in ("k_density", "k_scatter"):
            achieved = flops_contraction / n_top / (ms_launch * 1e-3) / 1e12
            roof = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf,
                    "peak_source": "FP64 cuBLAS DGEMM 4096^3 measured live (MEASURED_PEAKS.json has no FP64 figure; "
                                   "B200 nominal FP64 = FP64-tensor = 40 TFLOP/s)",
                    "algorithmic_flops_per_launch": flops_contraction / n_top}
        else:
            bytes_build = 32.0 * st["npts"] + 16.0 * st["sum_s2"] + 16.0 * nbf * nbf
            achieved = bytes_build / (ms_launch * n_top * 1e-3) / 1e9
            roof = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "peak_source": hbm_src, "algorithmic_bytes_per_build": bytes_build}
        roof["ms_per_launch"] = ms_launch
        roof["launches_per_build"] = n_top
        roof["traffic"] = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tr = json.load(fh)
            roof["traffic"] = tr.get(cfg.name, {}).get({"k_scatter": "k_vmat"}.get(top, top))  # timer slot -> kernel name
        except (OSError, ValueError):
            pass
        # whole-build view against both ceilings (all kernels of one build on rank 0)
        t_build = ms_dev * 1e-3
        bytes_fused = 32.0 * st["npts"] + 16.0 * st["sum_s2"] + 16.0 * nbf * nbf
        tile_bytes = 4.0 * 8.0 * 128.0 * st["sum_s"]  # phi + grad phi tiles as materialised (written once)
        roof["build"] = {"gemm_tflops": 4.0 * st["sum_ns2"] / t_build / 1e12,
                         "gemm_frac_of_dgemm": 4.0 * st["sum_ns2"] / t_build / 1e12 / peak_tf,
                         "fused_bytes_gbs": bytes_fused / t_build / 1e9, "tile_bytes_written": tile_bytes,
                         "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src}
        line = {"metric": METRIC, "value": cfg.npts / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "s_per_build": ms_dev * 1e-3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "with_tile_cache": {"ms_per_step": ms_cached, "value": cfg.npts / (ms_cached * 1e-3), "unit": UNIT,
                                    "note": "optional sxc_set_tile_cache(1): basis-function tiles resident across SCF iterations; "
                                            "not used for value / e2e / roofline"},
                "config": config_dict(cfg, world, {
                    "l2": "no flush: one build streams %.2f GB of phi/grad-phi tiles (>> 126 MB L2) between uses of any "
                          "input" % (tile_bytes / 1e9),
                    "sum_n_s2": st["sum_ns2"], "sum_n_s2_padded": st["sum_ns2_padded"], "s_max": st["s_max"], "s_mean": st["sum_s"] / max(1, st["nblocks"]),
                    "chunks": st["nchunks"], "shard_points": [s_["npts"] for s_ in stats_all]}),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_build * args.steps * world,
                "kernels_ms_per_build": per_build_ms, "roofline": roof,
                "result": {"E_xc": E0, "n_electrons_on_grid": ne0}}
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = time_oracle(cfg, ids, mix, steps=3, warmup=1, P=P)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "phases_s", "cpu_model",
                                                         "omp_proc_bind", "omp_places")}
                if cb.get("result") is not None:  # the sample was the whole grid: parity of this very build
                    Vo, Eo, No = cb["result"]
                    line["parity"] = {"dE_xc": float(abs(E0 - Eo)), "max_dV_xc": float(np.abs(V0 - Vo).max()),
                                      "dN_el": float(abs(ne0 - No)), "against": "CPU oracle, same P, full grid",
                                      "tolerance": {"dE_xc": 1e-9, "max_dV_xc": 1e-8}}
            except Exception as exc:  # the oracle is a checker; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                        "sample": "failed: %s" % exc}
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
