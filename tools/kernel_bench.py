#!/usr/bin/env python
"""Row f-4 measurement: the LR-TDDFT kernel sigma build (KernelSigmavector::calcF, KernelSigmavector.cpp:119-252) for nvec
trial vectors on one B200 through the host C ABI, next to the CPU oracle on a bounded sample of the same grid.

  python tools/kernel_bench.py [tetracene|h2o|water64] [nvec] [steps]
Prints one JSON line: device times from the library's CUDA-event timers (store set-up, contraction, integration), wall time
of the calls (host buffers: D uploaded, F downloaded), parity against the oracle on the sample."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from serenity_b200.inputs import make_config  # noqa: E402
from serenity_b200.inputs.configs import FUNCTIONALS  # noqa: E402
from serenity_b200.xc import XCContext  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tetracene"
    nvec = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    cfg = make_config(name)
    sub = cfg.subsystems[0]
    nb = sub.basis.nbf
    ids, mix = FUNCTIONALS[cfg.functional]
    ctx = XCContext(0)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    k = ctx.kernel_create(g, 1, True)
    ctx.kernel_add(k, f, [b], [sub.P])          # warm-up: plans, workspace
    ctx.kernel_add(k, f, [b], [sub.P], sign=-1.0)
    t0 = time.perf_counter()
    ctx.kernel_add(k, f, [b], [sub.P])
    add_wall = (time.perf_counter() - t0) * 1e3
    st = ctx.stats()
    add_dev = st["ms_total"]
    rng = np.random.default_rng(0)
    D = [rng.standard_normal((nb, nb)) * 1e-2 for _ in range(nvec)]
    ctx.kernel_sigma(g, b, nb, [k], D, 0)
    con, integ, wall = [], [], []
    per_kernel = {}
    for _ in range(steps):
        t0 = time.perf_counter()
        ctx.kernel_contract(g, b, [k], D, 0, False)
        s1 = ctx.stats()
        F = ctx.kernel_integrate(g, b, nb, nvec, 0)
        s2 = ctx.stats()
        wall.append((time.perf_counter() - t0) * 1e3)
        con.append(s1["ms_total"])
        integ.append(s2["ms_total"])
        per_kernel = {"contract": s1["ms_kernel"], "integrate": s2["ms_kernel"]}
    dev = float(np.median(con) + np.median(integ))
    flops = 4.0 * st["sum_ns2"] * nvec  # two contractions of 2 n s^2 per vector and block
    line = {"workload": cfg.description, "name": cfg.name, "functional": cfg.functional, "grid_points": cfg.npts, "nbf": nb,
            "nvec": nvec, "steps": steps, "store_device_ms": add_dev, "store_call_ms": add_wall,
            "contract_device_ms": float(np.median(con)), "integrate_device_ms": float(np.median(integ)),
            "sigma_device_ms": dev, "sigma_call_ms": float(np.median(wall)), "device_ms_per_vector": dev / nvec,
            "grid_pts_x_vectors_per_s": cfg.npts * nvec / (dev * 1e-3), "gemm_tflops": flops / (dev * 1e-3) / 1e12,
            "ms_kernel": per_kernel}
    if "--no-oracle" not in sys.argv:
        from oracle import pyoracle as orc
        nblk = min(int(os.environ.get("KERNEL_BENCH_SAMPLE_BLOCKS", "1024")), (cfg.npts + 127) // 128)
        npts = min(nblk * 128, cfg.npts)
        # the sample: the first nblk blocks of the same grid (blocks are independent); parity on that sub-grid
        og, ob = orc.Grid(cfg.xyz[:npts], cfg.w[:npts], 128), orc.Basis(sub.basis)
        func = orc.Functional(ids, mix)
        rho, gr, _, _ = orc.density_on_grid(ob, og, 1e-9, sub.P, 1)
        store = orc.kernel_store_r(func, rho, gr)
        t0 = time.perf_counter()
        F_ref = orc.kernel_integrate(ob, og, orc.kernel_contract(ob, og, store, D[0], 0, True), True)
        t_cpu = time.perf_counter() - t0
        # the reference's (i, j) screen at blockAveThreshold = 1e-11 drops terms the device multiplies: compare to both
        F_exact = orc.kernel_integrate(ob, og, orc.kernel_contract(ob, og, store, D[0], 0, True, block_ave_thr=0.0), True,
                                       block_ave_thr=0.0)
        c2 = XCContext(0)
        g2 = c2.set_grid(cfg.xyz[:npts], cfg.w[:npts], 128)
        b2 = c2.add_basis(sub.basis, 1e-9)
        k2 = c2.kernel_create(g2, 1, True)
        c2.kernel_add(k2, c2.set_functional(ids, mix), [b2], [sub.P])
        F2 = c2.kernel_sigma(g2, b2, nb, [k2], [D[0]], 0)[0]
        line["sample"] = "first %d blocks (%d points) of the grid, 1 trial vector" % (nblk, npts)
        line["cpu_oracle_s_sample"] = t_cpu
        line["cpu_threads"] = orc.max_threads()
        line["cpu_grid_pts_x_vectors_per_s"] = npts / t_cpu
        line["max_rel_dF_sample"] = float(np.abs(F2 - F_ref).max() / np.abs(F_ref).max())
        line["max_rel_dF_sample_unscreened_oracle"] = float(np.abs(F2 - F_exact).max() / np.abs(F_exact).max())
        c2.close()
    print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
