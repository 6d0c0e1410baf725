#!/usr/bin/env python
"""BASELINE config 4 measurement: one freeze-and-thaw iteration's grid work for the active subsystem on one B200 -
NAddFuncPotential (XC, PBE) + NAddFuncPotential (kinetic, PW91k) on the supersystem grid, environment density frozen
(FDEPotentials.cpp:43-61) - through the host C ABI (sxc_build_nadd), next to the CPU oracle on the box's cores.

  python tools/fde_bench.py [fde_water64|fde_dimer] [steps]
Prints one JSON line (device time from the library's CUDA-event timers, wall time of the call, parity vs the oracle)."""
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
    name = sys.argv[1] if len(sys.argv) > 1 else "fde_water64"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    cfg = make_config(name)
    act, env = cfg.subsystems
    ctx = XCContext(0)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    ba, be = ctx.add_basis(act.basis, 1e-9), ctx.add_basis(env.basis, 1e-9)
    funcs = {"xc": ctx.set_functional(*FUNCTIONALS[cfg.functional]), "kin": ctx.set_functional(*FUNCTIONALS[cfg.nadd_kin])}
    out = {}
    for key, f in funcs.items():  # first calls: environment density + plans (and the parity inputs)
        V, E = ctx.build_nadd(g, f, ba, act.P, [be], [env.P], env_frozen=False)
        out[key] = {"V": V, "E": E, "dev": [], "wall": []}
    # one freeze-and-thaw iteration = the XC object, then the kinetic object (they alternate, FDEPotentials.cpp:43-61)
    for _ in range(steps):
        for key, f in funcs.items():
            t0 = time.perf_counter()
            V, E = ctx.build_nadd(g, f, ba, act.P, [be], [env.P], env_frozen=True)
            out[key]["wall"].append((time.perf_counter() - t0) * 1e3)
            out[key]["dev"].append(ctx.stats()["ms_total"])
            out[key]["V"], out[key]["E"] = V, E
    for key in funcs:
        out[key]["device_ms"] = float(np.median(out[key]["dev"]))
        out[key]["call_ms"] = float(np.median(out[key]["wall"]))
    # optional tile residency (sxc_set_tile_cache): with a frozen environment the active system's basis-function tiles stay valid
    # across the XC and kinetic objects and across iterations - reported next to the default, not instead of it
    ctx.set_tile_cache(True)
    cached = []
    for _ in range(steps + 1):
        t = 0.0
        for f in funcs.values():
            ctx.build_nadd(g, f, ba, act.P, [be], [env.P], env_frozen=True)
            t += ctx.stats()["ms_total"]
        cached.append(t)
    ctx.set_tile_cache(False)
    line = {"workload": cfg.description, "name": cfg.name, "grid_points": cfg.npts, "nbf_active": act.basis.nbf,
            "nbf_environment": env.basis.nbf, "steps": steps,
            "nadd_xc_device_ms": out["xc"]["device_ms"], "nadd_kin_device_ms": out["kin"]["device_ms"],
            "nadd_xc_call_ms": out["xc"]["call_ms"], "nadd_kin_call_ms": out["kin"]["call_ms"],
            "iteration_device_ms": out["xc"]["device_ms"] + out["kin"]["device_ms"],
            "iteration_device_ms_with_tile_cache": float(np.median(cached[1:])),
            "grid_pts_per_s": cfg.npts / ((out["xc"]["device_ms"] + out["kin"]["device_ms"]) * 1e-3)}
    if "--no-oracle" not in sys.argv:
        from oracle import pyoracle as orc
        og, oa, oe = orc.Grid(cfg.xyz, cfg.w, 128), orc.Basis(act.basis), orc.Basis(env.basis)
        t_cpu = 0.0
        for key, fn in (("xc", cfg.functional), ("kin", cfg.nadd_kin)):
            t0 = time.perf_counter()
            V_ref, E_ref, _ = orc.build_nadd(oa, act.P, [(oe, env.P)], og, orc.Functional(*FUNCTIONALS[fn]))
            t_cpu += time.perf_counter() - t0
            line["max_dV_" + key] = float(np.abs(out[key]["V"] - V_ref).max())
            E = out[key]["E"]
            line["dE_" + key] = float(abs(E[0] - E[1] - E[2:].sum() - E_ref))  # E_nadd = E[tot] - E[act] - sum E[env]
        line["cpu_oracle_s"] = t_cpu  # (recomputes the environment density, which the device path caches)
        line["cpu_threads"] = orc.max_threads()
    print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
