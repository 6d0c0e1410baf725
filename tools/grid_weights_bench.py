#!/usr/bin/env python
"""Row f-1 measurement: partition weights (GridFactory.cpp:139-266) on the device vs the host restatement.

  python tools/grid_weights_bench.py [workload ...]     (default: tetracene water64 peptide; SSF, their BASELINE accuracy)
Prints one JSON line per workload: points, atoms, host seconds (all cores, gridweights.c), device kernel ms (CUDA
events inside sxc_partition_weights), whole C-ABI call ms (H2D + kernel + D2H), max |w_device - w_host| / max atomic weight."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from serenity_b200.inputs.configs import geometry_of  # noqa: E402
from serenity_b200.inputs.geometry import atomic_numbers  # noqa: E402
from serenity_b200.inputs.grid import host_partition_weights, reference_atom_grids  # noqa: E402
from serenity_b200.xc import XCContext  # noqa: E402

ACC = {"h2o": 4, "tetracene": 6, "water64": 4, "peptide": 6}


def main():
    names = sys.argv[1:] or ["tetracene", "water64", "peptide"]
    ctx = XCContext(0)
    for name in names:
        symbols, coords = geometry_of(name)
        zs = atomic_numbers(symbols)
        xyz, w0, parent = reference_atom_grids(symbols, coords, ACC[name])
        ctx.partition_weights("SSF", coords, xyz, parent, w0)  # warm-up (module load, allocations)
        t0 = time.perf_counter()
        got, ms = ctx.partition_weights("SSF", coords, xyz, parent, w0)
        call_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        ref = host_partition_weights("SSF", zs, coords, xyz, w0, parent)
        host_s = time.perf_counter() - t0
        dev = float(np.max(np.abs(got - ref)) / np.abs(w0).max())
        print(json.dumps({"workload": name, "flavour": "SSF", "accuracy": ACC[name], "atoms": len(zs), "points": int(w0.shape[0]),
                          "kept_points": int((got > 1e-14).sum()), "host_seconds": round(host_s, 4),
                          "host_threads": os.cpu_count(), "device_kernel_ms": round(ms, 4), "device_call_ms": round(call_ms, 3),
                          "points_per_s_device": w0.shape[0] / (ms * 1e-3), "speedup_kernel": host_s * 1e3 / ms,
                          "speedup_call": host_s * 1e3 / call_ms, "max_abs_dev_over_max_weight": dev}))
    ctx.close()


if __name__ == "__main__":
    main()
