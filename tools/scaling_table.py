#!/usr/bin/env python
"""Builds profiles/r01_scaling.md from the bench lines tools/scaling_run.sh left in gpurun_out/ (bench_r1_<workload>_n<N>.json)
and copies those lines to profiles/r01_scale_<workload>_n<N>.json.  usage: python tools/scaling_table.py [round_tag]"""
import glob
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = {}
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "bench_r1_*_n*.json"))):
    m = re.match(r"bench_r1_(\w+)_n(\d+)\.json", os.path.basename(path))
    line = next((l for l in open(path) if l.startswith("{")), None)
    if not m or not line:
        continue
    d = json.loads(line)
    rows.setdefault(m.group(1), {})[int(m.group(2))] = d
    shutil.copy(path, os.path.join(ROOT, "profiles", "%s_scale_%s_n%s.json" % (tag, m.group(1), m.group(2))))
out = ["# Round 1 - strong scaling on one 8 x B200 box (gpurun --gpus 8, tools/scaling_run.sh; table by tools/scaling_table.py)", "",
       "Grid blocks of ONE molecule sharded over N ranks (contiguous cost-balanced ranges), one NCCL all-reduce of [V|E|N] per build.",
       "`device` = P resident, V left in HBM (kernels + all-reduce); `e2e` = pinned host P -> H2D -> build -> all-reduce -> D2H.", "",
       "| workload | grid points | nb | N | device ms | speed-up | efficiency | Mpts/s | e2e ms | e2e Mpts/s |", "|---|---|---|---|---|---|---|---|---|---|"]
clocks = []
for wl in sorted(rows):
    base = rows[wl].get(1)
    for n in sorted(rows[wl]):
        d = rows[wl][n]
        sp = base["ms_per_step"] / d["ms_per_step"] if base else float("nan")
        out.append("| %s | %d | %d | %d | %.3f | %.2f | %.0f %% | %.1f | %.3f | %.1f |" % (
            wl, d["config"]["grid_points"], d["config"]["basis_functions"], n, d["ms_per_step"], sp, 100.0 * sp / n, d["value"] / 1e6,
            d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6))
        clocks.append("%s n=%d: %.0f MHz %s" % (wl, n, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
out += ["", "Clocks (rank 0, during the timed regions): " + ", ".join(clocks), ""]
open(os.path.join(ROOT, "profiles", tag + "_scaling.md"), "w").write("\n".join(out))
print("\n".join(out))
