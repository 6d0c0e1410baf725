#!/usr/bin/env python
"""Builds profiles/r02_scaling.md from the bench lines tools/scaling_run_r2.sh left in gpurun_out/ (r2_scale_n<N>.json) and copies
those lines to profiles/r02_scale_n<N>.json.  usage: python tools/scaling_table_r2.py"""
import glob
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = {}
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r2_scale_n*.json"))):
    m = re.match(r"r2_scale_n(\d+)\.json", os.path.basename(path))
    line = next((l for l in open(path) if l.startswith("{")), None)
    if m and line:
        lines[int(m.group(1))] = json.loads(line)
        shutil.copy(path, os.path.join(ROOT, "profiles", "r02_scale_n%s.json" % m.group(1)))
out = ["# Round 2 - strong scaling on one B200 box (gpurun --gpus N, tools/scaling_run_r2.sh; table by tools/scaling_table_r2.py)", "",
       "Grid blocks of ONE molecule sharded over N ranks (contiguous cost-balanced ranges), one ncclAllReduce of [V|E|N] per build issued",
       "inside the library (sxc_comm_init_rank); no torch.distributed call in a timed region.  `device` = P resident, V left in HBM",
       "(kernels + all-reduce, max over ranks); `e2e` = sxc_build_xc on caller-owned pageable host buffers on every rank.", "",
       "| workload | N | device ms | speed-up | efficiency | e2e ms | e2e speed-up | e2e efficiency | all-reduce ms | max/mean kernel time over ranks | parity |",
       "|---|---|---|---|---|---|---|---|---|---|---|"]
names = ["tetracene", "water64", "peptide"]


def entries(d):
    yield d["config"]["name"], d
    for w in d.get("workloads", []):
        yield w["name"], w


table = {}
for n, d in lines.items():
    for name, w in entries(d):
        table.setdefault(name, {})[n] = w
for name in names:
    if name not in table or 1 not in table[name]:
        continue
    base = table[name][1]
    for n in sorted(table[name]):
        w = table[name][n]
        su = base["ms_per_step"] / w["ms_per_step"]
        se = base["e2e"]["ms_per_step"] / w["e2e"]["ms_per_step"]
        par = w.get("parity") or {}
        out.append("| %s | %d | %.3f | %.2f | %.3f | %.3f | %.2f | %.3f | %.3f | %s | %s |" % (
            name, n, w["ms_per_step"], su, su / n, w["e2e"]["ms_per_step"], se, se / n, w["kernels_ms_per_build"].get("allreduce", 0.0),
            "%.3f" % w["rank_balance"]["max_over_mean"] if w.get("rank_balance") else "-",
            "dE %.1e, dV %.1e" % (par.get("dE_xc", float("nan")), par.get("max_dV_xc", float("nan"))) if par else "-"))
out += ["", "## Per-kernel time of rank 0 (ms per build) - where the efficiency goes", "",
        "| workload | N | k_basis | k_density | k_functional | k_form_g | scatter | finish | all-reduce | sum | ideal (N = 1 sum / N) |",
        "|---|---|---|---|---|---|---|---|---|---|---|"]
for name in names:
    if name not in table or 1 not in table[name]:
        continue
    k1 = table[name][1]["kernels_ms_per_build"]
    s1 = sum(v for k, v in k1.items() if k != "allreduce")
    for n in sorted(table[name]):
        k = table[name][n]["kernels_ms_per_build"]
        out.append("| %s | %d | %.3f | %.3f | %.3f | %.3f | %.3f | %.3f | %.3f | %.3f | %.3f |" % (
            name, n, k["k_basis"], k["k_density"], k["k_functional"], k["k_form_g"], k["k_scatter"], k["finish"], k["allreduce"],
            sum(k.values()), s1 / n))
fused = [(name, n) for name in names for n in sorted(table.get(name, {}))
         if table[name][n]["kernels_ms_per_build"]["k_form_g"] == 0.0 and table[name][n]["kernels_ms_per_build"]["k_scatter"] > 0.0]
if fused:
    out += ["", "Lines with k_form_g = 0 (" + ", ".join("%s N = %d" % x for x in fused) + ") were measured while the fused scatter `k_vmat_fg` "
            "(`SXC_VMAT=24`) was the default for whole-block shards: their scatter column contains the G phase.  On that pool of boxes "
            "the fused and the two-launch build take the same time to within 1 % (profiles/r02_scatter_fused.md)."]
clk = [(n, d.get("clocks", {})) for n, d in sorted(lines.items())]
out += ["", "Clocks under load: " + "; ".join("N=%d %s MHz %s" % (n, c.get("sm_mhz"), c.get("reasons")) for n, c in clk), ""]
open(os.path.join(ROOT, "profiles", "r02_scaling.md"), "w").write("\n".join(out))
print("\n".join(out))
