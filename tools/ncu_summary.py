#!/usr/bin/env python
"""Summarise ncu outputs into small tracked files under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.md>     per-kernel launch counts, device time and share
  python tools/ncu_summary.py full <report.ncu-rep> <out.md> [traffic.json workload [n<ranks>]]
        per-kernel roofline-relevant counters of an `ncu --set full` capture (read with `ncu -i ... --page raw --csv`)
"""
import collections
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (regs)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math-pipe"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long-sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short-sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg-throttle"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smem ld bank conflicts"),
]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, vi, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if k.startswith("sxc::"))
    with open(out, "w") as f:
        f.write("| kernel | launches | total ms | avg us | share of all | share of sxc kernels |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            own = "%.3f" % (v[1] / mine) if k.startswith("sxc::") else "-"
            f.write("| `%s` | %d | %.3f | %.1f | %.3f | %s |\n" % (k[:90], v[0], v[1] / 1e6, v[1] / v[0] / 1e3, v[1] / tot, own))
        f.write("\n(ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and "
                "serialised - compare shares, not absolutes)\n")


def full(rep, out, traffic_json=None, workload=None, ranks="n1"):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    seen, traffic = set(), {}
    with open(out, "w") as f:
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
            if name in seen:
                continue
            seen.add(name)
            f.write("### `%s`\n\n| metric | value |\n|---|---|\n" % name)
            for m, label in METRICS:
                if m in col:
                    f.write("| %s (`%s`) | %s %s |\n" % (label, m, r[col[m]], units[col[m]]))
            f.write("\n")

            def byt(m):
                v, u = float(r[col[m]].replace(",", "")), units[col[m]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
            short = name.replace("sxc::", "").split("<")[0]
            traffic[short] = byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum")
            if short in ("k_vmat_fg", "k_vmat_tma"):  # bench.py names the scatter contraction "k_vmat" whichever variant ran
                traffic.setdefault("k_vmat", traffic[short])
    if traffic_json:
        try:
            d = json.load(open(traffic_json))
        except (OSError, ValueError):
            d = {}
        d.setdefault(workload, {}).setdefault(ranks, {}).update(traffic)  # {workload: {"n<N>": {kernel: bytes per launch}}}
        json.dump(d, open(traffic_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
