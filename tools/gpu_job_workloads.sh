#!/bin/bash
# GPU-box job: the other BASELINE configs on one GPU (bench lines without the CPU leg) + the FDE iteration tool
mkdir -p gpurun_out
for wl in h2o water64 peptide; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_n1.json 2> gpurun_out/bench_${wl}_n1.err
  python - gpurun_out/bench_${wl}_n1.json $wl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); print(sys.argv[2], "pts=%d nb=%d ms=%.3f e2e=%.3f Mpts/s=%.1f frac=%.3f" % (d["config"]["grid_points"], d["config"]["basis_functions"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"] / 1e6, d["roofline"]["frac"]), {k: round(v, 3) for k, v in d["kernels_ms_per_build"].items()}, "pad=%.3f" % (d["config"]["sum_n_s2_padded"] / d["config"]["sum_n_s2"]))
PY
done
timeout 900 python tools/fde_bench.py fde_water64 10 > gpurun_out/fde_water64.json 2> gpurun_out/fde_water64.err; cat gpurun_out/fde_water64.json; tail -2 gpurun_out/fde_water64.err
timeout 300 python tools/fde_bench.py fde_dimer 10 > gpurun_out/fde_dimer.json 2> gpurun_out/fde_dimer.err; cat gpurun_out/fde_dimer.json
