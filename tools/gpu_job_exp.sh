#!/bin/bash
# GPU-box job: bench.py with each listed env setting ("name:VAR=1,VAR2=x" ...), no tests
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; var=${spec#*:}
  env ${var//,/ } timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - gpurun_out/bench_$name.json $name <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); print(sys.argv[2], "ms=%.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["kernels_ms_per_build"].items()}, d["result"])
PY
done
