#!/bin/bash
# GPU-box job: parity tests under an env setting ($1, "VAR=1"), then bench.py with and without it
mkdir -p gpurun_out
env $1 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for spec in "new:$1" "old:X=1"; do
  name=${spec%%:*}; var=${spec#*:}
  env $var timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - gpurun_out/bench_$name.json $name <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); print(sys.argv[2], "ms=%.3f e2e=%.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["kernels_ms_per_build"].items()}, d["result"])
PY
done
