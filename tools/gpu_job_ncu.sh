#!/bin/bash
# GPU-box job: source-level ncu capture of the kernels matching $1 (regex), skipping the warm-up builds.
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$1" --launch-skip ${2:-3} -c ${3:-1} \
  -o gpurun_out/${4:-cap} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_${4:-cap}.log 2>&1
tail -3 gpurun_out/ncu_${4:-cap}.log | cut -c1-300
