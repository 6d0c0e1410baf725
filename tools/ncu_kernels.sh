#!/bin/bash
# Development aid: one `ncu --set full` capture of the DMMA kernels of a tetracene build (1 GPU). usage: tools/ncu_kernels.sh <tag> [ENV=val ...]
tag=$1; shift
env "$@" ncu --set full --import-source on --clock-control none -k regex:'k_density|k_vmat' --launch-skip 8 -c 2 -f -o gpurun_out/$tag \
  python bench.py --workloads none --no-cpu-baseline --no-parity --no-e2e --steps 2 --warmup 3 > gpurun_out/$tag.log 2>&1
