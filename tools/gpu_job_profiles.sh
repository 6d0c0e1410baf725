#!/bin/bash
# GPU-box job that regenerates the evidence under profiles/: launch list, --set full capture of every kernel of one
# build, the default bench line and the reference arm.  Outputs land in gpurun_out/; summarise here with
#   python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_tetracene.md
#   python tools/ncu_summary.py full gpurun_out/full.ncu-rep profiles/r01_ncu_full_tetracene.md profiles/traffic.json tetracene
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_run.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_' --launch-skip 24 -c 8 -o gpurun_out/full -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/full_run.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 400 gpurun_out/bench_n1.json; tail -c 600 gpurun_out/bench_reference.json; ls -la gpurun_out | tail -12
