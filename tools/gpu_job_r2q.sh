#!/bin/bash
# e2e with 4 and 8 copy threads on the two large-matrix workloads
mkdir -p gpurun_out
L=gpurun_out/r2q.log; : > $L
for v in "SXC_COPY_THREADS=4" "SXC_COPY_THREADS=8" "SXC_COPY_THREADS=12"; do
for wl in water64 peptide; do
  echo "== $v $wl" >> $L; env $v timeout 300 python bench.py --workload $wl --workloads none --no-cpu-baseline --no-parity --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'pinned', round(d['e2e']['pinned']['ms_per_step'],3))" >> $L
done
done
cat $L
