#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2g.log; : > $L
step() { echo "== $*" >> $L; }
step smoke; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
rm -f gpurun_out/r2g_sweep.jsonl
for wl in tetracene water64 peptide; do
  bash tools/variant_sweep.sh gpurun_out/r2g_sweep.jsonl $wl "SXC_BASIS=1" "SXC_BASIS=2"
done
python tools/sweep_summary.py gpurun_out/r2g_sweep.jsonl >> $L
for v in "SXC_BASIS=1" "SXC_BASIS=2"; do
  step "emulate-world 8 $v"; env $v timeout 300 python bench.py --workloads none --no-cpu-baseline --no-e2e --no-parity --emulate-world 8 --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernels_ms_per_build'].items()})" >> $L
done
cat $L | cut -c1-300
