#!/bin/bash
# validation first (each step under its own timeout; stop at the first failure), then measurements
mkdir -p gpurun_out
L=gpurun_out/r2d.log; : > $L
step() { echo "== $*" >> $L; }
step smoke; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
step memcheck; timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|Invalid|smoke:" | head >> $L
step synccheck; timeout 300 compute-sanitizer --tool synccheck --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|Barrier|smoke:|divergent" | head >> $L
step tetracene; SXC_VMAT=24 timeout 300 python bench.py --workloads none --no-cpu-baseline --no-e2e --steps 5 --warmup 2 2>&1 | tail -c 400 >> $L
grep -q '"within": true' $L || { tail -8 $L | cut -c1-300; exit 1; }
step pytest; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log; tail -3 gpurun_out/r2d_pytest.log >> $L
rm -f gpurun_out/r2d_sweep.jsonl
for wl in tetracene water64 peptide; do
  bash tools/variant_sweep.sh gpurun_out/r2d_sweep.jsonl $wl "SXC_VMAT=24" "SXC_VMAT=24 SXC_FG_MODE=2" "SXC_VMAT=24 SXC_DPF=0" "SXC_VMAT=8 SXC_DPF=0"
done
python tools/sweep_summary.py gpurun_out/r2d_sweep.jsonl >> $L
# emulated rank 0 of 8 (segmented work items): fused against unfused
for v in "SXC_VMAT=24" "SXC_VMAT=8"; do
  step "emulate-world 8 $v"; env $v timeout 300 python bench.py --workloads none --no-cpu-baseline --no-e2e --no-parity --emulate-world 8 --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernels_ms_per_build'].items()})" >> $L
done
cat $L | cut -c1-400
