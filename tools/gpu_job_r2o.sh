#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2o.log; : > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
for w in 8 4; do
for v in "SXC_FG_SEG=1" "SXC_FG_SEG=0"; do
  echo "== emulate-world $w $v" >> $L; env $v timeout 300 python bench.py --workloads none --no-cpu-baseline --no-e2e --emulate-world $w --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernels_ms_per_build'].items()}, d.get('parity',{}).get('within'))" >> $L
done
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.log; tail -3 gpurun_out/r2o_pytest.log >> $L
bash tools/scaling_run_r2.sh 1 >> $L 2>&1
cat $L | cut -c1-330
