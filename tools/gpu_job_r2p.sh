#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2p.log; : > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
SXC_VMAT=24 timeout 300 python bench.py --workloads none --no-cpu-baseline --no-e2e --steps 10 --warmup 3 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tetracene', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernels_ms_per_build'].items()}, d.get('parity',{}).get('within'))" >> $L
for w in 8 4; do
for v in "SXC_FG_SEG=1" "SXC_FG_SEG=0"; do
  echo "== emulate-world $w $v" >> $L; env $v timeout 300 python bench.py --workloads none --no-cpu-baseline --no-e2e --no-parity --emulate-world $w --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernels_ms_per_build'].items()})" >> $L
done
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2p_pytest.log; tail -3 gpurun_out/r2p_pytest.log >> $L
cat $L | cut -c1-330
