#!/bin/bash
# round 2, third session: k_density epilogue with the 4 x 4 transposed lane reduction - GPU suite and the default bench line
mkdir -p gpurun_out
L=gpurun_out/r3b.log; : > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3b_pytest.log; tail -3 gpurun_out/r3b_pytest.log >> $L
timeout 900 python bench.py > gpurun_out/r3b_bench_n1.json 2> gpurun_out/r3b_bench_n1.err
cat $L; python - <<'P'
import json
d=json.loads(open('gpurun_out/r3b_bench_n1.json').read().strip().splitlines()[-1])
def show(n,w):
    e=w.get('e2e') or {}
    k=w['kernels_ms_per_build']
    print(n, round(w['ms_per_step'],3), 'e2e', round(e.get('ms_per_step',0),3), 'dens', round(k['k_density'],3), 'vmat', round(k['k_scatter'],3), 'frac', round(w['roofline']['frac'],3), (w.get('parity') or {}).get('within'))
show('tetracene', d)
for w in d['workloads']: show(w['name'], w)
P
