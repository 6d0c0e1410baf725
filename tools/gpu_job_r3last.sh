#!/bin/bash
# last check of HEAD (round 2, third session): smoke, the GPU suite, the default bench line
mkdir -p gpurun_out
L=gpurun_out/r3last.log; : > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3last_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3last_pytest.log; tail -3 gpurun_out/r3last_pytest.log >> $L
timeout 600 python bench.py > gpurun_out/r3last_bench_n1.json 2> gpurun_out/r3last_bench_n1.err
cat $L; python - <<'P'
import json
d=json.loads(open('gpurun_out/r3last_bench_n1.json').read().strip().splitlines()[-1])
for n,w in [('tetracene',d)]+[(w['name'],w) for w in d['workloads']]:
    e=w.get('e2e') or {}
    print(n, round(w['ms_per_step'],3), 'e2e', round(e.get('ms_per_step',0),3), round(w['roofline']['frac'],3), (w.get('parity') or {}).get('within'))
P
