#!/bin/bash
# round 2, third session: streamed pageable downloads (staged_d2h) - GPU suite, default bench line, copy-thread sweep on the
# 18.9 MB matrix of (H2O)64, and the --set full capture of (H2O)64 that roofline.traffic of that workload reads
mkdir -p gpurun_out
L=gpurun_out/r3a.log; : > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3a_pytest.log; tail -3 gpurun_out/r3a_pytest.log >> $L
timeout 900 python bench.py > gpurun_out/r3a_bench_n1.json 2> gpurun_out/r3a_bench_n1.err
rm -f gpurun_out/r3a_sweep.jsonl
bash tools/variant_sweep.sh gpurun_out/r3a_sweep.jsonl water64 "SXC_COPY_THREADS=4" "SXC_COPY_THREADS=16" >> $L 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'^k_|k_vmat|k_basis' --launch-skip 18 -c 6 -o gpurun_out/r3a_full_water64 -f \
  python bench.py --workload water64 --steps 2 --warmup 3 --workloads none --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r3a_full_water64.log 2>&1
cat $L; python - <<'P'
import json
d=json.loads(open('gpurun_out/r3a_bench_n1.json').read().strip().splitlines()[-1])
print('tetracene', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'pinned', round(d['e2e']['pinned']['ms_per_step'],3), d['parity']['within'])
for w in d['workloads']:
    e=w.get('e2e',{})
    print(w['name'], round(w['ms_per_step'],3), 'e2e', round(e.get('ms_per_step',0),3), 'pinned', round(e.get('pinned',{}).get('ms_per_step',0),3), w.get('parity',{}).get('within'))
for l in open('gpurun_out/r3a_sweep.jsonl'):
    v=json.loads(l); ln=v['line']
    print(v['variant'], ln.get('ms_per_step'), ln.get('e2e',{}).get('ms_per_step'))
P
