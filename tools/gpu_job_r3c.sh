#!/bin/bash
# round 2, third session: k_functional at 3 / 4 / 5 CTAs per SM (SXC_FUNC) on tetracene (B3LYP), fde_water64 (PBE + PW91k) and the peptide (PBE)
mkdir -p gpurun_out
rm -f gpurun_out/r3c_func.jsonl
for v in 0 4 5; do
  line=$(SXC_FUNC=$v timeout 600 python bench.py --workloads fde_water64,peptide --no-cpu-baseline --no-parity --no-e2e --steps 10 --warmup 3 2>gpurun_out/r3c.err)
  echo "{\"SXC_FUNC\": $v, \"line\": ${line:-null}}" >> gpurun_out/r3c_func.jsonl
done
python - <<'P'
import json
for l in open('gpurun_out/r3c_func.jsonl'):
    d=json.loads(l); ln=d['line']
    if not ln: print(d['SXC_FUNC'],'FAILED'); continue
    row=[('tetracene',ln)]+[(w['name'],w) for w in ln['workloads']]
    print(d['SXC_FUNC'], [(n, round(w['ms_per_step'],3), round(w['kernels_ms_per_build']['k_functional'],4)) for n,w in row])
P
tail -3 gpurun_out/r3c.err
