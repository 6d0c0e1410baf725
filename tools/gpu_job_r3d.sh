#!/bin/bash
# round 2, third session: L2 prefetch of k_density's epilogue rows (SXC_DPF=1) on the small-s workloads, where the epilogue is a third of the kernel
mkdir -p gpurun_out
rm -f gpurun_out/r3d_dpf.jsonl
for v in 0 1; do
  line=$(SXC_DPF=$v timeout 300 python bench.py --workload water64 --workloads peptide --no-cpu-baseline --no-parity --no-e2e --steps 10 --warmup 3 2>gpurun_out/r3d.err)
  echo "{\"SXC_DPF\": $v, \"line\": ${line:-null}}" >> gpurun_out/r3d_dpf.jsonl
done
python - <<'P'
import json
for l in open('gpurun_out/r3d_dpf.jsonl'):
    d=json.loads(l); ln=d['line']
    if not ln: print(d['SXC_DPF'],'FAILED'); continue
    row=[(ln['config']['name'],ln)]+[(w['name'],w) for w in ln['workloads']]
    print(d['SXC_DPF'], [(n, round(w['ms_per_step'],3), round(w['kernels_ms_per_build']['k_density'],4)) for n,w in row])
P
tail -2 gpurun_out/r3d.err
