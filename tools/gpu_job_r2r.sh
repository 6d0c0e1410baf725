#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2r.log; : > $L
rm -f gpurun_out/r2r_sweep.jsonl
for rep in 1 2; do
for wl in tetracene water64 peptide; do
  bash tools/variant_sweep.sh gpurun_out/r2r_sweep.jsonl $wl "SXC_VMAT=24" "SXC_VMAT=8" "SXC_VMAT=16"
done
done
python tools/sweep_summary.py gpurun_out/r2r_sweep.jsonl >> $L
echo "== trace" >> $L
SXC_TRACE=1 timeout 300 python bench.py --workloads none --no-cpu-baseline --no-parity --steps 6 --warmup 3 2>&1 >/dev/null | grep "sxc_build_xc:" | tail -8 >> $L
SXC_TRACE=1 timeout 300 python bench.py --workload water64 --workloads none --no-cpu-baseline --no-parity --steps 6 --warmup 3 2>&1 >/dev/null | grep "sxc_build_xc:" | tail -8 >> $L
cat $L | cut -c1-260
