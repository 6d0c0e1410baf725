#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2k.log; : > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
SXC_VMAT=24 timeout 300 python bench.py --workloads none --no-cpu-baseline --no-e2e --steps 5 --warmup 2 2>&1 | tail -c 300 >> $L
grep -q '"within": true' $L || { tail -8 $L | cut -c1-300; exit 1; }
rm -f gpurun_out/r2k_sweep.jsonl
for wl in tetracene water64 peptide; do
  bash tools/variant_sweep.sh gpurun_out/r2k_sweep.jsonl $wl "SXC_FG_PRE=0" "SXC_FG_PRE=15" "SXC_FG_PRE=30" "SXC_FG_PRE=45" "SXC_FG_PRE=60" "SXC_FG_PRE=100"
done
python tools/sweep_summary.py gpurun_out/r2k_sweep.jsonl >> $L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2k_pytest.log; tail -3 gpurun_out/r2k_pytest.log >> $L
cat $L | cut -c1-300
