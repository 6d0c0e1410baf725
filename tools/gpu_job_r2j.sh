#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2j.log; : > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
rm -f gpurun_out/r2j_sweep.jsonl
for wl in tetracene water64 peptide; do
  bash tools/variant_sweep.sh gpurun_out/r2j_sweep.jsonl $wl "SXC_VMAT=24" "SXC_VMAT=24 SXC_FG_MODE=2"
done
python tools/sweep_summary.py gpurun_out/r2j_sweep.jsonl >> $L
cat $L | cut -c1-300
