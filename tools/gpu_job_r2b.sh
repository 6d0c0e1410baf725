#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2b_sweep.jsonl
for wl in tetracene peptide; do
  bash tools/variant_sweep.sh gpurun_out/r2b_sweep.jsonl $wl "SXC_VMAT=24" "SXC_VMAT=24 SXC_FG_MODE=1" "SXC_VMAT=83" "SXC_VMAT=8"
done
python tools/sweep_summary.py gpurun_out/r2b_sweep.jsonl
