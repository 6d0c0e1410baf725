#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2l.log; : > $L
rm -f gpurun_out/r2l_sweep.jsonl
bash tools/variant_sweep.sh gpurun_out/r2l_sweep.jsonl tetracene "SXC_FG_PRE=0" "SXC_FG_PRE=0 SXC_FG_MODE=16" "SXC_FG_PRE=0 SXC_FG_MODE=32" "SXC_FG_PRE=0 SXC_FG_MODE=48" "SXC_FG_PRE=0 SXC_FG_MODE=1"
python tools/sweep_summary.py gpurun_out/r2l_sweep.jsonl >> $L
cat $L | cut -c1-300
