#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2c_sweep.jsonl
for wl in tetracene water64 peptide; do
  bash tools/variant_sweep.sh gpurun_out/r2c_sweep.jsonl $wl "SXC_VMAT=24" "SXC_VMAT=24 SXC_FG_MODE=2"
done
python tools/sweep_summary.py gpurun_out/r2c_sweep.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
