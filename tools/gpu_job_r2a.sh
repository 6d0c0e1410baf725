#!/bin/bash
# GPU-box job (round 2): parity tests first, then the variant sweep of the scatter kernels on three workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
rm -f gpurun_out/r2a_sweep.jsonl
for wl in tetracene water64 peptide; do
  bash tools/variant_sweep.sh gpurun_out/r2a_sweep.jsonl $wl "SXC_VMAT=24" "SXC_VMAT=8" "SXC_VMAT=16"
done
python tools/sweep_summary.py gpurun_out/r2a_sweep.jsonl
