#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2m.log; : > $L
SXC_DPF=4 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
rm -f gpurun_out/r2m_sweep.jsonl
for wl in tetracene water64 peptide; do
  bash tools/variant_sweep.sh gpurun_out/r2m_sweep.jsonl $wl "SXC_DPF=0" "SXC_DPF=4"
done
python tools/sweep_summary.py gpurun_out/r2m_sweep.jsonl >> $L
SXC_DPF=4 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.log; tail -3 gpurun_out/r2m_pytest.log >> $L
cat $L | cut -c1-300
