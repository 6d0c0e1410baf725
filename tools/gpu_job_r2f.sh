#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2f.log; : > $L
step() { echo "== $*" >> $L; }
step smoke; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
rm -f gpurun_out/r2f_sweep.jsonl
bash tools/variant_sweep.sh gpurun_out/r2f_sweep.jsonl tetracene "SXC_DPF=0" "SXC_DPF=2" "SXC_DPF=1"
bash tools/variant_sweep.sh gpurun_out/r2f_sweep.jsonl peptide "SXC_DPF=0" "SXC_DPF=2"
python tools/sweep_summary.py gpurun_out/r2f_sweep.jsonl >> $L
for v in "SXC_SEG_WAVES=3" "SXC_SEG_WAVES=2" "SXC_SEG_WAVES=4" "SXC_SEG_WAVES=6" "SXC_SEG_WAVES=1"; do
  step "emulate-world 8 $v"; env $v timeout 300 python bench.py --workloads none --no-cpu-baseline --no-e2e --no-parity --emulate-world 8 --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernels_ms_per_build'].items()})" >> $L
done
# sanitizers on the small workloads (k_density, k_vmat_fg / k_vmat_tma, k_vmat_ab, k_grad_contract, k_hessq)
for tool in racecheck synccheck; do
  step "$tool smoke"; timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Barrier|smoke:" | head -8 >> $L
  step "$tool ab + gradient + segmented"; SXC_SEG_WAVES=3 timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest -q -x "tests/test_ab_potential.py::test_gpu_ab_reference_kats" "tests/test_xc_gradient.py::test_gpu_gradient_matches_oracle" "tests/test_gpu_parity.py::test_golden_density_hessian" "tests/test_gpu_parity.py::test_shards_sum_to_full_build" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Barrier|passed|failed" | head -8 >> $L
done
cat $L | cut -c1-300
