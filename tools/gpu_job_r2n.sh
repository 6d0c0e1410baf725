#!/bin/bash
# compute-sanitizer on the small workloads that reach every DMMA kernel; summaries for profiles/r02_sanitizer.md
mkdir -p gpurun_out
L=gpurun_out/r2n.log; : > $L
T="tests/test_ab_potential.py::test_gpu_ab_reference_kats tests/test_xc_gradient.py::test_gpu_gradient_matches_oracle tests/test_gpu_parity.py::test_golden_density_hessian tests/test_gpu_parity.py::test_shards_sum_to_full_build tests/test_gpu_configs.py::test_config1_h2o_accuracy4"
for tool in memcheck synccheck; do
  echo "== $tool" >> $L
  timeout 900 compute-sanitizer --tool $tool --print-limit 3 python -m pytest -q -x $T 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|Barrier" | head -6 >> $L
done
echo "== racecheck (analysis report)" >> $L
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 200 python -m pytest -q -x $T > gpurun_out/r2n_racecheck_full.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r2n_racecheck_full.log >> $L
# hazards grouped by kernel and by the pair of source lines
python - <<'PY' >> $L
import re, collections
txt = open("gpurun_out/r2n_racecheck_full.log").read()
blocks = re.split(r"=========\s*\n", txt)
cnt = collections.Counter()
for b in re.findall(r"========= (?:Error|Warning): .*?(?=\n=========\s*\n|\Z)", txt, flags=re.S):
    kern = re.search(r"in (?:void )?(sxc::\w+)", b)
    lines = sorted(set(re.findall(r"(\w+\.cuh?):(\d+)", b)))
    kind = re.search(r"(Race reported|Potential \w+ hazard|\w+ hazard)", b)
    cnt[(kern.group(1) if kern else "?", kind.group(1) if kind else "?", tuple(lines[:4]))] += 1
for (k, kind, ls), n in cnt.most_common(30):
    print(n, k, kind, ls)
PY
cat $L | cut -c1-250
