#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r3final.log; : > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3final_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3final_pytest.log; tail -3 gpurun_out/r3final_pytest.log >> $L
timeout 900 python bench.py > gpurun_out/r3final_bench_n1.json 2> gpurun_out/r3final_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3final_bench_reference.json 2> gpurun_out/r3final_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3final_launches.csv \
  python bench.py --steps 10 --warmup 3 --workloads none --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r3final_launches_run.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_|k_vmat|k_basis' --launch-skip 18 -c 6 -o gpurun_out/r3final_full_tetracene -f \
  python bench.py --steps 2 --warmup 3 --workloads none --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r3final_full_tetracene.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'^k_|k_vmat|k_basis' --launch-skip 18 -c 6 -o gpurun_out/r3final_full_peptide -f \
  python bench.py --workload peptide --steps 2 --warmup 3 --workloads none --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r3final_full_peptide.log 2>&1
true
cat $L; tail -c 200 gpurun_out/r3final_bench_n1.json; echo; tail -c 300 gpurun_out/r3final_bench_reference.json
