#!/bin/bash
# last check of HEAD: smoke, the GPU suite, the default bench line and the reference arm
mkdir -p gpurun_out
L=gpurun_out/r2last.log; : > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1 || { tail -5 $L; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2last_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2last_pytest.log; tail -3 gpurun_out/r2last_pytest.log >> $L
timeout 900 python bench.py > gpurun_out/r2last_bench_n1.json 2> gpurun_out/r2last_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2last_bench_reference.json 2> gpurun_out/r2last_bench_reference.err
cat $L; tail -c 300 gpurun_out/r2last_bench_n1.json; echo; tail -c 300 gpurun_out/r2last_bench_reference.json
