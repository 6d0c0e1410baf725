#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "two_rank or cpp or bridge or communicator" > gpurun_out/r2_2gpu_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_2gpu_pytest.log
tail -4 gpurun_out/r2_2gpu_pytest.log
bash tools/scaling_run_r2.sh 2
cp gpurun_out/r2_scale_n2.json gpurun_out/r2_scale_n2_packed.json
