#!/bin/bash
# round 2, third session, two GPUs: the GPU suite (incl. the two-rank NCCL and group tests) and the scaling line at N = 2 with the
# final build (streamed pageable downloads on rank 0, k_density epilogue, k_functional_occ<4>)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3n2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3n2_pytest.log; tail -3 gpurun_out/r3n2_pytest.log
sed 's/r2_scale_n/r3_scale_n/' tools/scaling_run_r2.sh > /tmp/scaling_run_r3.sh
bash /tmp/scaling_run_r3.sh 2
