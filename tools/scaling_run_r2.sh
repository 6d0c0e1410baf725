#!/bin/bash
# usage (under gpurun --gpus N): tools/scaling_run_r2.sh N [extra bench args]  -> gpurun_out/r2_scale_n<N>.json (one bench line: the
# tetracene headline + water64 + peptide at N ranks, parity of every workload against the oracle on rank 0)
n=$1; shift
mkdir -p gpurun_out
out=gpurun_out/r2_scale_n$n.json
if [ "$n" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --workloads water64,peptide --no-cpu-baseline "$@" > $out 2> ${out%.json}.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530 + n)) \
    bench.py --gpus $n --steps 20 --warmup 3 --workloads water64,peptide --no-cpu-baseline "$@" > $out 2> ${out%.json}.err
fi
python - "$out" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        for w in [d] + d.get("workloads", []):
            print(w.get("name", d["config"]["name"]), "n=%d" % d["n_gpus"], "ms=%.3f" % w["ms_per_step"], "e2e=%.3f" % w["e2e"]["ms_per_step"],
                  {k: round(v, 3) for k, v in w["kernels_ms_per_build"].items()}, (w.get("parity") or {}).get("within"),
                  (w.get("rank_balance") or {}).get("max_over_mean"))
PY
tail -3 ${out%.json}.err | cut -c1-300
