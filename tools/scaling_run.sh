#!/bin/bash
# usage (on the GPU box, under gpurun --gpus 8): tools/scaling_run.sh "<n list>" <workload> <steps>
# runs bench.py under torchrun for each n and prints a one-line summary per run; JSON lines go to gpurun_out/
mkdir -p gpurun_out
for n in $1; do
  out=gpurun_out/bench_r1_$2_n$n.json
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps $3 --warmup 3 --workload $2 --no-cpu-baseline > $out 2> ${out%.json}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530 + n)) \
      bench.py --gpus $n --steps $3 --warmup 3 --workload $2 > $out 2> ${out%.json}.err
  fi
  python - "$out" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["config"]["name"], "n=%d" % d["n_gpus"], "ms=%.3f" % d["ms_per_step"], "e2e=%.3f" % d["e2e"]["ms_per_step"],
              "Mpts/s=%.1f" % (d["value"] / 1e6), d["config"]["shard_points"], d["result"])
PY
done
