#!/bin/bash
# Development aid: time kernel variants (environment switches read by sxc_create) on one workload; one JSON line per variant.
#   tools/variant_sweep.sh <out.jsonl> <workload> "VAR=val VAR=val" "VAR=val" ...
out=$1; shift
wl=$1; shift
for v in "$@"; do
  line=$(env $v timeout 600 python bench.py --workload $wl --workloads none --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/sweep_last.err)
  if [ -z "$line" ]; then line="{\"failed\": true, \"stderr\": $(tail -c 400 gpurun_out/sweep_last.err | python -c 'import json,sys; print(json.dumps(sys.stdin.read()))')}"; fi
  echo "{\"variant\": \"$v\", \"workload\": \"$wl\", \"line\": $line}" >> $out
done
