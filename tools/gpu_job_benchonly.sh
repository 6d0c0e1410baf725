#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2last_bench_n1.json 2> gpurun_out/r2last_bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2last_bench_n1.json"))
print(round(d["ms_per_step"], 3), d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
for w in d["workloads"]:
    print(w["name"], round(w["ms_per_step"], 3), w["roofline"]["kernel"], round(w["roofline"]["frac"], 3), w["roofline"].get("launches_per_build"), w["parity"].get("within"))
PY
