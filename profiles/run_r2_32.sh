#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_runtime.py -m gpu -q -x --timeout 300 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_x$i.json 2> gpurun_out/bench_x$i.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_x$i.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stage_ms_per_call"), d["roofline"]["frac"])
PY
done
