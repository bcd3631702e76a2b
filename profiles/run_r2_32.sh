#!/bin/bash
set -x
mkdir -p gpurun_out
run() {
env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "$1 rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
    print("cfg $1", d["value"], d["ms_per_step"], d.get("stage_ms_per_call"), d["roofline"]["frac"])
except Exception as e:
    print("cfg $1 ERR", e)
PY
}
run JMB_DECODE_SPAWN_LEVEL=0
run JMB_DECODE_SPAWN_LEVEL=1
run JMB_DECODE_SPAWN_LEVEL=2
run JMB_DECODE_SPAWN_LEVEL=3
