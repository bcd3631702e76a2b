#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pipeline-depth 3 > gpurun_out/bench_r2_p3.json 2> gpurun_out/bench_r2_p3.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_r2_p3.err
python - <<'PY'
import json
for n in ("bench_r2_p3",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stage_ms_per_call"), d["roofline"]["frac"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
