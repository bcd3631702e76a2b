#!/bin/bash
set -x
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_o$i.json 2> gpurun_out/bench_r2_o$i.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for n in ("bench_r2_o1", "bench_r2_o2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stage_ms_per_call"), d["roofline"]["frac"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
