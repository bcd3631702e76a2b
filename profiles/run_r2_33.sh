#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python profiles/sa_bench.py > gpurun_out/sa_bench.txt 2>&1; cat gpurun_out/sa_bench.txt
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_modules_gpu.py tests/test_golden_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["ms_per_step"], d.get("stage_ms_per_call"), d["roofline"]["frac"], d["roofline"]["achieved"])
PY
