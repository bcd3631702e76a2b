#!/bin/bash
set -x
mkdir -p gpurun_out
run() {
JMB_FPS_CFG=$1 JMB_FPS_CFG4K=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fps_$1_$2.json 2> gpurun_out/bench_fps_$1_$2.err; echo "cfg $1 $2 rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_fps_$1_$2.json").read().strip().splitlines()[-1])
    print("cfg $1 $2", d["value"], d["ms_per_step"], d.get("stage_ms_per_call"), d["roofline"]["frac"])
except Exception as e:
    print("cfg $1 $2 ERR", e)
PY
}
run 0 0
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_runtime.py tests/test_golden_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python profiles/ref_cuda_timing.py > gpurun_out/ref_cuda_timing.log 2>&1; grep -i "furthest" gpurun_out/ref_cuda_timing.log | head -3
