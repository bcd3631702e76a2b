#!/bin/bash
# round 2: full GPU test pass + bench with the image decoder inside the step (default), dense-map and image-stack variants
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python profiles/ref_cuda_timing.py > gpurun_out/ref_cuda_timing.log 2>&1; echo "ref timing rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_q.json 2> gpurun_out/bench_r2_q.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --image-map dense > gpurun_out/bench_r2_q_dense.json 2> gpurun_out/bench_r2_q_dense.err; echo "bench dense rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --image-stack > gpurun_out/bench_r2_q_stack.json 2> gpurun_out/bench_r2_q_stack.err; echo "bench stack rc=$?"
tail -3 gpurun_out/bench_r2_q_stack.err
python - <<'PY'
import json
for n in ("bench_r2_q", "bench_r2_q_dense", "bench_r2_q_stack"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stage_ms"), d["roofline"]["achieved"], d["roofline"]["frac"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
