set -x
timeout 600 python -m pytest --timeout=120 tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/pytest_gpu.log
timeout 200 python profiles/ref_cuda_timing.py gpurun_out/ref_cuda_timing.json > gpurun_out/ref_timing.log 2>&1; grep "ball_query\|three_nn" gpurun_out/ref_timing.log
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/launch_table_r2_i.txt > gpurun_out/bench_r2_i.json 2> gpurun_out/bench_r2_i.err; echo bench rc=$?
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_i.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stage_ms_per_call"), d["roofline"]["achieved"], d["roofline"]["frac"])
PY
