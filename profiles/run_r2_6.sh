set -x
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_table_r2_d.txt > gpurun_out/bench_r2_d.json 2> gpurun_out/bench_r2_d.err; echo bench rc=$?; tail -3 gpurun_out/bench_r2_d.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pipeline --dump-launches gpurun_out/launch_table_r2_d_nopipe.txt > gpurun_out/bench_r2_d_nopipe.json 2> gpurun_out/bench_r2_d_nopipe.err; echo bench rc=$?
python - <<'PY'
import json
for f in ("bench_r2_d","bench_r2_d_nopipe"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stage_ms_per_call"), d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["all_tensor_kernels"])
    except Exception as e:
        print(f, "failed", e)
PY
sort -rn gpurun_out/launch_table_r2_d_nopipe.txt | head -40
