set -x
timeout 300 python -m pytest --timeout=60 tests/test_tc_gpu.py tests/test_head_gpu.py tests/test_modules_gpu.py tests/test_runtime.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/pytest_gpu.log
timeout 120 python profiles/sa_bench.py > gpurun_out/sa_bench.txt 2>&1; cat gpurun_out/sa_bench.txt
JMB_SA_DEBUG=1 timeout 60 python profiles/sa_bench.py > /dev/null 2> gpurun_out/sa_timeline.txt
head -c 1500 gpurun_out/sa_timeline.txt
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_table_r2_f.txt > gpurun_out/bench_r2_f.json 2> gpurun_out/bench_r2_f.err; echo bench rc=$?; tail -3 gpurun_out/bench_r2_f.err
python - <<'PY'
import json
for f in ("bench_r2_f",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stage_ms_per_call"), d["roofline"]["achieved"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
