set -x
mkdir -p gpurun_out
python tests/golden/make_golden_gpu.py gpurun_out/ref_gpu.npz > gpurun_out/golden.log 2>&1; echo golden rc=$?
cp gpurun_out/ref_gpu.npz tests/golden/ 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python profiles/sa_bench.py > gpurun_out/sa_bench.txt 2>&1; cat gpurun_out/sa_bench.txt
JMB_SA_DEBUG=1 timeout 300 python profiles/sa_bench.py > /dev/null 2> gpurun_out/sa_timeline.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; echo bench rc=$?; tail -3 gpurun_out/bench_r2_b.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/bench_r2_b_nopipe.json 2> gpurun_out/bench_r2_b_nopipe.err; echo bench rc=$?
timeout 600 python bench.py --workload affinity-sharded --steps 10 --warmup 3 > gpurun_out/bench_r2_aff1.json 2> gpurun_out/bench_r2_aff1.err; echo bench rc=$?; tail -3 gpurun_out/bench_r2_aff1.err
python - <<'PY'
import json
for f in ("bench_r2_b", "bench_r2_b_nopipe", "bench_r2_aff1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stage_ms_per_call"), d["roofline"]["achieved"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
