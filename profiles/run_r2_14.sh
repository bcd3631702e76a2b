set -x
timeout 200 python profiles/tc_bench.py > gpurun_out/tc_bench_r2.txt 2>&1; tail -45 gpurun_out/tc_bench_r2.txt
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pipeline --dump-launches gpurun_out/launch_table_r2_g_nopipe.txt > gpurun_out/bench_r2_g_nopipe.json 2> gpurun_out/bench_r2_g_nopipe.err; echo rc=$?
sort -rn gpurun_out/launch_table_r2_g_nopipe.txt | head -50
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_g_nopipe.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("stage_ms_per_call"))
PY
