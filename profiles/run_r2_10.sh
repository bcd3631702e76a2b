set -x
timeout 120 python profiles/sa_bench.py > gpurun_out/sa_bench.txt 2>&1; cat gpurun_out/sa_bench.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sa_fused_kernel -s 2 -c 1 -o gpurun_out/prof_sa0_v7 -f python profiles/sa_bench.py > gpurun_out/ncu_sa0.log 2>&1; echo ncu rc=$?; tail -3 gpurun_out/ncu_sa0.log
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_table_r2_f.txt > gpurun_out/bench_r2_f.json 2> gpurun_out/bench_r2_f.err; echo bench rc=$?
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_f.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stage_ms_per_call"), d["roofline"]["achieved"], d["roofline"]["frac"])
PY
