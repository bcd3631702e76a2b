set -x
for cfg in 0 6 1; do
JMB_FPS_CFG=$cfg timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fps$cfg.json 2> gpurun_out/bench_fps$cfg.err; echo rc=$?
done
python - <<'PY'
import json
for c in (0,6,1):
    d = json.loads(open(f"gpurun_out/bench_fps{c}.json").read().strip().splitlines()[-1])
    print(c, d["value"], d["ms_per_step"], d.get("stage_ms_per_call"))
PY
