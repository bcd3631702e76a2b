#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e_n2.json 2> gpurun_out/bench_e2e_n2.err; echo e2e rc=$?; tail -3 gpurun_out/bench_e2e_n2.err
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo e2e1 rc=$?
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --serial-e2e > gpurun_out/bench_serial.json 2> gpurun_out/bench_serial.err; echo serial rc=$?
python - <<'PY'
import json
for f in ("bench_final","bench_e2e_n2","bench_serial"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), (d.get("collective") or {}).get("median_us"), d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
