#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e_n2.json 2> gpurun_out/bench_e2e_n2.err; echo e2e rc=$?; tail -2 gpurun_out/bench_e2e_n2.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e_n1.json 2> gpurun_out/bench_e2e_n1.err; echo e2e1 rc=$?
timeout 600 python bench.py --workload ops --steps 10 --warmup 3 > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err; echo ops rc=$?; tail -2 gpurun_out/bench_ops.err
timeout 600 python bench.py --workload affinity-sharded --steps 20 --warmup 5 > gpurun_out/bench_aff_n1.json 2> gpurun_out/bench_aff_n1.err; echo aff rc=$?
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --rois proposal > gpurun_out/bench_prop.json 2> gpurun_out/bench_prop.err; echo prop rc=$?
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/bench_nopipe.json 2> gpurun_out/bench_nopipe.err; echo nopipe rc=$?
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo nograph rc=$?
python - <<'PY'
import json
for f in ("bench_e2e_n1","bench_e2e_n2","bench_ops","bench_aff_n1","bench_prop","bench_nopipe","bench_nograph"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), (d.get("collective") or {}).get("median_us"), d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
