#!/bin/bash
# HEAD on 8 GPUs of one box: e2e at N = 8, 4, 2, 1 back to back
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e_n$n.json 2> gpurun_out/bench_e2e_n$n.err; echo e2e $n rc=$?
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e_n1.json 2> gpurun_out/bench_e2e_n1.err; echo e2e 1 rc=$?
python - <<'PY'
import json
for f in ("bench_e2e_n1","bench_e2e_n2","bench_e2e_n4","bench_e2e_n8"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), (d.get("collective") or {}).get("median_us"), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
