set -x
nvidia-smi -L | wc -l
timeout 600 python -m pytest --timeout=300 tests/test_parallel_gpu.py -m gpu -q > gpurun_out/pytest_parallel_gpu.log 2>&1; echo pytest rc=$?
tail -6 gpurun_out/pytest_parallel_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload affinity-sharded --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_aff_n2.json 2> gpurun_out/bench_aff_n2.err; echo aff rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_e2e_n2.json 2> gpurun_out/bench_e2e_n2.err; echo e2e rc=$?; tail -2 gpurun_out/bench_e2e_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo ref rc=$?
python - <<'PY'
import json
for f in ("bench_aff_n2","bench_e2e_n2","bench_ref_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("collective"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
