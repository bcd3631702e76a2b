set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -x > gpurun_out/pytest_parallel_gpu.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_parallel_gpu.log
for N in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload affinity-sharded --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_aff_n$N.json 2> gpurun_out/bench_r2_aff_n$N.err; echo aff N=$N rc=$?
done
timeout 600 python bench.py --gpus 1 --workload affinity-sharded --steps 20 --warmup 5 > gpurun_out/bench_r2_aff_n1.json 2> gpurun_out/bench_r2_aff_n1.err
for N in 2 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_e2e_n$N.json 2> gpurun_out/bench_r2_e2e_n$N.err; echo e2e N=$N rc=$?; tail -2 gpurun_out/bench_r2_e2e_n$N.err
done
python - <<'PY'
import json
for f in ("bench_r2_aff_n1","bench_r2_aff_n2","bench_r2_aff_n4","bench_r2_e2e_n2","bench_r2_e2e_n4"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("collective"))
    except Exception as e:
        print(f, "failed", e)
PY
