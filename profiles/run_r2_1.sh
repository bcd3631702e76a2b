set -x
mkdir -p gpurun_out
nvidia-smi -L
python tests/golden/make_golden_gpu.py gpurun_out/ref_gpu.npz > gpurun_out/golden.log 2>&1; echo golden rc=$?
cp gpurun_out/ref_gpu.npz tests/golden/ 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python profiles/ref_cuda_timing.py gpurun_out/ref_cuda_timing.json > gpurun_out/ref_timing.log 2>&1; echo timing rc=$?
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; echo bench rc=$?
tail -5 gpurun_out/golden.log
