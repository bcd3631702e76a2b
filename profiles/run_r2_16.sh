set -x
timeout 600 python -m pytest --timeout=120 tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/pytest_gpu.log
