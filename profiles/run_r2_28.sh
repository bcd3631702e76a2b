#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_runtime.py tests/test_modules_gpu.py tests/test_golden_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
bash profiles/run_r2_27.sh
