#!/bin/bash
# round 2: sparse image decoder — tests, timing next to the dense cuDNN formulation
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_decode.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_decode.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_decode.log
timeout 600 python profiles/decode_bench.py > gpurun_out/decode_bench.log 2>&1; echo "decode bench rc=$?"
tail -20 gpurun_out/decode_bench.log
