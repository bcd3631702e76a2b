#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_image_decode.py -m gpu -q -x --timeout 120 -k "basic_block" > gpurun_out/pytest_conv.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_conv.log
timeout 600 python profiles/decode_bench.py > gpurun_out/decode_bench.log 2>&1; echo "decode bench rc=$?"
tail -24 gpurun_out/decode_bench.log
