#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 4 -c 1 -f -o gpurun_out/prof_tc_gemm_conv python profiles/conv_only.py > gpurun_out/conv_only.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/conv_only.log
