#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_table_o.txt > gpurun_out/bench_r2_o3.json 2> gpurun_out/bench_r2_o3.err; echo "bench rc=$?"
head -60 gpurun_out/launch_table_o.txt
