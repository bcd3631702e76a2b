#!/bin/bash
set -x
mkdir -p gpurun_out
for d in 0 1 2; do
JMB_DG_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches_$d.csv python profiles/decode_only.py 2 > gpurun_out/decode_only.log 2>&1; echo "rc=$?"
grep "dg_decode" gpurun_out/decode_launches_$d.csv | tail -1 | awk -F'","' '{print $NF}'
done
