#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches.csv python profiles/decode_only.py 3 > gpurun_out/decode_only.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dg_decode -s 1 -c 1 -o gpurun_out/dg_decode python profiles/decode_only.py 2 >> gpurun_out/decode_only.log 2>&1; echo "rc=$?"
grep -v "^==" gpurun_out/decode_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:60], r['Metric Value'])
" | tail -12
