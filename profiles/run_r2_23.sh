#!/bin/bash
# round 2 HEAD: ncu launch list of exactly one pipelined step (shares), with full kernel names for the elementwise kernels
set -x
mkdir -p gpurun_out/ev
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ev/launches_step_head.csv python profiles/stage_profile.py step > gpurun_out/ev/launches_step_head.log 2>&1; echo rc=$?
python profiles/launch_table.py gpurun_out/ev/launches_step_head.csv 60 > gpurun_out/ev/launches_step_head.txt
head -45 gpurun_out/ev/launches_step_head.txt
