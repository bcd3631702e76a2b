set -x
mkdir -p gpurun_out/ev
# launch list of exactly one pipelined step (cold-cache, serialised: compare SHARES)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ev/launches_step.csv python profiles/stage_profile.py step > gpurun_out/ev/launches_step.log 2>&1; echo rc=$?
# one full-set capture per kernel family (the first launch of each inside one step / one affinity stage)
for k in sa_fused_kernel ball_query_kernel three_nn_kernel fps_cluster_kernel fps_kernel roipool3d_kernel nms_greedy_batched_kernel proposal_select_kernel feature_gather_kernel three_interpolate_kernel rcnn_input; do
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/ev/prof_$k python profiles/stage_profile.py step > gpurun_out/ev/ncu_$k.log 2>&1; echo $k rc=$?
done
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm_kernel -c 1 -f -o gpurun_out/ev/prof_tc_gemm_link python profiles/stage_profile.py affinity > gpurun_out/ev/ncu_tc_gemm.log 2>&1; echo tc rc=$?
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pair_corr_kernel -c 1 -f -o gpurun_out/ev/prof_pair_corr python profiles/stage_profile.py affinity > gpurun_out/ev/ncu_pair_corr.log 2>&1; echo pc rc=$?
ls -la gpurun_out/ev | head -40
