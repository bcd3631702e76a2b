set -x
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_head_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python profiles/sa_bench.py > gpurun_out/sa_bench.txt 2>&1; cat gpurun_out/sa_bench.txt
JMB_SA_DEBUG=1 timeout 300 python profiles/sa_bench.py > /dev/null 2> gpurun_out/sa_timeline.txt
head -c 1800 gpurun_out/sa_timeline.txt
