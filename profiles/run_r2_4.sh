set -x
JMB_SA_LOCKSTEP=1 timeout 300 python profiles/sa_bench.py > gpurun_out/sa_bench_lockstep.txt 2>&1; cat gpurun_out/sa_bench_lockstep.txt
JMB_SA_LOCKSTEP=1 JMB_SA_DEBUG=1 timeout 300 python profiles/sa_bench.py > /dev/null 2> gpurun_out/sa_timeline_lockstep.txt
head -c 2500 gpurun_out/sa_timeline_lockstep.txt
