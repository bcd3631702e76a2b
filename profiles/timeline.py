"""Kernel timeline of ONE e2e step via torch.profiler (CUPTI): per-stream busy time, idle gaps on the main stream,
and the raw (stream, start_us, dur_us, name) list for offline reading.
usage: python profiles/timeline.py [frames] > gpurun_out/timeline.txt"""
import sys, json, os, time, torch
sys.path.insert(0, '.')
import bench
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
host_np = bench.make_inputs_e2e(0, B)
suite = bench.FusionE2E(dev, B, host_np)
host = {k: torch.from_numpy(v) for k, v in host_np.items() if k != 'img'}
d = bench.to_device_e2e(host, dev, torch, non_blocking=False)
for _ in range(3): suite.step(d)
torch.cuda.synchronize()
# host enqueue time of a step vs its device time
t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); suite.step(d); e1.record(); t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"# host enqueue {t_enq*1e3:.2f} ms, device {e0.elapsed_time(e1):.2f} ms")
GRAPH = len(sys.argv) > 2 and sys.argv[2] == 'graph'
if GRAPH:
    from jmodt_b200.runtime import CapturedPath
    cap = CapturedPath(suite.step, d, warmup=1)
    for _ in range(2): cap.replay()
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    cap.replay() if GRAPH else suite.step(d)
    torch.cuda.synchronize()
path = 'gpurun_out/trace.json'
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
ev.sort(key=lambda e: e['ts'])
t0 = ev[0]['ts']
print(f"# {len(ev)} device activities, span {(ev[-1]['ts']+ev[-1]['dur']-t0)/1e3:.2f} ms")
for e in ev:
    print(f"{e['args'].get('stream', -1):3d} {e['ts']-t0:10.1f} {e['dur']:9.1f} {e['name'][:90]}")
os.remove(path)
