"""Tuning aid: times the cluster-FPS configurations selectable through JMB_FPS_CFG (see csrc/sampling.cu)."""
import os, subprocess, sys
if len(sys.argv) > 1:
    import numpy as np, torch
    sys.path.insert(0, '.')
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    from jmodt_b200 import synth
    from oracle import cref
    b, n, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    pts = synth.make_batch(0, b, n_points=n, with_image=False)["pts"]
    x = torch.from_numpy(pts).cuda()
    idx = pu.farthest_point_sample(x, m)
    ok = np.array_equal(idx[:1].cpu().numpy(), cref.fps(pts[:1], m))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): pu.farthest_point_sample(x, m)
    e1.record(); torch.cuda.synchronize()
    print(f"cfg={os.environ.get('JMB_FPS_CFG','0')} b={b} n={n} m={m}: {e0.elapsed_time(e1)/5:.3f} ms  exact={ok}")
else:
    for (b, n, m) in ([(8, 4096, 1024)] if os.environ.get('FPS_TUNE_L1') else [(8, 16384, 4096), (8, 4096, 1024), (1, 16384, 4096)]):
        for cfg in (range(0, 10) if n == 4096 else range(0, 8)):
            env = dict(os.environ, JMB_FPS_CFG=str(cfg))
            r = subprocess.run([sys.executable, __file__, str(b), str(n), str(m)], env=env, capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr.strip()[-300:])
