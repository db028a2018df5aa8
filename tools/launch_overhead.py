#!/usr/bin/env python
"""How long does the host take to ISSUE one bench step (ctypes marshalling of ~8 C-ABI calls) compared with the GPU time
of the step?  If the two are close the bench is launch-bound, not kernel-bound."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsdf_b200 import scenes, sharded

n = (512,) * 3
eng = sharded.ShardedEngine(n, (3000.0,) * 3)
cams = [scenes.orbit_camera(i, 1000) for i in range(60)]
frames = [torch.from_numpy(scenes.render_depth(c)).cuda() for c in cams]
for i in range(10):
    eng.integrate(frames[i], cams[i]); eng.raycast(640, 480, cams[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(10, 60):
    eng.integrate(frames[i], cams[i]); eng.raycast(640, 480, cams[i])
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host issue time {1e3 * (t1 - t0) / 50:.3f} ms/step, device time {e0.elapsed_time(e1) / 50:.3f} ms/step")
