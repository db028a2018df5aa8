import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tsdf_b200 import scenes, sharded
n = (512,) * 3
W, H = 640, 480
world = 2
ranks = [sharded.ShardedEngine(n, (3000.0,) * 3, rank=r, world=world) for r in range(world)]
for i in range(6):
    cam = scenes.orbit_camera(i * 1000 // 6, 1000)
    d = torch.from_numpy(scenes.render_depth(cam)).cuda()
    for e in ranks: e.integrate(d, cam)
for f in (0, 125):
    cam = scenes.orbit_camera(f, 1000)
    d = torch.from_numpy(scenes.render_depth(cam)).cuda()
    for e in ranks:
        e.integrate(d, cam)
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); e.march(W, H, cam); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
        print(f"frame {f} rank {e.rank}: march {np.median(ts):.1f} us", flush=True)
