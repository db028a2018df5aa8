import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from tsdf_b200 import scenes, sharded
eng = sharded.ShardedEngine((512,)*3, (3000.0,)*3)
nb = 64**3
for f in [5, 24, 125, 250, 375, 500, 625, 750, 875]:
    cam = scenes.orbit_camera(f, 1000)
    d = torch.from_numpy(scenes.render_depth(cam)).cuda()
    eng.integrate(d, cam)
    eng.raycast(640, 480, cam)
    torch.cuda.synchronize()
    words = eng.occ[2*nb:2*nb+16].cpu().numpy().view(np.uint32)
    print(f"frame {f}: tile_counter {words[0]} queued {words[1]} claimed {words[2]} tiles_finished {words[3]}")
