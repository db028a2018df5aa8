import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tsdf_b200 import scenes, sharded
eng = sharded.ShardedEngine((512,) * 3, (3000.0,) * 3)
nb = 64 ** 3
prev = 0
changed = []
for i in range(210):
    cam = scenes.orbit_camera(i, 1000)
    d = torch.from_numpy(scenes.render_depth(cam)).cuda()
    eng.integrate(d, cam)
    c = int(eng.occ[:nb].sum().item())
    changed.append(c - prev)
    prev = c
ch = np.array(changed)
print("flagged after 210 frames:", prev, "frames with new bricks:", int((ch[5:] > 0).sum()), "of", len(ch) - 5, "new bricks per frame (5..):", ch[5:25].tolist(), "...", ch[-20:].tolist())
