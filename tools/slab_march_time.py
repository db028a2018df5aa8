#!/usr/bin/env python
"""Per-slab march time of the Z-sharded raycast, emulated on one GPU: the 512^3 volume after a few orbit frames is cut into
`world` slabs (own planes + halo), each slab is marched by tsdf_b200_raycast_slab from its own copy and occupancy grid, and the
call is timed with CUDA events.  The slowest slab is what a frame of the sharded raycast waits for."""
import argparse, os, sys
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsdf_b200 import scenes, sharded
from tsdf_b200.capi import lib, check, fptr, fvec, colmajor

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--frame", type=int, default=10)
ap.add_argument("--reps", type=int, default=7)
args = ap.parse_args()
n = (512,) * 3
eng = sharded.ShardedEngine(n, (3000.0,) * 3)
for f in range(args.frame + 1):
    cam = scenes.orbit_camera(f, 1000)
    eng.integrate(torch.from_numpy(scenes.render_depth(cam)).cuda(), cam)
torch.cuda.synchronize()
W, H = 640, 480
pose = np.asarray(cam.pose, np.float32)
smin = eng.offset.copy(); smax = (eng.offset + eng.physical).astype(np.float32)
full = eng.dist.view(n[2], n[1] * n[0])
p = lambda t: C.c_void_p(t.data_ptr())
out = []
for r, (z0, z1) in enumerate(sharded.shard_ranges(n[2], args.world)):
    zs1 = min(z1 + 1, n[2])
    slab = full[z0:zs1].contiguous().view(-1)
    occ = torch.zeros(lib.tsdf_b200_occupancy_bytes(n[0], n[1], zs1 - z0), dtype=torch.uint8, device="cuda")
    check(lib.tsdf_b200_occupancy_rebuild(p(slab), n[0], n[1], zs1 - z0, eng.trunc, p(occ), None))
    keys = torch.empty(H * W, dtype=torch.int64, device="cuda")
    ts = []
    for i in range(args.reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        e0.record()
        check(lib.tsdf_b200_raycast_slab(p(slab), *n, z0, zs1 - z0, z0, z1, fptr(eng.voxel), fptr(smin), fptr(smax), eng.trunc,
                                         fptr(fvec(pose[:3, 3])), fptr(colmajor(pose[:3, :3])), fptr(colmajor(cam.kinv)), W, H,
                                         p(eng.table), p(occ), p(keys), None, 1, None))
        e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    hits = int((keys != 0x7fffffffffffffff).sum().item())
    nb = ((n[0] + 7) // 8) * ((n[1] + 7) // 8) * ((zs1 - z0 + 7) // 8)
    words = occ[2 * nb:2 * nb + 32].cpu().numpy().view(np.uint32)
    out.append(np.median(ts))
    print(f"slab {r} planes [{z0},{z1}): {np.median(ts):7.1f} us  hits {hits:6d}  set aside {words[1]:6d}  cap {words[4]}", flush=True)
print(f"slowest slab {max(out):.1f} us, mean {np.mean(out):.1f} us")
