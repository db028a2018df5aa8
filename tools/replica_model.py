#!/usr/bin/env python
"""Cost model of the image-sharded multi-GPU path, measured on ONE GPU (tuning aid): the ranks of a `world`-GPU run are
emulated side by side, and each phase is timed per rank with CUDA events.  Peer stores go to local memory here, so the
push time is a lower bound; bricks pushed per rank show how evenly the NVLink traffic would spread."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsdf_b200 import scenes, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--configs", default="2:256,2:32,4:128,4:32,8:64,8:32,8:16", help="world:slab_planes,...")
ap.add_argument("--frames", default="0,125,250,500")
ap.add_argument("--prefill", type=int, default=6)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
n = (args.size,) * 3
W, H = 640, 480


def timed(fn):
    ts = []
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


for cfg in args.configs.split(","):
    world, slab = (int(x) for x in cfg.split(":"))
    ranks = [sharded.ShardedEngine(n, (3000.0,) * 3, rank=r, world=world, layout="replica", slab=slab) for r in range(world)]
    for e in ranks:
        e.connect(W, H, peers=ranks)
    for i in range(args.prefill):
        cam = scenes.orbit_camera(i * 1000 // max(args.prefill, 1), 1000)
        d = torch.from_numpy(scenes.render_depth(cam)).cuda()
        for e in ranks:
            e.integrate(d, cam)
    for f in [int(x) for x in args.frames.split(",")]:
        cam = scenes.orbit_camera(f, 1000)
        d = torch.from_numpy(scenes.render_depth(cam)).cuda()
        t_int = [timed(lambda: e.integrate(d, cam, restage=False)) for e in ranks]
        merged = ranks[0].flags().clone()
        for e in ranks[1:]:
            merged = torch.maximum(merged, e.flags())
        for e in ranks:
            e.flags().copy_(merged)
        bricks = [e.push(count=True) for e in ranks]
        t_push = [timed(lambda: e.push()) for e in ranks]
        t_ray = [timed(lambda: e.march_tiles(W, H, cam)) for e in ranks]
        print(f"world {world} slab {slab:3d} frame {f:4d}: flagged {int(merged.sum())} pushed {sum(bricks)} bricks "
              f"({sum(bricks) * 2048 / 1e6:.1f} MB), max/rank {max(bricks)} ({max(bricks) * 2048 * (world - 1) / 1e6:.1f} MB out)  "
              f"integrate max {max(t_int):6.1f} us  push(x{world} local) max {max(t_push):6.1f} us  "
              f"march tiles (incl. 3 distance passes) max {max(t_ray):6.1f} mean {np.mean(t_ray):6.1f} us", flush=True)
    for e in ranks:
        e.close()
    del ranks
    torch.cuda.empty_cache()
