#!/usr/bin/env python
"""One orbit frame under a profiler: prefill the 512^3 volume, then integrate / raycast the given frame a few times.
  ncu --set full -k regex:integrate_rigid -s 2 -c 1 -o gpurun_out/prof python tools/prof_frame.py --frame 500 --what integrate"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsdf_b200 import scenes, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--frame", type=int, default=500)
ap.add_argument("--what", default="integrate")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--prefill", type=int, default=8)
args = ap.parse_args()
eng = sharded.ShardedEngine((args.size,) * 3, (3000.0,) * 3)
for i in range(args.prefill):
    cam = scenes.orbit_camera(i * 1000 // max(args.prefill, 1), 1000)
    eng.integrate(torch.from_numpy(scenes.render_depth(cam)).cuda(), cam)
cam = scenes.orbit_camera(args.frame, 1000)
d = torch.from_numpy(scenes.render_depth(cam)).cuda()
eng.integrate(d, cam)
torch.cuda.synchronize()
for r in range(args.reps):
    if args.what in ("integrate", "both"):
        eng.integrate(d, cam, restage=False)
    if args.what in ("raycast", "both"):
        eng.raycast(640, 480, cam)
torch.cuda.synchronize()
