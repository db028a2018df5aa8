#!/usr/bin/env python
"""Kernel-level timing of integrate / raycast on a few orbit frames (tuning aid, not the bench contract)."""
import argparse, os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsdf_b200 import scenes, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--frames", default="0,125,250,375,500,625,750,875")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--prefill", type=int, default=8, help="orbit frames fused before timing (evenly spaced)")
args = ap.parse_args()
n = (args.size,) * 3
eng = sharded.ShardedEngine(n, (3000.0,) * 3)
W, H = 640, 480
for i in range(args.prefill):
    cam = scenes.orbit_camera(i * 1000 // max(args.prefill, 1), 1000)
    eng.integrate(torch.from_numpy(scenes.render_depth(cam)).cuda(), cam)
torch.cuda.synchronize()
peak = 6552.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
fracs = []
for f in [int(x) for x in args.frames.split(",")]:
    cam = scenes.orbit_camera(f, 1000)
    d = torch.from_numpy(scenes.render_depth(cam)).cuda()
    nu = eng.integrate(d, cam, count=True)
    ti, tr = [], []
    for r in range(args.reps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda._sleep(400000)      # keep the GPU busy while the launches queue up: the events then bracket device time only
        e[0].record(); eng.integrate(d, cam, restage=False); e[1].record(); eng.raycast(W, H, cam); e[2].record()
        torch.cuda.synchronize()
        ti.append(e[0].elapsed_time(e[1])); tr.append(e[1].elapsed_time(e[2]))
    ns = eng.raycast(W, H, cam, count=True)
    st = eng.last_ray_stats()
    ti, tr = float(np.median(ti)), float(np.median(tr))
    gbs = (16.0 * nu + W * H * 2) / (ti * 1e-3) / 1e9
    print(f"frame {f:4d}: upd {nu/ (args.size**3):5.1%}  integrate {ti*1e3:7.1f} us  {gbs:7.1f} GB/s ({gbs/peak:5.1%})   "
          f"raycast+normals {tr*1e3:7.1f} us  samples {ns:10d}  hits {st['hit_pixels']}", flush=True)
    fracs.append(gbs / peak)
print(f"integrate roofline fraction over these frames: min {min(fracs):.3f} mean {sum(fracs)/len(fracs):.3f}")
