#!/usr/bin/env python
"""Distribution of raycast loop iterations per ray (run with TSDF_B200_DEBUG_ITERS=1): who sets the kernel's tail?"""
import os, sys
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from tsdf_b200 import scenes
import gpu_util as G

n = (512,) * 3
dv = G.DeviceVolume(n, (3000.0,) * 3)
for i in range(8):
    cam = scenes.orbit_camera(i * 125, 1000)
    dv.integrate(scenes.render_depth(cam), cam.inv_pose, cam.k, cam.kinv, count=False)
for f in (0, 60, 125, 500):
    cam = scenes.orbit_camera(f, 1000)
    V, N, it, ns = dv.raycast(640, 480, cam.pose, cam.kinv, fastdiv=True)
    it = it.reshape(480, 640)
    hit = ~np.isnan(V[:, 0]).reshape(480, 640)
    q = np.percentile(it, [50, 90, 99, 99.9, 100])
    tiles = it.reshape(120, 4, 80, 8).max(axis=(1, 3))
    print(f"frame {f}: iterations/ray median {q[0]:.0f} p90 {q[1]:.0f} p99 {q[2]:.0f} p99.9 {q[3]:.0f} max {q[4]:.0f}; "
          f"sum {it.sum()/1e6:.1f}M; warp-tile max: mean {tiles.mean():.0f} p99 {np.percentile(tiles, 99):.0f} max {tiles.max()}; "
          f"hits {hit.sum()}; worst ray at {np.unravel_index(it.argmax(), it.shape)} hit={hit.flat[it.argmax()]}", flush=True)
