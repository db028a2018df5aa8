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
def classes(cam):
    """Instrumented build only (make dbg; TSDF_B200_LIB=tsdf_b200/libtsdf_b200_dbg.so): per-ray iteration classes, tile clocks."""
    out = {}
    for mode, name in ((1, "iters"), (2, "eval"), (3, "l1"), (4, "l2"), (7, "l3"), (5, "tile_ns"), (6, "tile_t0")):
        os.environ["TSDF_B200_DEBUG_ITERS"] = str(mode)
        out[name] = dv.raycast(640, 480, cam.pose, cam.kinv, fastdiv=True)[2].reshape(480, 640).astype(np.int64)
    os.environ["TSDF_B200_DEBUG_ITERS"] = "1"
    return out


for f in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,125,250,375,500,625".split(","))]:
    cam = scenes.orbit_camera(f, 1000)
    if "dbg" in os.environ.get("TSDF_B200_LIB", ""):
        c = classes(cam)
        w = np.unravel_index(c["iters"].argmax(), c["iters"].shape)
        tns = c["tile_ns"].reshape(120, 4, 80, 8).max(axis=(1, 3))
        tit = c["iters"].reshape(120, 4, 80, 8).max(axis=(1, 3))
        t0 = c["tile_t0"].reshape(120, 4, 80, 8).max(axis=(1, 3))
        t0 = t0 - t0.min()
        order = np.argsort(tns.ravel())[::-1][:8]
        print(f"frame {f}: worst ray {w}: iters {c['iters'][w]} = eval {c['eval'][w]} + l1 {c['l1'][w]} + l2 {c['l2'][w]} (l3 skips {c['l3'][w]}); "
              f"all rays: eval {c['eval'].sum()/1e6:.2f}M l1 {c['l1'].sum()/1e6:.2f}M l2 {c['l2'].sum()/1e6:.2f}M")
        print("   slowest tiles (us, max iters, start us): " + ", ".join(
            f"{tns.ravel()[i]/1e3:.0f}/{tit.ravel()[i]}/{t0.ravel()[i]/1e3:.0f}" for i in order) +
            f"; last tile end {((t0 + tns).max())/1e3:.0f} us; ns per iteration of slowest tiles {np.mean([tns.ravel()[i]/max(tit.ravel()[i],1) for i in order]):.0f}")
        # per-class iteration counts of the worst 1% of rays
        thr = np.percentile(c["iters"], 99)
        m = c["iters"] >= thr
        print(f"   top 1% rays (>= {thr:.0f} iters): mean eval {c['eval'][m].mean():.0f} l1 {c['l1'][m].mean():.0f} l2 {c['l2'][m].mean():.0f}")
    V, N, it, ns = dv.raycast(640, 480, cam.pose, cam.kinv, fastdiv=True)
    it = it.reshape(480, 640)
    hit = ~np.isnan(V[:, 0]).reshape(480, 640)
    q = np.percentile(it, [50, 90, 99, 99.9, 100])
    tiles = it.reshape(120, 4, 80, 8).max(axis=(1, 3))
    print(f"frame {f}: iterations/ray median {q[0]:.0f} p90 {q[1]:.0f} p99 {q[2]:.0f} p99.9 {q[3]:.0f} max {q[4]:.0f}; "
          f"sum {it.sum()/1e6:.1f}M; warp-tile max: mean {tiles.mean():.0f} p99 {np.percentile(tiles, 99):.0f} max {tiles.max()}; "
          f"hits {hit.sum()}; worst ray at {np.unravel_index(it.argmax(), it.shape)} hit={hit.flat[it.argmax()]}", flush=True)
