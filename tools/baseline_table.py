#!/usr/bin/env python
"""every baseline of section 2 on this box, per BASELINE.json config (1: 128^3 x 10 fixed pose; 2: 256^3
orbit; 3: 512^3 orbit) — CPU restatement (1 thread / all cores), reference CUDA as shipped (-G) and -O3 (oracle/_ref), and the
new kernels through the level-2 C-ABI with pageable host buffers (the reference's own calling convention).  JSON to stdout.
Frames are bounded (the reference's raycast takes ~0.7 s per 512^3 frame); medians after one warm-up frame."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tsdf_b200 import Volume, scenes
from oracle import oracle, refcuda

W, H, PHYS = 640, 480, (3000.0,) * 3


def cams_for(cfg, n):
    if cfg == 1:
        return [scenes.fixed_pose_camera() for _ in range(n)]
    total = 200 if cfg == 2 else 1000
    return [scenes.orbit_camera(i, total) for i in range(n)]


def timed(fn, frames):
    ti, tr = [], []
    for j, (cam, depth) in enumerate(frames):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn[0](cam, depth); torch.cuda.synchronize(); t1 = time.perf_counter()
        fn[1](cam); torch.cuda.synchronize(); t2 = time.perf_counter()
        if j > 0:
            ti.append(t1 - t0); tr.append(t2 - t1)
    return {"integrate_ms": 1e3 * float(np.median(ti)), "raycast_ms": 1e3 * float(np.median(tr)),
            "frames_per_s": 1.0 / (float(np.median(ti)) + float(np.median(tr))), "frames_timed": len(ti)}


out = {"host": {"nproc": os.cpu_count()}, "gpu": torch.cuda.get_device_name(0), "configs": {}}
for cfg, size, n_gpu, n_ref, n_cpu in ((1, 128, 10, 10, 4), (2, 256, 12, 6, 3), (3, 512, 12, 4, 0)):
    n = (size,) * 3
    cams = cams_for(cfg, max(n_gpu, n_ref, n_cpu))
    frames = [(c, scenes.render_depth(c, W, H)) for c in cams]
    row = {"volume": f"{size}^3", "pose": "fixed" if cfg == 1 else "orbit"}
    vol = Volume(n, PHYS)
    V, N = np.empty((H * W, 3), np.float32), np.empty((H * W, 3), np.float32)
    row["tsdf_b200_level2_pageable"] = timed((lambda c, d: vol.integrate(d, c.inv_pose, c.k, c.kinv),
                                              lambda c: vol.raycast(W, H, c.pose, c.kinv, V, N)), frames[:n_gpu])
    row["voxels_rewritten_last_frame"], row["samples_evaluated_last_frame"] = vol.stats()
    row["hit_pixels_last_frame"] = int((~np.isnan(V[:, 0])).sum())
    vol.close()
    for tag, label in (("O3", "ref_cuda_O3"), ("G", "ref_cuda_G_as_shipped")):
        if not refcuda.available(tag):
            continue
        lib = refcuda.RefLib(tag)
        with refcuda.quiet():
            rv = refcuda.RefVolume(lib, n, PHYS)
            row[label] = timed((lambda c, d: rv.integrate(d, c.k, c.pose), lambda c: rv.raycast(W, H, c.k, c.pose)), frames[:n_ref])
            rv.close()
    if n_cpu:
        for threads, label in ((1, "cpu_restatement_1_thread"), (os.cpu_count(), "cpu_restatement_all_cores")):
            cores = oracle.set_threads(threads)
            ov = oracle.OracleVolume(n, PHYS)
            r = timed((lambda c, d: ov.integrate(d, c.inv_pose, c.k, c.kinv), lambda c: ov.raycast(W, H, c.pose, c.kinv, want_khit=False)),
                      frames[:n_cpu])
            r["threads"] = cores
            row[label] = r
    out["configs"][str(cfg)] = row
print(json.dumps(out))
