#!/usr/bin/env python
"""BASELINE configs[4] as one timed pipeline on this process's GPU(s): per frame bilateral filter (16-bit depth) ->
integrate -> raycast + normals; marching cubes of the fused volume at the end (per Z-shard when launched under torchrun).
Not the bench contract (bench.py is) — the numbers go to DESIGN.md section 2.4.

  python tools/pipeline_bench.py [--size 512] [--frames 100]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pipeline_bench.py
"""
import argparse, ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tsdf_b200 import scenes, sharded
from tsdf_b200.capi import lib, check

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--frames", type=int, default=100)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--sigma-colour", type=float, default=30.0)
ap.add_argument("--sigma-space", type=float, default=2.0)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
W, H = 640, 480
n = (args.size,) * 3
eng = sharded.ShardedEngine(n, (3000.0,) * 3, rank, world, exchange="peer")
stream = eng.stream

# look-up tables of BilateralFilter's constructor (reference src/BilateralFilter.cpp:15-42), built on the host
radius = int(np.ceil(np.float32(args.sigma_space) * np.float32(1.5)))
ks = 2 * radius + 1
xs = np.arange(-radius, radius + 1, dtype=np.float32)
kern = np.exp(-(xs[:, None] ** 2 + xs[None, :] ** 2) / np.float32(args.sigma_space) ** 2).astype(np.float32).reshape(-1)
simi = np.exp(-np.arange(65536, dtype=np.float32) / np.float32(args.sigma_colour) ** 2).astype(np.float32)
d_kern, d_simi = torch.from_numpy(kern).cuda(), torch.from_numpy(simi).cuda()

total = args.warmup + args.frames
cams = [scenes.orbit_camera(i, 1000) for i in range(total)]
raw = [torch.from_numpy(scenes.render_depth(c, W, H)).cuda() for c in cams]
filt = torch.empty((H, W), dtype=torch.uint16, device="cuda")


def frame(i):
    check(lib.tsdf_b200_bilateral_u16(C.c_void_p(raw[i].data_ptr()), C.c_void_p(filt.data_ptr()), W, H, C.c_void_p(d_kern.data_ptr()),
                                      ks, C.c_void_p(d_simi.data_ptr()), simi.size, stream), "bilateral")
    eng.integrate(filt, cams[i])
    eng.raycast(W, H, cams[i])


for i in range(args.warmup):
    frame(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for i in range(args.warmup, total):
    frame(i)
e1.record()
mesh = eng.extract_mesh()          # first call: table upload, kernel loading
torch.cuda.synchronize()
del mesh                           # (the timed call then reuses the block in torch's caching allocator instead of a cudaMalloc)
e1b = torch.cuda.Event(enable_timing=True)
e1b.record()
mesh = eng.extract_mesh()
e2.record()
torch.cuda.synchronize()
t_frames, t_mc = e0.elapsed_time(e1), e1b.elapsed_time(e2)
tri = torch.tensor([mesh.shape[0] // 3], device="cuda", dtype=torch.int64)
t = torch.tensor([t_frames, t_mc], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(tri)
if rank == 0:
    print(json.dumps({"workload": f"{args.size}^3, {args.frames} frames of the sphere+wall orbit: bilateral(u16, {ks}x{ks}) + integrate + "
                                  f"raycast + normals per frame, marching cubes at the end", "n_gpus": world,
                      "frames_per_s": args.frames / (float(t[0]) * 1e-3), "ms_per_frame": float(t[0]) / args.frames,
                      "marching_cubes_ms": float(t[1]), "triangles": int(tri.item())}))
if world > 1:
    dist.destroy_process_group()
