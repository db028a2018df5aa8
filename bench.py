#!/usr/bin/env python
"""bench.py — depth frames/s of the TSDF hot path (integrate + raycast) at 512^3, 640x480.

One step = one depth frame of BASELINE.json configs[2] (512^3 volume, 3000 mm cube, 1000-frame orbit
of the analytic sphere + wall scene): integrate the frame, then raycast (+ normals) from its pose.

  value     frames/s with the depth frames already resident in HBM (level-1 C-ABI launches on one
            stream, CUDA events around the K steps).
  e2e       the same K frames through the level-2 C-ABI with HOST buffers — the call path of the
            reference's TSDFVolume::integrate / ::raycast: pinned-host depth H2D inside integrate,
            vertex + normal maps D2H inside raycast, every call synchronous.
  roofline  integrate kernel: algorithmic bytes (16 B x voxels rewritten + the depth frame) / CUDA-event
            time of the integrate launches, against the measured HBM peak.
  cpu_baseline / --impl reference: the CPU restatement of the reference kernels (oracle/, the reference
            has no CPU path of its own) on all host cores, on a bounded sample of the same frame.

Multi-GPU (torchrun, --gpus N): the volume is sharded along Z, see DESIGN.md.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 640, 480
PHYS = (3000.0, 3000.0, 3000.0)
ORBIT_FRAMES = 1000
METRIC = "depth frames/sec (integrate+raycast) at 512^3 vol, 640x480"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="voxels per side (512 = the headline workload)")
    ap.add_argument("--layout", default="contiguous", choices=["contiguous", "interleaved", "replica"],
                    help="multi-GPU layout (tsdf_b200/sharded.py): Z-slabs + key all-reduce, or surface replicas + image tiles")
    ap.add_argument("--slab", type=int, default=0, help="planes per slab for the interleaved / replica layouts (0: size / gpus)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def frame_inputs(i):
    from tsdf_b200 import scenes
    cam = scenes.orbit_camera(i % ORBIT_FRAMES, ORBIT_FRAMES)
    return cam, scenes.render_depth(cam, W, H)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md's clocks line, via NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
                "reasons": sorted(self.reasons)}


def cpu_sample(size, frame_index, rows_step, z_frac, threads=None):
    """Bounded sample of one frame on the host cores with the CPU restatement: integrate a 1/z_frac Z-slab and
    raycast every rows_step-th image row of the 512^3 frame; times are scaled back to a whole frame."""
    from oracle import oracle
    cores = oracle.set_threads(threads or (os.cpu_count() or 1))
    cam, depth = frame_inputs(frame_index)
    ov = getattr(cpu_sample, "_vol", None)
    if ov is None or ov.size != (size,) * 3:
        ov = oracle.OracleVolume((size,) * 3, PHYS)
        # give the raycast a surface to find: fuse this frame once into the whole volume (untimed)
        ov.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
        cpu_sample._vol = ov
    z0 = (size // 2) - (size // z_frac) // 2
    t0 = time.perf_counter()
    ov.integrate(depth, cam.inv_pose, cam.k, cam.kinv, z0, z0 + size // z_frac)
    t1 = time.perf_counter()
    ov.raycast(W, H, cam.pose, cam.kinv, want_khit=False, y_begin=rows_step // 2, y_step=rows_step, want_normals=False)
    t2 = time.perf_counter()
    frame_s = (t1 - t0) * z_frac + (t2 - t1) * rows_step
    return frame_s, cores, (t1 - t0) * z_frac, (t2 - t1) * rows_step


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores (CPU restatement, kind "port")."""
    if rank != 0:
        return
    size = args.size
    rows_step, z_frac = 16, 8
    for i in range(args.warmup):
        cpu_sample(size, i, rows_step, z_frac)
    times = []
    for i in range(args.steps):
        fs, cores, _, _ = cpu_sample(size, args.warmup + i, rows_step, z_frac)
        times.append(fs)
    ms = 1e3 * float(np.mean(times))
    val = 1e3 / ms
    sample = (f"per step: integrate a {size // z_frac}-plane Z-slab (1/{z_frac} of {size}^3) + raycast every {rows_step}th "
              f"image row of the same orbit frame; times scaled x{z_frac} / x{rows_step} to a whole frame")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{size}^3 volume, 640x480, {ORBIT_FRAMES}-frame orbit (BASELINE configs[2])",
                   "note": "reference has no CPU path; this is the line-by-line CPU restatement of its CUDA kernels"},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import ctypes as C
    from tsdf_b200 import capi, Volume
    from tsdf_b200.capi import lib, check, fptr, fvec, colmajor
    from tsdf_b200 import sharded

    size, K, Wm = args.size, args.steps, args.warmup
    n = (size, size, size)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    # ---- inputs: K+W orbit frames, rendered on the host, resident in HBM before timing ---------------
    cams, frames = [], []
    for i in range(Wm + K):
        cam, depth = frame_inputs(i)
        cams.append(cam)
        frames.append(depth)
    d_frames = [torch.from_numpy(f).cuda() for f in frames]

    slab = args.slab if args.slab > 0 else max(8, (size // max(world, 1)) // 8 * 8)
    eng = sharded.ShardedEngine(n, PHYS, rank, world, stream=stream.cuda_stream, layout=args.layout, slab=slab)

    def step(i, count=False):
        eng.integrate(d_frames[i], cams[i], count=count)
        eng.raycast(W, H, cams[i])

    for i in range(Wm):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize()
    ev0.record(stream)
    for s in range(K):
        i = Wm + s
        eng.stage(d_frames[i])                      # culling pyramid of the frame (2 small launches)
        iev[s][0].record(stream)                    # events bracket the integrate kernel alone (roofline)
        eng.integrate(d_frames[i], cams[i], count=False, restage=False)
        iev[s][1].record(stream)
        eng.raycast(W, H, cams[i])
    ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    t_int_ms = [a.elapsed_time(b) for a, b in iev]
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        dist.barrier()
    ms_per_step = total_ms / K
    value = 1e3 / ms_per_step

    # ---- untimed: voxels rewritten per timed frame (depends on geometry only) -> algorithmic bytes ----
    n_upd = [eng.integrate(d_frames[Wm + s], cams[Wm + s], count=True) for s in range(K)]
    b_alg = [16.0 * u + W * H * 2 for u in n_upd]
    ray_stats = eng.last_ray_stats()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = sum(b_alg) / (sum(t_int_ms) * 1e-3) / 1e9
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("size") == size:
            traffic = tr.get("integrate_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "integrate_rigid_kernel",
                "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback",
                "algorithmic_bytes_per_launch": float(np.mean(b_alg)),
                "integrate_ms_per_launch": float(np.mean(t_int_ms)),
                "voxels_rewritten_per_frame": float(np.mean(n_upd)), "per_gpu": world > 1}

    # ---- e2e: level-2 C-ABI with host buffers (rank 0's view; every rank does the same work) ----------
    e2e = None
    if not args.no_e2e and world == 1:
        eng.close()
        vol = Volume(n, PHYS)
        pin = [torch.from_numpy(f).pin_memory() for f in frames]
        pin_np = [p.numpy() for p in pin]
        hv = torch.empty((H * W, 3), dtype=torch.float32).pin_memory()
        hn = torch.empty((H * W, 3), dtype=torch.float32).pin_memory()
        hv_np, hn_np = hv.numpy(), hn.numpy()
        mats = [(colmajor(c.inv_pose), colmajor(c.k), colmajor(c.kinv), colmajor(c.pose)) for c in cams]
        # argument marshalling done once: a C or C++ caller of the C-ABI has none of it (ctypes pointer objects cost
        # microseconds each, during which the GPU would idle inside the timed region)
        args_i = [(C.c_void_p(pin_np[i].ctypes.data), fptr(m[0]), fptr(m[1]), fptr(m[2]), fptr(m[3])) for i, m in enumerate(mats)]
        hv_p, hn_p = C.c_void_p(hv_np.ctypes.data), C.c_void_p(hn_np.ctypes.data)
        f_int, f_ray, handle = lib.tsdf_b200_volume_integrate, lib.tsdf_b200_volume_raycast, vol._h

        def e2e_step(i):
            depth_p, ip, k, kinv, pose = args_i[i]
            check(f_int(handle, depth_p, W, H, ip, k, kinv))
            check(f_ray(handle, W, H, pose, kinv, hv_p, hn_p))

        for i in range(Wm):
            e2e_step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(K):
            e2e_step(Wm + s)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / K
        e2e = {"value": 1e3 / e2e_ms, "unit": "frames/s", "h2d_bytes_per_step": W * H * 2,
               "d2h_bytes_per_step": 2 * W * H * 3 * 4, "ms_per_step": e2e_ms,
               "api": "tsdf_b200_volume_integrate + tsdf_b200_volume_raycast (host buffers, synchronous)"}
        vol.close()
    elif world > 1 and not args.no_e2e:
        e2e = eng.e2e(frames, cams, Wm, K, W, H)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_sample(size, Wm, 8, 4)                       # warm-up (allocates + fuses one frame)
        fs, cores, ti, tr_ = cpu_sample(size, Wm + 1, 8, 4)
        cpu_baseline = {"value": 1.0 / fs, "unit": "frames/s", "cores": cores, "kind": "port",
                        "sample": f"one orbit frame at {size}^3: integrate a {size // 4}-plane Z-slab (x4) + raycast every 8th "
                                  f"image row (x8) with the CPU restatement (OpenMP, all host cores); "
                                  f"integrate {ti * 1e3:.0f} ms + raycast {tr_ * 1e3:.0f} ms per whole frame"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{size}^3 volume / 3000 mm, 640x480 depth, {ORBIT_FRAMES}-frame orbit of sphere+wall "
                                   f"(BASELINE configs[2]); step = integrate + raycast + normals of one frame",
                       "cache": "volume (1 GiB dist+weight at 512^3) is larger than L2, no flush needed",
                       "parallelism": "single GPU" if world == 1 else
                                      (f"Z-slab sharding over {world} GPUs, key all-reduce(min)" if args.layout != "replica" else
                                       f"Z-slabs of {slab} planes over {world} GPUs, surface bricks pushed to per-GPU replicas over "
                                       f"NVLink, image tiles sharded")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": eng.launches_per_step * K,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "raycast": ray_stats,
        }
        print(json.dumps(out))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        eng.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
