#!/usr/bin/env python
"""bench.py — depth frames/s of the TSDF hot path (integrate + raycast) at 512^3, 640x480.

One step = one depth frame of BASELINE.json configs[2] (512^3 volume, 3000 mm cube, 1000-frame orbit
of the analytic sphere + wall scene): integrate the frame, then raycast (+ normals) from its pose.

  value     frames/s with the depth frames already resident in HBM (level-1 C-ABI launches on one
            stream, CUDA events around the K steps).
  e2e       the same K frames through the level-2 C-ABI with HOST buffers — the call path of the
            reference's TSDFVolume::integrate / ::raycast: pinned-host depth H2D inside integrate,
            vertex + normal maps D2H inside raycast, every call synchronous.  e2e_pageable: the same
            with plain malloc'ed buffers (what kinfu's DepthImage / Eigen matrices are).
  roofline  integrate kernel: algorithmic bytes (16 B x voxels rewritten + the depth frame) / CUDA-event
            time of the integrate launches, against the measured HBM peak; `orbit` repeats that on a
            stratified sample of the 1000-frame orbit (frames 0, 125, .., 875) with min / mean.
  raycast   rays/s and samples/s of the march (+ continuation + normals) over the timed frames; samples
            evaluated against the samples the reference's fixed-step march executes for the same frame.
  cpu_baseline / --impl reference: the CPU restatement of the reference kernels (oracle/, the reference
            has no CPU path of its own) on all host cores, WHOLE frames of the same workload.
  ref_cuda  the reference's own CUDA classes and kernels (oracle/_ref: rebuilt for sm_100 as shipped, -G,
            and with -O3) on the same frames at the same size, host buffers, on this GPU.

Multi-GPU (torchrun, --gpus N): the volume is sharded along Z; `extra.scale_1024` is the
same step at 1024^3 (BASELINE configs[3]).
"""
import argparse
import contextlib
import importlib.util
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
                    help="multi-GPU layout (tsdf_b200/sharded.py): Z-slabs + key exchange, or surface replicas + image tiles")
    ap.add_argument("--exchange", default="peer", choices=["peer", "allreduce"],
                    help="multi-GPU key exchange of the contiguous layout: atomics into rank 0's key map over NVLink, or NCCL all-reduce(min)")
    ap.add_argument("--slab", type=int, default=0, help="planes per slab for the interleaved / replica layouts (0: size / gpus)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-orbit", action="store_true", help="skip the stratified-orbit roofline sample")
    ap.add_argument("--no-1024", action="store_true", help="skip extra.scale_1024")
    return ap.parse_args()


def _scenes():
    """tsdf_b200/scenes.py loaded as a plain file: the reference arm must not import the tsdf_b200 package (its
    __init__ loads libtsdf_b200.so, and nothing of the product may be mapped into the reference arm's process)."""
    mod = sys.modules.get("_bench_scenes")
    if mod is None:
        spec = importlib.util.spec_from_file_location("_bench_scenes", os.path.join(ROOT, "tsdf_b200", "scenes.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["_bench_scenes"] = mod
    return mod


def frame_inputs(i):
    scenes = _scenes()
    cam = scenes.orbit_camera(i % ORBIT_FRAMES, ORBIT_FRAMES)
    return cam, scenes.render_depth(cam, W, H)


def make_config(size, world, layout="contiguous", slab=0):
    """The workload description — identical for the b200 arm and the reference arm."""
    return {"workload": f"{size}^3 volume / 3000 mm, 640x480 depth, {ORBIT_FRAMES}-frame orbit of sphere+wall "
                        f"(BASELINE configs[2]); step = integrate + raycast + normals of one frame",
            "cache": f"volume ({size ** 3 * 8 / 2 ** 30:g} GiB dist+weight) is larger than L2, no flush needed",
            "parallelism": ("single GPU" if world <= 1 else
                            (f"Z-slab sharding over {world} GPUs, one key exchange (min) per frame" if layout != "replica" else
                             f"Z-slabs of {slab} planes over {world} GPUs, surface bricks pushed to per-GPU replicas over NVLink, "
                             f"image tiles sharded"))}


def default_slab(args, world):
    return args.slab if args.slab > 0 else max(8, (args.size // max(world, 1)) // 8 * 8)


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
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
                "reasons": sorted(self.reasons)}


# ---- CPU restatement (oracle/): cpu_baseline leg and --impl reference ---------------------------------------------------
class CpuArm:
    """Whole frames of the workload on the host cores: integrate the frame into the whole volume, then raycast every
    pixel (+ normals) from its pose — the same step the GPU arm times."""

    def __init__(self, size, threads=None):
        from oracle import oracle
        self.cores = oracle.set_threads(threads or (os.cpu_count() or 1))
        self.vol = oracle.OracleVolume((size,) * 3, PHYS)
        self.size = size

    def frame(self, i):
        cam, depth = frame_inputs(i)
        t0 = time.perf_counter()
        self.vol.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
        t1 = time.perf_counter()
        _, _, _, marched = self.vol.raycast(W, H, cam.pose, cam.kinv, want_khit=False)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, marched


def host_info():
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"nproc": os.cpu_count(), "cpu_model": model}


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores (CPU restatement, kind "port"), whole frames."""
    if rank != 0:
        return
    size = args.size
    arm = CpuArm(size)
    for i in range(args.warmup):
        arm.frame(i)
    t_int, t_ray = [], []
    for s in range(args.steps):
        a, b, _ = arm.frame(args.warmup + s)
        t_int.append(a); t_ray.append(b)
    ms = 1e3 * float(np.mean(t_int) + np.mean(t_ray))
    val = 1e3 / ms
    sample = (f"every step is one WHOLE orbit frame at {size}^3: integrate all {size} planes + raycast all {W}x{H} rays + normals "
              f"(integrate {1e3 * np.mean(t_int):.0f} ms + raycast {1e3 * np.mean(t_ray):.0f} ms per frame, OpenMP on {arm.cores} threads)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(size, max(world, args.gpus), args.layout, default_slab(args, max(world, args.gpus))),
        "note": "the reference has no CPU path (SURVEY.md fact 2); this is the line-by-line CPU restatement of its CUDA kernels",
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": arm.cores, "kind": "port", "sample": sample, "host": host_info()},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- the reference's own CUDA path (oracle/_ref), timed next to the product on the same frames -------------------------------
@contextlib.contextmanager
def _stdout_to_devnull():
    """The reference prints from host code (std::cout per call) and from the device (one printf per ray that reaches the
    4402-sample cap, GPURaycaster.cu:370): fd 1 goes to /dev/null while it runs."""
    sys.stdout.flush()
    saved = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(null, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(null)
        os.close(saved)


def time_ref_cuda(size, first_frame, n_frames):
    import torch
    from oracle import refcuda
    out = {}
    for tag, label in (("O3", "O3"), ("G", "G_as_shipped")):
        if not refcuda.available(tag):
            out[label] = {"unavailable": f"oracle/_ref/libref_cuda_{tag}.so not built"}
            continue
        lib = refcuda.RefLib(tag)
        with _stdout_to_devnull():
            rv = refcuda.RefVolume(lib, (size,) * 3, PHYS)
            ti, tr = [], []
            for j in range(n_frames + 1):
                cam, depth = frame_inputs(first_frame + j)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rv.integrate(depth, cam.k, cam.pose)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                rv.raycast(W, H, cam.k, cam.pose)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                if j > 0:                                   # first call: lazy module load
                    ti.append(t1 - t0); tr.append(t2 - t1)
            rv.close()
        out[label] = {"frames_per_s": 1.0 / (float(np.median(ti)) + float(np.median(tr))),
                      "integrate_ms": 1e3 * float(np.median(ti)), "raycast_ms": 1e3 * float(np.median(tr)), "frames": n_frames}
    out["how"] = (f"reference TSDFVolume::integrate / ::raycast (its own classes and kernels, /root/reference/src compiled for sm_100 by "
                  f"oracle/build_ref.sh: -G as shipped in kinfu.make:64, and -O3 -fmad=false) at {size}^3 on orbit frames "
                  f"{first_frame + 1}..{first_frame + n_frames}, host buffers, per-call malloc/copy/sync as written, wall clock, median")
    return out



def class_layer_e2e(size, frames, cams, warmup, steps, W, H):
    """The same frames through the drop-in C++ classes (TSDFVolume::integrate / ::raycast, Eigen result matrices declared per
    frame, as the reference's kinfu.cpp does): tools/class_e2e.cpp, built by `make class_e2e`.  None when the binary is absent."""
    import struct, subprocess, tempfile
    exe = os.path.join(ROOT, "build", "class_e2e")
    if not os.path.exists(exe):
        return None
    n = warmup + steps
    try:
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
            f.write(struct.pack("<4If", size, W, H, n, PHYS[0]))
            f.write(np.asarray(cams[0].k, np.float32).tobytes())
            for i in range(n):
                f.write(np.asarray(cams[i].pose, np.float32).tobytes())
                f.write(np.ascontiguousarray(frames[i], np.uint16).tobytes())
            path = f.name
        out = subprocess.run([exe, path, str(warmup)], capture_output=True, text=True, timeout=600)
        os.unlink(path)
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not lines:
            return {"error": f"rc {out.returncode}: " + (out.stdout + out.stderr)[-300:]}
        r = json.loads(lines[-1])
        return {"value": r["frames_per_s"], "unit": "frames/s", "ms_per_step": r["ms_per_frame"], "frames": r["frames"],
                "api": "TSDFVolume::integrate + TSDFVolume::raycast of the drop-in C++ class layer, DepthImage pixels and the Eigen "
                       "result matrices (declared per frame, like kinfu.cpp:174-181) in the library's pinned pool (tools/class_e2e.cpp)"}
    except Exception as e:      # the tool is an extra: never fail the bench line for it
        return {"error": repr(e)[:200]}

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

    K, Wm = args.steps, args.warmup
    stream = torch.cuda.current_stream()

    def run_device_resident(size, K, Wm, orbit_sample):
        """The timed loop with frames resident in HBM.  Returns a dict of measurements (rank-local; times max-reduced)."""
        n = (size, size, size)
        cams, frames = [], []
        for i in range(Wm + K):
            cam, depth = frame_inputs(i)
            cams.append(cam)
            frames.append(depth)
        d_frames = [torch.from_numpy(f).cuda() for f in frames]
        slab = args.slab if args.slab > 0 else max(8, (size // max(world, 1)) // 8 * 8)
        eng = sharded.ShardedEngine(n, PHYS, rank, world, stream=stream, layout=args.layout, slab=slab, exchange=args.exchange)
        for i in range(Wm):
            eng.integrate(d_frames[i], cams[i])
            eng.raycast(W, H, cams[i])
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(K)]
        token = torch.zeros(1, device="cuda")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.all_reduce(token)                      # stream-ordered rendezvous: every rank's ev0 follows it on the device
        ev0.record(stream)
        for s in range(K):
            i = Wm + s
            eng.stage(d_frames[i])                      # culling pyramid of the frame
            iev[s][0].record(stream)                    # events bracket the integrate kernel alone (roofline) ...
            eng.integrate(d_frames[i], cams[i], count=False, restage=False)
            iev[s][1].record(stream)                    # ... and the raycast (+ normals, + exchange when sharded)
            eng.raycast(W, H, cams[i])
            iev[s][2].record(stream)
        if world > 1:
            dist.all_reduce(token)                      # the step count is done when the slowest rank is
        ev1.record(stream)
        torch.cuda.synchronize()
        clocks = sampler.stop()
        total_ms = ev0.elapsed_time(ev1)
        t_int_ms = [a.elapsed_time(b) for a, b, _ in iev]
        t_ray_ms = [b.elapsed_time(c) for _, b, c in iev]
        if world > 1:
            t = torch.tensor([total_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        # ---- untimed: voxels rewritten per timed frame (geometry only) -> algorithmic bytes; samples evaluated ----
        n_upd = [eng.integrate(d_frames[Wm + s], cams[Wm + s], count=True) for s in range(K)]
        n_samples = [eng.raycast(W, H, cams[Wm + s], count=True) for s in range(min(K, 8))]
        ray_stats = eng.last_ray_stats() or {}
        orbit = None
        if orbit_sample and world == 1:
            orbit = []
            for f in range(0, ORBIT_FRAMES, ORBIT_FRAMES // 8):
                cam, depth = frame_inputs(f)
                d = torch.from_numpy(depth).cuda()
                upd = eng.integrate(d, cam, count=True)
                ts = []
                for _ in range(5):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda._sleep(400000)     # the launches queue up behind a busy GPU: the events bracket device time only
                    a.record(stream); eng.integrate(d, cam, restage=False); b.record(stream)
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                orbit.append((f, upd, float(np.median(ts))))
        res = {"total_ms": total_ms, "t_int_ms": t_int_ms, "t_ray_ms": t_ray_ms, "n_upd": n_upd, "n_samples": n_samples,
               "ray_stats": ray_stats, "clocks": clocks, "launches_per_step": eng.launches_per_step, "orbit": orbit,
               "frames": frames, "cams": cams, "slab": slab}
        res["e2e_sharded"] = eng.e2e(frames, cams, Wm, K, W, H) if (world > 1 and not args.no_e2e and size == args.size) else None
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        eng.close()
        if world > 1:
            dist.barrier()
        return res

    size = args.size
    n = (size, size, size)
    R = run_device_resident(size, K, Wm, orbit_sample=not args.no_orbit)
    ms_per_step = R["total_ms"] / K
    value = 1e3 / ms_per_step
    frames, cams = R["frames"], R["cams"]

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    b_alg = [16.0 * u + W * H * 2 for u in R["n_upd"]]
    achieved = sum(b_alg) / (sum(R["t_int_ms"]) * 1e-3) / 1e9
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    traffic = prof.get("integrate_dram_bytes_per_launch") if prof.get("size") == size else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": prof.get("source") if traffic else None,
                "kernel": "integrate_rigid_kernel",
                "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback (B200_PROFILING.md)",
                "frac_of_nominal_8000": achieved / 8000.0,
                "algorithmic_bytes_per_launch": float(np.mean(b_alg)),
                "integrate_ms_per_launch": float(np.mean(R["t_int_ms"])),
                "voxels_rewritten_per_frame": float(np.mean(R["n_upd"])), "per_gpu": world > 1,
                "timed_frames": f"orbit frames {Wm}..{Wm + K - 1}"}
    if R["orbit"]:
        rows = []
        for f, upd, ms in R["orbit"]:
            gbs = (16.0 * upd + W * H * 2) / (ms * 1e-3) / 1e9
            rows.append({"frame": f, "voxels_rewritten": upd, "integrate_us": 1e3 * ms, "GB/s": gbs, "frac": gbs / peak})
        roofline["orbit"] = {"frames": rows, "frac_min": min(r["frac"] for r in rows),
                             "frac_mean": float(np.mean([r["frac"] for r in rows])),
                             "what": "stratified sample of the 1000-frame orbit (every 125th frame), median of 5 launches each"}

    # ---- raycast work metrics (rays/s and samples/s, no DRAM fraction) -----------------------------
    ray_ms = float(np.mean(R["t_ray_ms"]))
    raycast = dict(R["ray_stats"])
    raycast.update({"ms_per_frame": ray_ms, "rays_per_s": W * H / (ray_ms * 1e-3),
                    "samples_evaluated": float(np.mean(R["n_samples"])) if R["n_samples"] else None,
                    "samples_per_s": (float(np.mean(R["n_samples"])) / (ray_ms * 1e-3)) if R["n_samples"] else None,
                    "includes": "brick distance transform + march + continuation + normals" + (" + key exchange + resolve" if world > 1 else ""),
                    "l1_hit_rate_pct": prof.get("raycast_l1_hit_pct"), "l2_hit_rate_pct": prof.get("raycast_l2_hit_pct"),
                    "hit_rate_source": prof.get("raycast_source")})

    # ---- e2e: level-2 C-ABI with host buffers ---------------------------------------------------------------------------
    e2e, e2e_pageable, e2e_classes = R["e2e_sharded"], None, None
    if not args.no_e2e and world == 1:
        vol = Volume(n, PHYS)
        mats = [(colmajor(c.inv_pose), colmajor(c.k), colmajor(c.kinv), colmajor(c.pose)) for c in cams]
        f_int, f_ray, handle = lib.tsdf_b200_volume_integrate, lib.tsdf_b200_volume_raycast, vol._h

        def loop(depth_arrays, hv_np, hn_np):
            # argument marshalling done once: a C or C++ caller of the C-ABI has none of it (ctypes pointer objects cost
            # microseconds each, during which the GPU would idle inside the timed region)
            args_i = [(C.c_void_p(depth_arrays[i].ctypes.data), fptr(m[0]), fptr(m[1]), fptr(m[2]), fptr(m[3])) for i, m in enumerate(mats)]
            hv_p, hn_p = C.c_void_p(hv_np.ctypes.data), C.c_void_p(hn_np.ctypes.data)

            def step(i):
                depth_p, ip, k, kinv, pose = args_i[i]
                check(f_int(handle, depth_p, W, H, ip, k, kinv))
                check(f_ray(handle, W, H, pose, kinv, hv_p, hn_p))

            check(lib.tsdf_b200_volume_clear(handle))
            for i in range(Wm):
                step(i)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for s in range(K):
                step(Wm + s)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) * 1e3 / K

        pin = [torch.from_numpy(f).pin_memory() for f in frames]
        hv = torch.empty((H * W, 3), dtype=torch.float32).pin_memory()
        hn = torch.empty((H * W, 3), dtype=torch.float32).pin_memory()
        e2e_ms = loop([p.numpy() for p in pin], hv.numpy(), hn.numpy())
        e2e = {"value": 1e3 / e2e_ms, "unit": "frames/s", "h2d_bytes_per_step": W * H * 2,
               "d2h_bytes_per_step": 2 * W * H * 3 * 4, "ms_per_step": e2e_ms,
               "api": "tsdf_b200_volume_integrate + tsdf_b200_volume_raycast (pinned host buffers, synchronous)"}
        pg_ms = loop(frames, np.empty((H * W, 3), np.float32), np.empty((H * W, 3), np.float32))
        e2e_pageable = {"value": 1e3 / pg_ms, "unit": "frames/s", "ms_per_step": pg_ms,
                        "api": "the same calls with pageable (malloc'ed) depth and result buffers — what kinfu's DepthImage::data() "
                               "and Eigen matrices are (the driver stages the copies)"}
        vol.close()
        e2e_classes = class_layer_e2e(size, frames, cams, Wm, K, W, H)

    # ---- baselines on the same box: CPU restatement (whole frames), reference CUDA --------------------------------------
    cpu_baseline, ref_cuda = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(size)
        arm.frame(Wm)                                    # warm-up frame (also gives the raycast a surface)
        ti, tr, marched = arm.frame(Wm + 1)
        cpu_baseline = {"value": 1.0 / (ti + tr), "unit": "frames/s", "cores": arm.cores, "kind": "port",
                        "sample": f"ONE whole orbit frame at {size}^3 (after one warm-up frame): integrate {ti * 1e3:.0f} ms + raycast "
                                  f"{tr * 1e3:.0f} ms with the CPU restatement of the reference kernels (OpenMP, {arm.cores} threads)",
                        "host": host_info()}
        raycast["oracle_samples"] = marched
        raycast["oracle_samples_what"] = "trilinear samples the reference's fixed-step march executes for one such frame (CPU restatement)"
        del arm
    if rank == 0 and world == 1 and not args.no_ref_cuda:
        try:
            ref_cuda = time_ref_cuda(size, Wm, 3)
            if "frames_per_s" in ref_cuda.get("O3", {}):
                ref_cuda["speedup_e2e_pageable_vs_O3"] = (e2e_pageable or e2e or {}).get("value", 0.0) / ref_cuda["O3"]["frames_per_s"]
            if "frames_per_s" in ref_cuda.get("G_as_shipped", {}):
                ref_cuda["speedup_e2e_pageable_vs_G"] = (e2e_pageable or e2e or {}).get("value", 0.0) / ref_cuda["G_as_shipped"]["frames_per_s"]
        except Exception as ex:                       # the baseline must never take the bench line down
            ref_cuda = {"unavailable": f"{type(ex).__name__}: {ex}"}

    # ---- BASELINE configs[3]: the same step at 1024^3 (8 GiB of dist+weight), so that the scaling run shows it -----------
    extra = {}
    if not args.no_1024 and size == 512:
        try:
            k2, w2 = min(K, 10), 3
            R2 = run_device_resident(1024, k2, w2, orbit_sample=False)
            b2 = [16.0 * u + W * H * 2 for u in R2["n_upd"]]
            g2 = sum(b2) / (sum(R2["t_int_ms"]) * 1e-3) / 1e9
            extra["scale_1024"] = {"value": 1e3 * k2 / R2["total_ms"], "unit": "frames/s", "n_gpus": world, "steps": k2, "warmup": w2,
                                   "ms_per_step": R2["total_ms"] / k2, "integrate_ms": float(np.mean(R2["t_int_ms"])),
                                   "raycast_ms": float(np.mean(R2["t_ray_ms"])), "integrate_frac_of_hbm_peak": g2 / peak,
                                   "workload": "1024^3 volume / 3000 mm, 640x480, same orbit (BASELINE configs[3])", "per_gpu_roofline": world > 1}
        except Exception as ex:
            extra["scale_1024"] = {"unavailable": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        cfg = make_config(size, world, args.layout, default_slab(args, world))
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "clocks": R["clocks"], "e2e": e2e, "e2e_pageable": e2e_pageable, "e2e_classes": e2e_classes, "gpu_launches": R["launches_per_step"] * K,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "ref_cuda": ref_cuda, "raycast": raycast, "extra": extra,
        }
        print(json.dumps(out))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
