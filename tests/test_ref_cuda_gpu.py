"""GPU tests that PIN THE ORACLE: the reference's own CUDA code (oracle/_ref, built from /root/reference/src by
oracle/build_ref.sh with -O3 -fmad=false, and as shipped with -G) is run on the B200 and compared bit for bit
with the CPU restatement (oracle/tsdf_oracle.c) and with the sm_100a kernels, through the same host calls
kinfu.cpp makes (TSDFVolume::integrate with a Camera, TSDFVolume::raycast)."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import assert_bits_equal, random_rigid_pose, random_depth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import refcuda
    if not refcuda.available("O3"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return refcuda


def scaled(cam, s):
    k = cam.k.copy()
    k[:2] *= s
    return k


def test_reference_clear_and_file_match_fixture_and_oracle(R, tmp_path):
    """The reference's clear() + save_to_file on the GPU reproduce the reference's committed fixture
    (TestData/t_100_2000_50.tsdf) — and so does the oracle (tests/test_oracle_cpu.py)."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "t_100_2000_50.json")))
    lib = R.RefLib("O3")
    v = R.RefVolume(lib, g["size"], g["physical"])
    assert int(np.float32(v.trunc).view(np.uint32)) == g["trunc_bits"]
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    d, w = v.read()
    assert sha(d) == g["dist_sha256"] and sha(w) == g["weight_sha256"]
    assert sha(v.read_deformation()) == g["deform_sha256"]
    v.offset(*g["offset"])
    path = str(tmp_path / "ref.tsdf")
    v.save(path)
    raw = open(path, "rb").read()
    assert len(raw) == g["file_bytes"] and raw[:68].hex() == g["header_hex"]
    # and the product loads the reference-written file
    from tsdf_b200 import Volume
    mine = Volume.load(path)
    d2, w2 = mine.read()
    assert_bits_equal(d2, d, "dist loaded from reference file")
    assert mine.size == tuple(g["size"]) and mine.trunc == v.trunc
    mine.close(); v.close()


@pytest.mark.parametrize("tag", ["O3", "G"])
def test_reference_cuda_equals_oracle_and_kernels(R, tag):
    """integrate x3 + raycast through the reference classes == oracle == tsdf_b200, bit for bit."""
    if not R.available(tag):
        pytest.skip(f"libref_cuda_{tag}.so missing")
    from oracle import oracle
    from tsdf_b200 import Volume, scenes
    lib = R.RefLib(tag)
    n = (64, 64, 64) if tag == "O3" else (48, 48, 48)
    w, h = (320, 240) if tag == "O3" else (160, 120)
    s = w / 640.0
    rv = R.RefVolume(lib, n, (3000, 3000, 3000))
    ov = oracle.OracleVolume(n, (3000, 3000, 3000), with_deformation=True)
    mv = Volume(n, (3000.0, 3000.0, 3000.0))
    assert rv.trunc == ov.trunc == mv.trunc
    for f in (0, 3, 7):
        cam = scenes.orbit_camera(f, 12)
        k = scaled(cam, s)
        kinv, inv_pose = lib.camera_matrices(k, cam.pose)      # what the reference Camera feeds its kernels
        depth = scenes.render_depth(cam, w, h)
        rv.integrate(depth, k, cam.pose)
        ov.integrate(depth, inv_pose, k, kinv)
        mv.integrate(depth, inv_pose, k, kinv)
    dr, wr = rv.read()
    dm, wm = mv.read()
    assert_bits_equal(dr, ov.dist, f"[{tag}] reference CUDA dist vs oracle")
    assert_bits_equal(wr, ov.weight, f"[{tag}] reference CUDA weight vs oracle")
    assert_bits_equal(dm, dr, f"[{tag}] tsdf_b200 dist vs reference CUDA")
    assert_bits_equal(wm, wr, f"[{tag}] tsdf_b200 weight vs reference CUDA")
    Vr, Nr = rv.raycast(w, h, k, cam.pose)
    Vo, No, ko, so = ov.raycast(w, h, cam.pose, kinv)
    Vm, Nm = mv.raycast(w, h, cam.pose, kinv)
    assert (ko >= 0).sum() > 100
    assert_bits_equal(Vr, Vo, f"[{tag}] reference CUDA vertices vs oracle")
    assert_bits_equal(Nr, No, f"[{tag}] reference CUDA normals vs oracle")
    assert_bits_equal(Vm, Vr, f"[{tag}] tsdf_b200 vertices vs reference CUDA")
    assert_bits_equal(Nm, Nr, f"[{tag}] tsdf_b200 normals vs reference CUDA")
    # north-star wording: SDF within 1e-4 relative, hit voxel indices bit-exact (implied by the above)
    assert np.allclose(dm, dr, rtol=1e-4, atol=0)
    hv = oracle.hit_voxels(Vm, ov.offset, ov.voxel, n[0], n[1])
    hr = oracle.hit_voxels(Vr, ov.offset, ov.voxel, n[0], n[1])
    assert np.array_equal(hv, hr)
    rv.close(); mv.close()


def test_reference_cuda_random_poses_sphere(R):
    """Sphere SDF uploaded with set_distance_data (the reference's create_sphere_in_TSDF route), random cameras,
    offset volume: reference raycast == oracle == tsdf_b200."""
    from oracle import oracle
    from tsdf_b200 import Volume
    from test_parity_gpu import sphere_sdf
    lib = R.RefLib("O3")
    rng = np.random.default_rng(77)
    n = (72, 64, 56)
    phys = (2800.0, 3000.0, 2600.0)
    rv = R.RefVolume(lib, n, phys)
    ov = oracle.OracleVolume(n, phys)
    mv = Volume(n, phys)
    for vol in (rv,):
        vol.offset(100.0, -40.0, 60.0)
    ov.offset[:] = (100.0, -40.0, 60.0)
    mv.set_offset(100.0, -40.0, 60.0)
    sdf = sphere_sdf(n, phys, ov.trunc, (1400, 1500, 1300), 700)
    rv.set_distance_data(sdf); ov.dist[:] = sdf; mv.set_distance_data(sdf)
    for _ in range(3):
        cam = random_rigid_pose(rng, centre=(1500, 1460, 1360), radius=(1800.0, 4500.0))
        k = scaled(cam, 0.25)
        kinv, inv_pose = lib.camera_matrices(k, cam.pose)
        Vr, Nr = rv.raycast(160, 120, k, cam.pose)
        Vo, No, ko, so = ov.raycast(160, 120, cam.pose, kinv)
        Vm, Nm = mv.raycast(160, 120, cam.pose, kinv)
        assert_bits_equal(Vr, Vo, "reference CUDA vertices vs oracle")
        assert_bits_equal(Vm, Vr, "tsdf_b200 vertices vs reference CUDA")
        assert_bits_equal(Nm, Nr, "tsdf_b200 normals vs reference CUDA")
        depth = random_depth(rng, 160, 120, lo=1000, hi=5000)
        rv.integrate(depth, k, cam.pose); ov.integrate(depth, inv_pose, k, kinv); mv.integrate(depth, inv_pose, k, kinv)
    dr, wr = rv.read()
    dm, wm = mv.read()
    assert_bits_equal(dr, ov.dist, "reference CUDA dist vs oracle")
    assert_bits_equal(dm, dr, "tsdf_b200 dist vs reference CUDA")
    assert_bits_equal(wm, wr, "tsdf_b200 weight vs reference CUDA")
    rv.close(); mv.close()


def test_reference_marching_cubes_equals_oracle_and_gpu(R):
    """the reference's own extract_surface (MarchingCubes/MarkAndSweepMC.cu:506-555, compiled
    unmodified into oracle/_ref) == the CPU oracle == tsdf_b200_mc_extract, vertex for vertex and bit for bit — on a
    noisy sphere uploaded with set_distance_data (offset volume, anisotropic voxels) and on a volume fused from orbit
    frames by the reference itself."""
    import ctypes as C
    import torch
    from oracle import oracle
    from tsdf_b200 import scenes
    from tsdf_b200.capi import lib as mylib, check, fptr
    from test_parity_gpu import sphere_sdf
    from test_mc_gpu import gpu_mc
    lib = R.RefLib("O3")
    rng = np.random.default_rng(5)

    def compare(rv, n, what):
        d, _ = rv.read()
        vox, off = rv.voxel, np.asarray(rv._offset, np.float32)
        want = rv.extract_surface()
        assert want.shape[0] > 1000 and want.shape[0] % 3 == 0
        assert_bits_equal(oracle.mc_extract(d, n, vox, off), want, f"{what}: oracle vs reference marching cubes")
        got = gpu_mc(mylib, check, fptr, torch.from_numpy(d).cuda(), n, 0, 0, n[2] - 1, vox, off)
        assert got.shape == want.shape
        assert_bits_equal(got, want, f"{what}: tsdf_b200_mc_extract vs reference marching cubes")

    n, phys = (72, 64, 56), (2800.0, 3000.0, 2600.0)
    rv = R.RefVolume(lib, n, phys)
    rv.offset(100.0, -40.0, 60.0); rv._offset = (100.0, -40.0, 60.0)
    sdf = sphere_sdf(n, phys, rv.trunc, (1400, 1500, 1300), 700)
    sdf = (sdf + rng.normal(scale=0.02 * float(rv.trunc), size=sdf.shape)).astype(np.float32)
    sdf[rng.integers(0, sdf.size, size=7)] = 0.0              # exact zeros count as outside (:139-146)
    rv.set_distance_data(sdf)
    compare(rv, n, "noisy sphere")
    rv.close()

    n, phys = (64, 64, 64), (3000.0, 3000.0, 3000.0)
    rv = R.RefVolume(lib, n, phys); rv._offset = (0.0, 0.0, 0.0)
    for f in (0, 3, 7):
        cam = scenes.orbit_camera(f, 12)
        k = scaled(cam, 0.5)
        rv.integrate(scenes.render_depth(cam, 320, 240), k, cam.pose)
    compare(rv, n, "fused orbit volume")
    rv.close()


def test_reference_render_to_depth_image_values(R, tmp_path):
    """GPURaycaster::render_to_depth_image (RayCaster/GPURaycaster.cu:555-606 — raycast, then
    (uint16_t)roundf(world_to_camera(vertex).z) per pixel, :577-581) through the drop-in class == the reference's own
    function, pixel for pixel."""
    import subprocess
    from test_classes_cpu import build_class_tests
    from test_parity_gpu import sphere_sdf
    from tsdf_b200 import scenes
    from tsdf_b200.capi import colmajor
    lib = R.RefLib("O3")
    n, phys, w, h = 64, 3000.0, 320, 240
    rv = R.RefVolume(lib, (n,) * 3, (phys,) * 3)
    sdf = sphere_sdf((n,) * 3, (phys,) * 3, rv.trunc, (1500, 1450, 1550), 900)
    rv.set_distance_data(sdf)
    exe = build_class_tests()
    rng = np.random.default_rng(11)
    total_hits = 0
    for i in range(3):
        cam = random_rigid_pose(rng, radius=(2200.0, 4200.0))
        k = scaled(cam, 0.5)
        want = rv.render_depth(w, h, k, cam.pose)
        sdf.tofile(tmp_path / "dist.f32")
        np.concatenate([colmajor(k), colmajor(cam.pose)]).astype(np.float32).tofile(tmp_path / "camera.f32")
        out = subprocess.run([exe, "--render-depth", str(tmp_path), str(n), str(phys), str(w), str(h)], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        got = np.fromfile(tmp_path / "depth.u16", np.uint16).reshape(h, w)
        assert np.array_equal(got, want), f"pose {i}: {(got != want).sum()} of {got.size} depth pixels differ"
        total_hits += int((want > 0).sum())
    assert total_hits > 20000
    rv.close()


class _Raw:
    """A raw device pointer as a torch tensor (no copy): full-size volumes are compared where they live."""

    def __init__(self, ptr, count, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _dev_i32(ptr, count):
    import torch
    return torch.as_tensor(_Raw(ptr, count, "<i4"), device="cuda")


def _fullsize_against_reference(R, size, frames, raycast_frames, w, h, slabs=0):
    """Integrates orbit frames into a size^3 volume with the reference's own classes (oracle/_ref, -O3 -fmad=false) and with
    the level-2 C-ABI, compares dist / weight on the device and vertices / normals / hit voxels on the host, bit for bit."""
    import torch
    from oracle import oracle
    from tsdf_b200 import Volume, scenes
    lib = R.RefLib("O3")
    n, phys = (size,) * 3, (3000.0,) * 3
    nv = size ** 3
    with R.quiet():
        rv = R.RefVolume(lib, n, phys)
    mv = Volume(n, phys)
    assert rv.trunc == mv.trunc
    V = {}
    for f in frames:
        cam = scenes.orbit_camera(f, 1000)
        kinv, inv_pose = lib.camera_matrices(cam.k, cam.pose)
        depth = scenes.render_depth(cam, w, h)
        with R.quiet():
            rv.integrate(depth, cam.k, cam.pose)
        mv.integrate(depth, inv_pose, cam.k, kinv)
        if f in raycast_frames:
            with R.quiet():                                   # capped rays print one line each (GPURaycaster.cu:370)
                Vr, Nr = rv.raycast(w, h, cam.k, cam.pose)
            Vm, Nm = mv.raycast(w, h, cam.pose, kinv)
            assert_bits_equal(Vm, Vr, f"{size}^3 frame {f}: vertices vs reference CUDA")
            assert_bits_equal(Nm, Nr, f"{size}^3 frame {f}: normals vs reference CUDA")
            hv = oracle.hit_voxels(Vm, np.zeros(3, np.float32), mv.voxel, size, size)
            hr = oracle.hit_voxels(Vr, np.zeros(3, np.float32), mv.voxel, size, size)
            assert np.array_equal(hv, hr), f"{size}^3 frame {f}: hit voxel indices"
            V[f] = (Vr, int((~np.isnan(Vr[:, 0])).sum()), cam, kinv)
        torch.cuda.synchronize()
        assert torch.equal(_dev_i32(rv.distance_ptr, nv), _dev_i32(mv.distance_ptr, nv)), f"{size}^3 frame {f}: dist vs reference CUDA"
        assert torch.equal(_dev_i32(rv.weight_ptr, nv), _dev_i32(mv.weight_ptr, nv)), f"{size}^3 frame {f}: weight vs reference CUDA"
    assert sum(v[1] for v in V.values()) > 50000
    if slabs:
        # the multi-GPU raycast (Z-slabs, key min, resolve), ranks emulated on this GPU, against the reference's vertices
        import gpu_util
        dv = gpu_util.DeviceVolume.__new__(gpu_util.DeviceVolume)
        dv.n, dv.physical = n, np.asarray(phys, np.float32)
        dv.voxel, dv.trunc = mv.voxel, mv.trunc
        dv.offset = np.zeros(3, np.float32)
        dv.dist = torch.as_tensor(_Raw(mv.distance_ptr, nv), device="cuda")
        dv.table = torch.empty(4416, dtype=torch.float32, device="cuda")
        from tsdf_b200.capi import lib as mylib, check
        import ctypes as C
        check(mylib.tsdf_b200_ray_table(dv.trunc, C.c_void_p(dv.table.data_ptr()), None))
        f = max(V)
        Vr, _, cam, kinv = V[f]
        Vs, _, _ = gpu_util.raycast_sharded_on_one_gpu(dv, w, h, cam.pose, kinv, slabs)
        assert_bits_equal(Vs, Vr, f"{size}^3 frame {f}: {slabs}-slab raycast vs reference CUDA")
    with R.quiet():
        rv.close()
    mv.close()


def test_reference_cuda_512_headline_config(R):
    """BASELINE configs[2] at its real size: 512^3 / 3000 mm, 640x480, orbit frames 3, 250, 500, 750 fused by the
    reference's own CUDA path and by tsdf_b200; raycast after the first and the last of them."""
    _fullsize_against_reference(R, 512, [3, 250, 500, 750], {3, 750}, 640, 480)


def test_reference_cuda_1024_slabs(R):
    """BASELINE configs[3] at its real size: 1024^3 (8 GiB of dist + weight here, 35 GiB in the reference with its
    deformation and colour arrays): two orbit frames, single-GPU raycast and the 8-slab sharded raycast (ranks emulated)
    against the reference's vertices."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~50 GiB of free device memory")
    _fullsize_against_reference(R, 1024, [3, 130], {130}, 640, 480, slabs=8)


def test_reference_cuda_timing_at_256(R):
    """SURVEY.md section 8d, "also time": the reference's own CUDA path (its classes and kernels, rebuilt -O3 for sm_100,
    host buffers, per-call allocations as shipped) next to tsdf_b200's level-2 calls on the same frames, 256^3 (BASELINE
    configs[1]).  The volumes must agree bit for bit; the times go to gpurun_out/ref_cuda_timing.json for DESIGN.md."""
    import time
    import torch
    from tsdf_b200 import Volume, scenes
    lib = R.RefLib("O3")
    n, phys, w, h = (256, 256, 256), (3000.0, 3000.0, 3000.0), 640, 480
    rv = R.RefVolume(lib, n, phys)
    mv = Volume(n, phys)
    frames = [scenes.orbit_camera(f, 1000) for f in range(6)]
    depths = [scenes.render_depth(c, w, h) for c in frames]
    mats = [lib.camera_matrices(c.k, c.pose) for c in frames]
    t = {"ref_integrate_ms": [], "ref_raycast_ms": [], "b200_integrate_ms": [], "b200_raycast_ms": []}
    Vbuf, Nbuf = np.zeros((h * w, 3), np.float32), np.zeros((h * w, 3), np.float32)
    for i, cam in enumerate(frames):
        kinv, inv_pose = mats[i]
        torch.cuda.synchronize()
        t0 = time.perf_counter(); rv.integrate(depths[i], cam.k, cam.pose); torch.cuda.synchronize(); t1 = time.perf_counter()
        Vr, Nr = rv.raycast(w, h, cam.k, cam.pose); torch.cuda.synchronize(); t2 = time.perf_counter()
        mv.integrate(depths[i], inv_pose, cam.k, kinv); torch.cuda.synchronize(); t3 = time.perf_counter()
        Vm, Nm = mv.raycast(w, h, cam.pose, kinv, Vbuf, Nbuf); torch.cuda.synchronize(); t4 = time.perf_counter()
        if i >= 2:                      # first calls: lazy kernel loading, self-test, allocations
            t["ref_integrate_ms"].append((t1 - t0) * 1e3); t["ref_raycast_ms"].append((t2 - t1) * 1e3)
            t["b200_integrate_ms"].append((t3 - t2) * 1e3); t["b200_raycast_ms"].append((t4 - t3) * 1e3)
    assert_bits_equal(Vm, Vr, "vertices, tsdf_b200 vs reference CUDA at 256^3")
    assert_bits_equal(Nm, Nr, "normals, tsdf_b200 vs reference CUDA at 256^3")
    dr, wr = rv.read()
    dm, wm = mv.read()
    assert_bits_equal(dm, dr, "dist at 256^3")
    assert_bits_equal(wm, wr, "weight at 256^3")
    out = {k: float(np.median(v)) for k, v in t.items()}
    out["config"] = "256^3 / 3000 mm, 640x480, orbit frames 2..5, host buffers (pageable), synchronous calls"
    out["ref_frames_per_s"] = 1e3 / (out["ref_integrate_ms"] + out["ref_raycast_ms"])
    out["b200_frames_per_s"] = 1e3 / (out["b200_integrate_ms"] + out["b200_raycast_ms"])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.path.isdir(os.path.join(root, "gpurun_out")):
        json.dump(out, open(os.path.join(root, "gpurun_out", "ref_cuda_timing.json"), "w"))
    assert out["b200_frames_per_s"] > out["ref_frames_per_s"]
    rv.close(); mv.close()
