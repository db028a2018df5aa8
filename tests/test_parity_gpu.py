"""GPU parity tests: the sm_100a kernels, called through the C-ABI, against the CPU oracle.

Bar: bit-exact distances, weights, vertices, normals, hit sample index k and hit voxel.
The north star asks for <= 1e-4 relative on SDF values; bit equality is the stronger statement and is
what is asserted.  Oracle-sized cases finish in seconds on the host; full-size cases (512^3) are
covered by size-independent properties in test_fullsize_gpu.py.
"""
import numpy as np
import pytest

from helpers import assert_bits_equal, random_rigid_pose, random_depth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gpu_util
    return gpu_util


def quarter_intrinsics(cam, s):
    k = cam.k.copy()
    k[:2] *= s
    kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
    return k, kinv


def sphere_sdf(n, physical, trunc, centre, radius):
    """create_sphere_in_TSDF of the reference (TestHelpers.cpp:18-61): analytic SDF clamped to +-trunc."""
    vs = np.asarray(physical, np.float64) / np.asarray(n)
    z, y, x = np.meshgrid(*(np.arange(m) + 0.5 for m in (n[2], n[1], n[0])), indexing="ij")
    d = np.sqrt((x * vs[0] - centre[0]) ** 2 + (y * vs[1] - centre[1]) ** 2 + (z * vs[2] - centre[2]) ** 2) - radius
    return np.clip(d, -trunc, trunc).astype(np.float32).reshape(-1)


# ----------------------------------------------------------------------------------- integrate
@pytest.mark.parametrize("n", [(64, 64, 64), (128, 128, 128), (36, 20, 28), (33, 17, 9), (4, 4, 4), (1, 1, 1), (130, 3, 5)])
def test_integrate_matches_oracle(G, n):
    from oracle import oracle
    from tsdf_b200 import scenes
    rng = np.random.default_rng(sum(n))
    dv = G.DeviceVolume(n, (3000, 3000, 3000))
    ov = oracle.OracleVolume(n, (3000, 3000, 3000))
    for f in range(3):
        cam = scenes.orbit_camera(f * 5 + 1, 16)
        k, kinv = quarter_intrinsics(cam, 0.5)
        depth = scenes.render_depth(cam, 320, 240)
        nu = dv.integrate(depth, cam.inv_pose, k, kinv)
        no = ov.integrate(depth, cam.inv_pose, k, kinv)
        assert nu == no
    assert_bits_equal(dv.dist.cpu().numpy(), ov.dist, "dist")
    assert_bits_equal(dv.weight.cpu().numpy(), ov.weight, "weight")


@pytest.mark.parametrize("variant", [1, 2])
def test_integrate_staging_variants_match_oracle(G, variant):
    """The TMA-staged kernel (3-D tensor maps, cp.async.bulk.tensor boxes of 128 x 4 x 1 voxels, mbarriers; boxes that hang
    over the volume's edge are clipped by the hardware) and the two-pass work-list form, on a ragged volume, against the
    oracle bit for bit."""
    from oracle import oracle
    from tsdf_b200 import scenes
    n = (132, 70, 37)
    dv = G.DeviceVolume(n, (3000, 2200, 1400))
    ov = oracle.OracleVolume(n, (3000, 2200, 1400))
    G.lib.tsdf_b200_debug_integrate_variant(variant)
    try:
        for f in range(3):
            cam = scenes.orbit_camera(f * 5 + 1, 16)
            k, kinv = quarter_intrinsics(cam, 0.5)
            depth = scenes.render_depth(cam, 320, 240)
            assert dv.integrate(depth, cam.inv_pose, k, kinv) == ov.integrate(depth, cam.inv_pose, k, kinv)
    finally:
        G.lib.tsdf_b200_debug_integrate_variant(0)
    assert float(ov.weight.sum()) > 1000
    assert_bits_equal(dv.dist.cpu().numpy(), ov.dist, "dist")
    assert_bits_equal(dv.weight.cpu().numpy(), ov.weight, "weight")


def test_integrate_fast_path_equals_general_kernel(G):
    """The rigid-camera kernel (approximate reciprocal + certainty test, exact fallback) against the general
    kernel on poses chosen to put many projections near rounding boundaries: a fronto-parallel camera whose
    principal point and focal length make voxel centres project onto half-integer pixel coordinates."""
    from oracle import oracle
    from tsdf_b200 import scenes
    rng = np.random.default_rng(2024)
    n = (64, 64, 32)
    for trial in range(4):
        if trial == 0:
            cam = scenes.PinholeCamera(64.0, 64.0, 32.5, 32.5)      # voxel (46.875 mm) at z=3000 -> exactly 1 px apart
            cam.move_to(1500.0 + 23.4375, 1500.0 + 23.4375, -3000.0 + 23.4375)
            w, h = 65, 65
            k, kinv = cam.k, cam.kinv
        else:
            cam = random_rigid_pose(rng)
            k, kinv = quarter_intrinsics(cam, 0.25)
            w, h = 160, 120
        depth = random_depth(rng, w, h, lo=2000, hi=6000, holes=0.05)
        fast = G.DeviceVolume(n, (3000,) * 3)
        gen = G.DeviceVolume(n, (3000,) * 3)
        ov = oracle.OracleVolume(n, (3000,) * 3)
        for rep in range(2):
            nf = fast.integrate(depth, cam.inv_pose, k, kinv)
            G.lib.tsdf_b200_debug_force_generic_integrate(1)
            try:
                ng = gen.integrate(depth, cam.inv_pose, k, kinv)
            finally:
                G.lib.tsdf_b200_debug_force_generic_integrate(0)
            no = ov.integrate(depth, cam.inv_pose, k, kinv)
            assert nf == ng == no
        assert_bits_equal(fast.dist.cpu().numpy(), ov.dist, f"fast dist trial {trial}")
        assert_bits_equal(gen.dist.cpu().numpy(), ov.dist, f"general dist trial {trial}")
        assert_bits_equal(fast.weight.cpu().numpy(), ov.weight, f"fast weight trial {trial}")


def test_integrate_config1_fixed_pose(G):
    """BASELINE config 1 geometry: 128^3, 640x480, fixed pose, identical frames (3 of the 10)."""
    from oracle import oracle
    from tsdf_b200 import scenes
    cam = scenes.fixed_pose_camera()
    depth = scenes.render_depth(cam)
    dv = G.DeviceVolume((128,) * 3, (3000,) * 3)
    ov = oracle.OracleVolume((128,) * 3, (3000,) * 3)
    for _ in range(3):
        assert dv.integrate(depth, cam.inv_pose, cam.k, cam.kinv) == ov.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
    assert_bits_equal(dv.dist.cpu().numpy(), ov.dist, "dist")
    assert_bits_equal(dv.weight.cpu().numpy(), ov.weight, "weight")


def test_integrate_random_poses_and_depth(G):
    from oracle import oracle
    rng = np.random.default_rng(1234)
    n = (72, 56, 40)
    dv = G.DeviceVolume(n, (2500, 3000, 2000), offset=(100.0, -50.0, 25.5))
    ov = oracle.OracleVolume(n, (2500, 3000, 2000))
    ov.offset[:] = (100.0, -50.0, 25.5); ov.clear()
    for f in range(6):
        cam = random_rigid_pose(rng, centre=(1350, 1450, 1025))
        k, kinv = quarter_intrinsics(cam, 0.25)
        depth = random_depth(rng, 160, 120)
        assert dv.integrate(depth, cam.inv_pose, k, kinv) == ov.integrate(depth, cam.inv_pose, k, kinv)
    assert_bits_equal(dv.dist.cpu().numpy(), ov.dist, "dist")
    assert_bits_equal(dv.weight.cpu().numpy(), ov.weight, "weight")


def test_integrate_general_matrices(G):
    """Non-affine inverse pose (w != 1), skewed K and a K^-1 whose third row is not (0,0,1), camera
    inside the volume (voxels behind the camera still project, TSDFVolume.cu has no cull)."""
    from oracle import oracle
    rng = np.random.default_rng(99)
    n = (40, 40, 40)
    dv = G.DeviceVolume(n, (3000,) * 3)
    ov = oracle.OracleVolume(n, (3000,) * 3)
    for f in range(4):
        cam = random_rigid_pose(rng, radius=(100.0, 1400.0))
        ip = cam.inv_pose.copy()
        ip[3] = (1e-5, -2e-5, 3e-5, 1.01)
        k = cam.k.copy(); k[:2] *= 0.25; k[0, 1] = 0.7; k[2] = (1e-4, -1e-4, 1.0)
        kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
        depth = random_depth(rng, 160, 120, lo=1, hi=3000, holes=0.3)
        assert dv.integrate(depth, ip, k, kinv) == ov.integrate(depth, ip, k, kinv)
    assert_bits_equal(dv.dist.cpu().numpy(), ov.dist, "dist")
    assert_bits_equal(dv.weight.cpu().numpy(), ov.weight, "weight")


def test_integrate_degenerate_inputs(G):
    """All-zero depth (nothing fused), camera plane cutting the volume (img.z = 0 -> NaN/inf pixel
    coordinates; NaN converts to pixel 0 on the GPU), maximum depth value."""
    from oracle import oracle
    from tsdf_b200 import scenes
    n = (32, 32, 32)
    dv = G.DeviceVolume(n, (3000,) * 3)
    ov = oracle.OracleVolume(n, (3000,) * 3)
    cam = scenes.PinholeCamera()
    cam.move_to(1500.0, 1500.0, 1500.0 + 0.5 * 93.75)      # camera z exactly on a voxel-centre plane
    k, kinv = quarter_intrinsics(cam, 0.25)
    zero = np.zeros((120, 160), np.uint16)
    assert dv.integrate(zero, cam.inv_pose, k, kinv) == 0 == ov.integrate(zero, cam.inv_pose, k, kinv)
    full = np.full((120, 160), 65535, np.uint16)
    assert dv.integrate(full, cam.inv_pose, k, kinv) == ov.integrate(full, cam.inv_pose, k, kinv)
    # principal point at pixel (0,0) so the NaN -> 0 conversion lands inside the image
    k2 = k.copy(); k2[0, 2] = 0; k2[1, 2] = 0
    kinv2 = np.linalg.inv(k2.astype(np.float64)).astype(np.float32)
    cam.move_to(1500.0 + 0.5 * 93.75, 1500.0 + 0.5 * 93.75, 1500.0 + 0.5 * 93.75)
    d3 = np.full((120, 160), 1000, np.uint16)
    assert dv.integrate(d3, cam.inv_pose, k2, kinv2) == ov.integrate(d3, cam.inv_pose, k2, kinv2)
    assert_bits_equal(dv.dist.cpu().numpy(), ov.dist, "dist")
    assert_bits_equal(dv.weight.cpu().numpy(), ov.weight, "weight")


def test_integrate_deformation_array_path(G):
    """Stored deformation nodes (the reference's only path) == analytic grid, then a perturbed field."""
    import torch
    from oracle import oracle
    from tsdf_b200 import scenes
    n = (48, 40, 32)
    cam = scenes.orbit_camera(2, 16)
    k, kinv = quarter_intrinsics(cam, 0.25)
    depth = scenes.render_depth(cam, 160, 120)
    a = G.DeviceVolume(n, (3000,) * 3, with_deformation=True)
    b = G.DeviceVolume(n, (3000,) * 3)
    ov = oracle.OracleVolume(n, (3000,) * 3, with_deformation=True)
    assert_bits_equal(a.deform.cpu().numpy(), ov.deform, "deformation grid")
    assert a.integrate(depth, cam.inv_pose, k, kinv) == b.integrate(depth, cam.inv_pose, k, kinv)
    ov.integrate(depth, cam.inv_pose, k, kinv)
    assert_bits_equal(a.dist.cpu().numpy(), b.dist.cpu().numpy(), "dist array-vs-analytic")
    assert_bits_equal(a.dist.cpu().numpy(), ov.dist, "dist")
    rng = np.random.default_rng(5)
    ov.deform[:] = ov.deform + rng.normal(0, 20, ov.deform.size).astype(np.float32)
    a.deform.copy_(torch.from_numpy(ov.deform))
    assert a.integrate(depth, cam.inv_pose, k, kinv) == ov.integrate(depth, cam.inv_pose, k, kinv)
    assert_bits_equal(a.dist.cpu().numpy(), ov.dist, "dist deformed")
    assert_bits_equal(a.weight.cpu().numpy(), ov.weight, "weight deformed")


def test_integrate_z_ranges_compose(G):
    """Z-slab decomposition (the multi-GPU sharding unit): slabs fused separately == whole volume."""
    from tsdf_b200 import scenes
    n = (64, 48, 40)
    cam = scenes.orbit_camera(3, 16)
    k, kinv = quarter_intrinsics(cam, 0.25)
    depth = scenes.render_depth(cam, 160, 120)
    whole = G.DeviceVolume(n, (3000,) * 3)
    parts = G.DeviceVolume(n, (3000,) * 3)
    total = whole.integrate(depth, cam.inv_pose, k, kinv)
    got = sum(parts.integrate(depth, cam.inv_pose, k, kinv, z0, z1) for z0, z1 in ((0, 13), (13, 14), (14, 40)))
    assert got == total
    assert_bits_equal(whole.dist.cpu().numpy(), parts.dist.cpu().numpy(), "dist")
    assert_bits_equal(whole.weight.cpu().numpy(), parts.weight.cpu().numpy(), "weight")


def test_clear_and_golden_fixture(G):
    import hashlib, json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "t_100_2000_50.json")))
    dv = G.DeviceVolume(g["size"], g["physical"], with_deformation=True)
    sha = lambda t: hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()
    assert int(np.float32(dv.trunc).view(np.uint32)) == g["trunc_bits"]
    assert sha(dv.dist) == g["dist_sha256"] and sha(dv.weight) == g["weight_sha256"] and sha(dv.deform) == g["deform_sha256"]


# ------------------------------------------------------------------------------------- raycast
def check_raycast(G, dv, ov, w, h, pose, kinv, what):
    from oracle import oracle
    Vo, No, ko, so = ov.raycast(w, h, pose, kinv)
    results = {}
    for skip in (False, True):
        for fastdiv in (False, True):
            V, N, kh, ns = dv.raycast(w, h, pose, kinv, skip=skip, fastdiv=fastdiv)
            tag = f"{what} skip={skip} fastdiv={fastdiv}"
            assert np.array_equal(kh, ko), tag + ": hit sample index k differs"
            assert_bits_equal(V, Vo, tag + ": vertices")
            assert_bits_equal(N, No, tag + ": normals")
            hv = oracle.hit_voxels(V, ov.offset, ov.voxel, ov.size[0], ov.size[1])
            ho = oracle.hit_voxels(Vo, ov.offset, ov.voxel, ov.size[0], ov.size[1])
            assert np.array_equal(hv, ho), tag + ": hit voxel index differs"
            if not skip:
                assert ns == so, tag + ": sample count differs from the reference march"
            else:
                assert ns <= so
            results[(skip, fastdiv)] = ns
    return results, int((ko >= 0).sum()), so


def test_raycast_sphere_reference_poses(G):
    """The two live camera placements of Test_TSDF_RayCast.cpp (:416-431, :568-581): sphere SDF in a
    volume, camera at (450,150,150)->(150,150,150) and (-150,150,450)->(150,150,150), scaled x10."""
    from oracle import oracle
    from tsdf_b200 import scenes
    n = (96, 96, 96)
    dv = G.DeviceVolume(n, (3000,) * 3)
    ov = oracle.OracleVolume(n, (3000,) * 3)
    sdf = sphere_sdf(n, (3000,) * 3, ov.trunc, (1500, 1500, 1500), 800)
    ov.dist[:] = sdf
    dv.upload_dist(sdf)
    for pos in ((4500, 1500, 1500), (-1500, 1500, 4500), (1500, 1500, -2500), (1500, 5200, 1500)):
        cam = scenes.PinholeCamera()
        cam.move_to(*pos)
        cam.look_at(1500, 1500, 1500)
        k, kinv = quarter_intrinsics(cam, 0.5)
        res, hits, so = check_raycast(G, dv, ov, 320, 240, cam.pose, kinv, f"sphere from {pos}")
        assert hits > 1000
        assert res[(True, True)] < 0.5 * so      # skipping must actually skip


def test_raycast_after_integration(G):
    """Integrate the synthetic scene from several poses, raycast from each (config-2 style, small)."""
    from oracle import oracle
    from tsdf_b200 import scenes
    n = (80, 80, 80)
    dv = G.DeviceVolume(n, (3000,) * 3)
    ov = oracle.OracleVolume(n, (3000,) * 3)
    for f in range(0, 16, 3):
        cam = scenes.orbit_camera(f, 16)
        k, kinv = quarter_intrinsics(cam, 0.5)
        depth = scenes.render_depth(cam, 320, 240)
        dv.integrate(depth, cam.inv_pose, k, kinv)
        ov.integrate(depth, cam.inv_pose, k, kinv)
        check_raycast(G, dv, ov, 320, 240, cam.pose, kinv, f"orbit frame {f}")


def test_raycast_camera_inside_and_axis_aligned(G):
    """Origin inside the AABB (near_t = 0, NaN-tolerant far_t chain, GPURaycaster.cu:202-233), rays with
    exactly-zero direction components, offset volume, anisotropic voxels."""
    from oracle import oracle
    from tsdf_b200 import scenes
    n = (64, 48, 40)
    phys = (2000, 2400, 1600)
    off = (250.0, -100.0, 40.0)
    dv = G.DeviceVolume(n, phys, offset=off)
    ov = oracle.OracleVolume(n, phys)
    ov.offset[:] = off
    sdf = sphere_sdf(n, phys, ov.trunc, (1000, 1200, 800), 500)
    ov.dist[:] = sdf
    dv.upload_dist(sdf)
    cam = scenes.PinholeCamera(100.0, 100.0, 80.0, 60.0)     # principal point on a pixel: centre ray has dir (0,0,1)
    cam.move_to(1250.0, 1100.0, 100.0)                        # inside the (offset) volume
    check_raycast(G, dv, ov, 160, 120, cam.pose, cam.kinv, "inside, axis aligned")
    cam.move_to(1250.0, 1100.0, -900.0)
    check_raycast(G, dv, ov, 160, 120, cam.pose, cam.kinv, "outside, axis aligned")
    cam.move_to(5000.0, 5000.0, -900.0)                       # every ray misses the volume
    res, hits, so = check_raycast(G, dv, ov, 160, 120, cam.pose, cam.kinv, "all miss")
    assert hits == 0 and so == 0


def test_raycast_low_edge_extrapolation_and_negative_volume(G):
    """Random (non-SDF) data with sign changes everywhere, including the first half-voxel layer where the
    reference extrapolates (GPURaycaster.cu:87-99): skipping must never hide a hit."""
    from oracle import oracle
    from tsdf_b200 import scenes
    rng = np.random.default_rng(11)
    n = (40, 40, 40)
    dv = G.DeviceVolume(n, (3000,) * 3)
    ov = oracle.OracleVolume(n, (3000,) * 3)
    data = np.full(40 ** 3, ov.trunc, np.float32)
    grid = data.reshape(40, 40, 40)
    grid[:, :, 1] = 100 * ov.trunc            # extrapolation in the x=0 half voxel goes negative
    grid[20:, 5:30, 10:12] = -ov.trunc * rng.random((20, 25, 2)).astype(np.float32)
    grid[3, 3, 3] = np.nan
    grid[30, 30, 30] = 0.0
    ov.dist[:] = data
    dv.upload_dist(data)
    for pos in ((-800, 1500, 1400), (1500, 1500, -900), (3900, 1700, 3900), (1500, -2000, 1500)):
        cam = scenes.PinholeCamera()
        cam.move_to(*pos)
        cam.look_at(1500, 1500, 1500)
        k, kinv = quarter_intrinsics(cam, 0.25)
        check_raycast(G, dv, ov, 160, 120, cam.pose, kinv, f"synthetic data from {pos}")


def test_raycast_sample_cap(G):
    """Rays longer than 4402 samples end as misses at the cap (GPURaycaster.cu:369); a surface beyond the
    cap is NOT found, one just before it is."""
    from oracle import oracle
    from tsdf_b200 import scenes
    n = (512, 8, 8)
    phys = (3000, 46.875, 46.875)          # cubic 5.859375 mm voxels, a 3000 mm long bar
    dv = G.DeviceVolume(n, phys)
    ov = oracle.OracleVolume(n, phys)
    step = np.float64(np.float32(np.float64(ov.trunc) * 0.05))
    assert 4500 * step < 3000
    for x_wall in (4300 * step, 4500 * step):
        xc = (np.arange(512) + 0.5) * ov.voxel[0]
        data = np.broadcast_to(np.clip(x_wall - xc, -ov.trunc, ov.trunc).astype(np.float32), (8, 8, 512)).copy()
        ov.dist[:] = data.reshape(-1)
        dv.upload_dist(data.reshape(-1))
        cam = scenes.PinholeCamera(20000.0, 20000.0, 16.0, 12.0)    # near-parallel rays down the bar
        cam.move_to(-1.0, 23.4375, 23.4375)
        cam.look_at(3000.0, 23.4375, 23.4375)
        res, hits, so = check_raycast(G, dv, ov, 32, 24, cam.pose, cam.kinv, f"cap wall@{x_wall:.0f}")
        if x_wall > 4402 * step:
            assert hits == 0
        else:
            assert hits == 32 * 24


def test_raycast_random_volumes_stress(G):
    """Randomised stress of the exact skipping logic: blobby random fields with many sign changes, thin shells,
    isolated non-positive voxels and NaNs, random cameras (inside and outside), anisotropic voxels, offsets."""
    from oracle import oracle
    from tsdf_b200 import scenes
    rng = np.random.default_rng(4242)
    for trial in range(6):
        n = tuple(int(x) for x in rng.integers(17, 72, size=3))
        phys = tuple(float(x) for x in rng.uniform(1500, 3500, size=3))
        off = tuple(float(x) for x in rng.uniform(-200, 200, size=3))
        dv = G.DeviceVolume(n, phys, offset=off)
        ov = oracle.OracleVolume(n, phys)
        ov.offset[:] = off
        z, y, x = np.meshgrid(*(np.linspace(0, 1, m) for m in (n[2], n[1], n[0])), indexing="ij")
        f = np.zeros(x.shape)
        for _ in range(4):
            c = rng.uniform(0.1, 0.9, size=3); r = rng.uniform(0.1, 0.35)
            f = np.maximum(f, r - np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2))
        data = np.where(f > 0, -f * 4000, ov.trunc).astype(np.float32)
        data = np.clip(data, -ov.trunc, ov.trunc)
        sel = rng.random(data.shape) < 0.0005
        data[sel] = rng.choice(np.array([0.0, -1.0, np.nan, 1e-6, 1e9], np.float32), size=int(sel.sum()))
        if trial % 2:
            data[:, :, :2] = rng.uniform(-1, 1, size=data[:, :, :2].shape).astype(np.float32) * ov.trunc   # noisy low edge
        ov.dist[:] = data.reshape(-1)
        dv.upload_dist(data.reshape(-1))
        centre = np.asarray(off) + 0.5 * np.asarray(phys)
        for view in range(3):
            cam = random_rigid_pose(rng, centre=centre, radius=(300.0, 800.0) if view == 2 else (2000.0, 5000.0))
            k, kinv = quarter_intrinsics(cam, 0.25)
            check_raycast(G, dv, ov, 160, 120, cam.pose, kinv, f"stress trial {trial} view {view}")


def test_raycast_z_sharded_equals_whole(G):
    """The multi-GPU decomposition on one device: slab copies with a halo plane, per-slab occupancy, key
    min-reduction, resolve == the oracle's undivided march (vertices, hit index) for 1..5 shards."""
    from oracle import oracle
    from tsdf_b200 import scenes
    n = (64, 56, 72)
    dv = G.DeviceVolume(n, (3000,) * 3)
    ov = oracle.OracleVolume(n, (3000,) * 3)
    for f in (0, 5, 11):
        cam = scenes.orbit_camera(f, 16)
        k, kinv = quarter_intrinsics(cam, 0.5)
        depth = scenes.render_depth(cam, 320, 240)
        dv.integrate(depth, cam.inv_pose, k, kinv)
        ov.integrate(depth, cam.inv_pose, k, kinv)
    for f in (2, 6, 13):
        cam = scenes.orbit_camera(f, 16)
        k, kinv = quarter_intrinsics(cam, 0.5)
        Vo, No, ko, so = ov.raycast(320, 240, cam.pose, kinv)
        for world in (1, 2, 3, 5):
            for skip in (False, True):
                V, kh, ns = G.raycast_sharded_on_one_gpu(dv, 320, 240, cam.pose, kinv, world, skip=skip)
                assert np.array_equal(kh, ko), f"frame {f} world {world} skip {skip}: hit index"
                assert_bits_equal(V, Vo, f"frame {f} world {world} skip {skip}: vertices")
                if not skip:
                    assert ns <= 2 * so + 320 * 240 * 16 * world      # shards re-evaluate only samples near slab seams


def test_normals_kernel(G):
    import torch
    from oracle import oracle
    rng = np.random.default_rng(3)
    V = rng.normal(0, 1000, (48 * 64, 3)).astype(np.float32)
    V[rng.random(48 * 64) < 0.2] = np.nan
    V[100] = V[101]          # zero-length difference -> 0/0
    Vd = G.dev(V.reshape(-1))
    Nd = torch.empty_like(Vd)
    G.check(G.lib.tsdf_b200_normals(64, 48, G.ptr(Vd), G.ptr(Nd), None))
    torch.cuda.synchronize()
    assert_bits_equal(Nd.cpu().numpy(), oracle.normals(64, 48, V), "normals")


def test_reciprocal_division_selftest(G):
    import ctypes as C
    for b in (3000 / 512, 3000 / 128, 3000 / 1024, 20.0, 3000 / 200, 2500 / 72, 1.0, 0.37):
        bad = C.c_ulonglong(123)
        G.check(G.lib.tsdf_b200_selftest_division(np.float32(b), C.byref(bad)))
        assert bad.value == 0, f"fdiv_recip differs from IEEE division for {bad.value} numerators at b={b}"


# ------------------------------------------------------------------------------ level-2 object
def test_volume_object_end_to_end(G, tmp_path):
    """The host-buffer API kinfu reaches through TSDFVolume: integrate x N, raycast, save, load."""
    import json, os
    from oracle import oracle
    from tsdf_b200 import Volume, scenes
    n = (64, 64, 64)
    vol = Volume(n, (3000.0,) * 3)
    ov = oracle.OracleVolume(n, (3000,) * 3)
    assert vol.trunc == ov.trunc and vol.max_weight == 15.0
    for f in (0, 2, 5):
        cam = scenes.orbit_camera(f, 12)
        depth = scenes.render_depth(cam)
        vol.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
        nu = ov.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
        assert vol.stats()[0] == nu
    d, w = vol.read()
    assert_bits_equal(d, ov.dist, "dist"); assert_bits_equal(w, ov.weight, "weight")
    V, N = vol.raycast(640, 480, cam.pose, cam.kinv)
    Vo, No, ko, so = ov.raycast(640, 480, cam.pose, cam.kinv)
    assert_bits_equal(V, Vo, "vertices"); assert_bits_equal(N, No, "normals")
    assert 0 < vol.stats()[1] < so
    vol.set_skipping(False)
    V2, N2 = vol.raycast(640, 480, cam.pose, cam.kinv)
    assert_bits_equal(V2, Vo, "vertices (no skipping)")
    assert vol.stats()[1] == so
    # save -> load round trip keeps every bit, and a loaded volume integrates through its stored field
    path = str(tmp_path / "v.tsdf")
    vol.save(path)
    assert os.path.getsize(path) == 68 + 64 ** 3 * 35
    vol2 = Volume.load(path)
    d2, w2 = vol2.read()
    assert_bits_equal(d2, d, "dist after load"); assert_bits_equal(w2, w, "weight after load")
    cam = scenes.orbit_camera(7, 12)
    depth = scenes.render_depth(cam)
    vol2.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
    ov.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
    d2, w2 = vol2.read()
    assert_bits_equal(d2, ov.dist, "dist after load+integrate")
    V3, _ = vol2.raycast(640, 480, cam.pose, cam.kinv)
    assert_bits_equal(V3, ov.raycast(640, 480, cam.pose, cam.kinv)[0], "vertices after load")
    vol.close(); vol2.close()


def test_volume_save_matches_reference_file_layout(G, tmp_path):
    """A freshly constructed 100^3/2000mm volume with offset (50,50,50) saves the reference fixture's
    bytes (TestData/t_100_2000_50.tsdf) in every section except the uninitialised colours."""
    import hashlib, json, os
    from tsdf_b200 import Volume
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "t_100_2000_50.json")))
    vol = Volume(g["size"], g["physical"])
    vol.set_offset(*g["offset"])
    path = str(tmp_path / "t.tsdf")
    vol.save(path)
    raw = open(path, "rb").read()
    n = 100 ** 3
    assert len(raw) == g["file_bytes"]
    assert raw[:68].hex() == g["header_hex"]
    sha = lambda b: hashlib.sha256(b).hexdigest()
    assert sha(raw[68:68 + 4 * n]) == g["dist_sha256"]
    assert sha(raw[68 + 4 * n:68 + 8 * n]) == g["weight_sha256"]
    assert sha(raw[68 + 11 * n:]) == g["deform_sha256"]
    vol.close()


def test_volume_offset_and_clear_semantics(G):
    """offset() after construction is added on top of the stored grid (TSDFVolume.cu:343); clear()
    re-bakes the current offset into the grid (:841), so it is then applied twice."""
    from oracle import oracle
    from tsdf_b200 import Volume, scenes
    n = (40, 40, 40)
    vol = Volume(n, (3000.0,) * 3)
    ov = oracle.OracleVolume(n, (3000,) * 3, with_deformation=True)
    cam = scenes.orbit_camera(1, 12)
    depth = scenes.render_depth(cam)
    for step in range(2):
        vol.set_offset(120.0, -60.0, 33.0); ov.offset[:] = (120.0, -60.0, 33.0)
        vol.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
        ov.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
        d, w = vol.read()
        assert_bits_equal(d, ov.dist, f"dist step {step}"); assert_bits_equal(w, ov.weight, f"weight step {step}")
        V, N = vol.raycast(320, 240, cam.pose, cam.kinv)
        assert_bits_equal(V, ov.raycast(320, 240, cam.pose, cam.kinv)[0], f"vertices step {step}")
        vol.clear(); ov.clear()
    vol.close()


@pytest.mark.parametrize("world,slab", [(2, 16), (3, 8), (8, 16)])
def test_interleaved_slabs_equal_whole(G, world, slab):
    """Interleaved Z-slabs (the multi-GPU layout of ShardedEngine): every emulated rank integrates its own slabs (+ halo
    planes) into its own arrays and its own whole-volume occupancy grid, marches all rays over the cells it owns; the
    min over the ranks' keys, resolved, must equal the single-volume raycast bit for bit, and the ranks' planes must equal
    the single volume's planes."""
    import torch
    from tsdf_b200 import scenes, sharded
    n, phys = (96, 80, 112), (3000.0, 2500.0, 3000.0)
    whole = sharded.ShardedEngine(n, phys)
    ranks = [sharded.ShardedEngine(n, phys, rank=r, world=world, layout="interleaved", slab=slab) for r in range(world)]
    cams = [scenes.orbit_camera(i, 12) for i in (0, 2, 5)]
    counts = []
    for cam in cams:
        k = cam.k.copy(); k[:2] *= 0.5
        cam.k = k
        cam.kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
        depth = torch.from_numpy(scenes.render_depth(cam, 320, 240)).cuda()
        total = whole.integrate(depth, cam, count=True)
        assert sum(e.integrate(depth, cam, count=True) for e in ranks) == total
        counts.append(total)
    assert sum(counts) > 0
    wd = whole.dist.view(n[2], -1)
    for e in ranks:
        ld = e.dist.view(-1, n[0] * n[1])
        for j, (z0, z1) in enumerate(e.slabs):
            stored = (z1 - z0) + (1 if z1 < n[2] else 0)
            assert torch.equal(ld[j * (slab + 1): j * (slab + 1) + stored].view(torch.int32), wd[z0:z0 + stored].view(torch.int32))
    cam = cams[1]
    whole.raycast(320, 240, cam)
    keys = None
    for e in ranks:
        k = e.march(320, 240, cam).clone()
        keys = k if keys is None else torch.minimum(keys, k)
    ranks[0].keys.copy_(keys)
    ranks[0].resolve(320, 240, cam)
    torch.cuda.synchronize()
    assert_bits_equal(ranks[0].vertices.cpu().numpy(), whole.vertices.cpu().numpy(), "vertices")
    assert_bits_equal(ranks[0].normals.cpu().numpy(), whole.normals.cpu().numpy(), "normals")
    assert int((~torch.isnan(whole.vertices.view(-1, 3)[:, 0])).sum()) > 5000


@pytest.mark.parametrize("world,slab,n", [(2, 16, (96, 80, 112)), (3, 8, (96, 80, 112)), (8, 16, (96, 80, 112)), (4, 8, (90, 77, 61))])
def test_replica_image_sharding_equals_whole(G, world, slab, n):
    """layout="replica" (the multi-GPU path of bench.py --gpus N), ranks emulated on one GPU: every rank integrates its
    slabs, the brick flags are max-merged (the all-reduce), every rank pushes its surface bricks into every rank's replica
    (pre-filled with NaN here: a single voxel read outside the pushed set would poison a sample), marches its own pixel
    tiles and stores them into every rank's vertex map.  Each rank's vertex and normal maps must equal the single-volume
    raycast bit for bit."""
    import ctypes as C
    import torch
    from tsdf_b200 import scenes, sharded
    from tsdf_b200.capi import lib, check
    phys = (3000.0, 2500.0, 3000.0)
    whole = sharded.ShardedEngine(n, phys)
    ranks = [sharded.ShardedEngine(n, phys, rank=r, world=world, layout="replica", slab=slab) for r in range(world)]
    for e in ranks:
        check(lib.tsdf_b200_fill_f32(C.c_void_p(e.replica), n[0] * n[1] * n[2], float("nan"), e.stream))
        e.connect(320, 240, peers=ranks)
    assert all(e.replicas is not None and len(e.vmaps) == world for e in ranks)
    for frame in (0, 2, 5, 7):
        cam = scenes.orbit_camera(frame, 12)
        k = cam.k.copy(); k[:2] *= 0.5
        cam.k = k
        cam.kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
        depth = torch.from_numpy(scenes.render_depth(cam, 320, 240)).cuda()
        total = whole.integrate(depth, cam, count=True)
        assert sum(e.integrate(depth, cam, count=True) for e in ranks) == total
        whole.raycast(320, 240, cam)
        merged = ranks[0].flags().clone()
        for e in ranks[1:]:
            merged = torch.maximum(merged, e.flags())
        for e in ranks:
            e.flags().copy_(merged)
        pushed = sum(e.push(count=True) for e in ranks)
        for e in ranks:
            check(lib.tsdf_b200_fill_f32(C.c_void_p(e.vmap), 320 * 240 * 3, -1.0, e.stream))
        torch.cuda.synchronize()
        for e in ranks:
            e.march_tiles(320, 240, cam)
        for e in ranks:
            e.finish(320, 240)
        torch.cuda.synchronize()
        flagged = int(merged.sum().item())
        nb = [(x + 7) // 8 for x in n]
        low_face = nb[0] * nb[1] * nb[2] - (nb[0] - 1) * (nb[1] - 1) * (nb[2] - 1)      # always pushed (low-edge layer)
        assert flagged <= pushed <= 27 * flagged + low_face and pushed <= merged.numel()
        for e in ranks:
            assert_bits_equal(e.vertices.cpu().numpy(), whole.vertices.cpu().numpy(), f"vertices of rank {e.rank}, frame {frame}")
            assert_bits_equal(e.normals.cpu().numpy(), whole.normals.cpu().numpy(), f"normals of rank {e.rank}, frame {frame}")
    assert int((~torch.isnan(whole.vertices.view(-1, 3)[:, 0])).sum()) > 5000
    # the merged flags equal the single volume's flags: nothing a rank cannot see decides a brick
    assert torch.equal(merged, whole.occ[:merged.numel()])
    for e in ranks:
        e.close()
    whole.close()


def _load_engine_planes(e, planes, n, slab):
    """Writes a whole-volume distance array (nz, ny*nx) into a replica-layout engine's slabs (+ halo planes)."""
    ld = e.dist.view(-1, n[0] * n[1])
    for j, (z0, z1) in enumerate(e.slabs):
        stored = (z1 - z0) + (1 if z1 < n[2] else 0)
        ld[j * (slab + 1): j * (slab + 1) + stored].copy_(planes[z0:z0 + stored])


def test_replica_low_face_extrapolation_and_clear(G):
    """Advisor findings (round 1) on layout="replica".  (1) Voxel layer 0 of each axis is always evaluated and
    EXTRAPOLATED by the reference (u in [-0.5, 0)): two positive-band corners c0 = 0.2*trunc, c1 = trunc give a sample <= 0,
    i.e. a hit, inside a brick that is not flagged — the push must publish low-face bricks regardless of flags (replicas
    are NaN-poisoned here, so a brick that was not pushed shows).  (2) clear() must reset the replica: after a clear the
    bricks pushed earlier are no longer flagged nor pushed, but the low-edge layer of them is still read."""
    import ctypes as C
    import torch
    from tsdf_b200 import scenes, sharded
    from tsdf_b200.capi import lib, check
    world, slab, n, phys = 2, 16, (64, 64, 64), (3000.0, 3000.0, 3000.0)
    w, h = 320, 240
    whole = sharded.ShardedEngine(n, phys)
    ranks = [sharded.ShardedEngine(n, phys, rank=r, world=world, layout="replica", slab=slab) for r in range(world)]
    for e in ranks:
        check(lib.tsdf_b200_fill_f32(C.c_void_p(e.replica), n[0] * n[1] * n[2], float("nan"), e.stream))
        e.connect(w, h, peers=ranks)
    t = float(whole.trunc)
    vol = torch.full((n[2], n[1], n[0]), t, dtype=torch.float32, device="cuda")
    vol[:, :, 0] = 0.2 * t            # x layer 0: extrapolates to <= 0 within the first quarter voxel
    vol[:, 0, 8:40] = 0.3 * t         # part of y layer 0
    vol[0, 20:50, :] = 0.25 * t       # part of z layer 0
    planes = vol.view(n[2], -1)
    whole.dist.copy_(vol.view(-1))
    check(lib.tsdf_b200_occupancy_rebuild(C.c_void_p(whole.dist.data_ptr()), *n, whole.trunc, C.c_void_p(whole.occ.data_ptr()), whole.stream))
    for e in ranks:
        _load_engine_planes(e, planes, n, slab)
    nbricks = 8 * 8 * 8
    # every voxel is in the wide positive band: the only flagged bricks are on the low faces (bx, by or bz == 0), whose
    # tight band [0.8, 1.0001] * trunc the low-edge voxels violate — the raycast therefore evaluates there instead of skipping
    flags = whole.occ[:nbricks].view(8, 8, 8)
    assert int(flags[1:, 1:, 1:].sum().item()) == 0 and int(flags.sum().item()) > 0
    for e in ranks:                       # (the ranks' planes were written directly: give them the flags an integrate would have set)
        e.flags().copy_(whole.occ[:nbricks])

    def run(cam):
        whole.raycast(w, h, cam)
        for e in ranks:
            e.push()
        torch.cuda.synchronize()
        for e in ranks:
            e.march_tiles(w, h, cam)
        for e in ranks:
            e.finish(w, h)
        torch.cuda.synchronize()
        for e in ranks:
            assert_bits_equal(e.vertices.cpu().numpy(), whole.vertices.cpu().numpy(), f"vertices of rank {e.rank}")
        return int((~torch.isnan(whole.vertices.view(-1, 3)[:, 0])).sum())

    hits = 0
    for pos in ((-2500.0, 1400.0, 1300.0), (1600.0, -2500.0, 1500.0), (1500.0, 1500.0, -2500.0)):
        cam = scenes.PinholeCamera(295.5, 295.0, 165.5, 117.3)
        cam.move_to(*pos)
        cam.look_at(1500.0, 1500.0, 1500.0)
        hits += run(cam)
    assert hits > 20000                  # the extrapolated low-edge hits exist (single-GPU path == oracle elsewhere)

    # (2) integrate a frame (surface bricks get flagged and pushed), clear, integrate a different view: same as the whole
    for e in ranks:
        e.clear()
    whole.clear()
    cams = [scenes.orbit_camera(i, 12) for i in (1, 7)]
    for cam in cams:
        k = cam.k.copy(); k[:2] *= 0.5
        cam.k = k
        cam.kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
    for phase, cam in enumerate(cams):
        depth = torch.from_numpy(scenes.render_depth(cam, w, h)).cuda()
        whole.integrate(depth, cam)
        for e in ranks:
            e.integrate(depth, cam)
        merged = torch.maximum(ranks[0].flags(), ranks[1].flags()).clone()
        for e in ranks:
            e.flags().copy_(merged)
        assert run(cam) > 3000
        # a different view of the pre-clear surface would hit stale voxels if the replica kept them
        if phase == 0:
            for e in ranks:
                e.clear()
            whole.clear()
    for e in ranks:
        e.close()
    whole.close()


def test_mirrored_vertex_map_in_pinned_memory(G):
    """tsdf_b200_raycast_mirrored: the copy of the vertex map written into pinned host memory during the march (tile-wise
    16-byte stores from the march, single pixels from the continuation) equals the device vertex map bit for bit; and
    the level-2 volume gives the same result for a pinned and for a pageable result buffer."""
    import ctypes as C
    import torch
    from tsdf_b200 import scenes, sharded, Volume
    from tsdf_b200.capi import lib, check, fptr, fvec, colmajor
    n, phys = (96, 80, 112), (3000.0, 2500.0, 3000.0)
    eng = sharded.ShardedEngine(n, phys)
    vol = Volume(n, phys)
    w, h = 320, 240
    for frame in (0, 3, 6):
        cam = scenes.orbit_camera(frame, 12)
        k = cam.k.copy(); k[:2] *= 0.5
        cam.k = k
        cam.kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
        depth = scenes.render_depth(cam, w, h)
        eng.integrate(torch.from_numpy(depth).cuda(), cam)
        check(lib.tsdf_b200_volume_integrate(vol._h, depth.ctypes.data, w, h, fptr(colmajor(cam.inv_pose)), fptr(colmajor(cam.k)),
                                             fptr(colmajor(cam.kinv))))
    eng.raycast(w, h, cam)
    want = eng.vertices.cpu().numpy()
    dvert = torch.empty(w * h * 3, dtype=torch.float32, device="cuda")
    mirror = torch.full((w * h * 3,), -7.0, dtype=torch.float32).pin_memory()
    pose = np.asarray(cam.pose, np.float32)
    smin = eng.offset.copy()
    smax = (eng.offset + eng.physical).astype(np.float32)
    check(lib.tsdf_b200_raycast_mirrored(C.c_void_p(eng.dist.data_ptr()), *n, fptr(eng.voxel), fptr(smin), fptr(smax), eng.trunc,
                                         fptr(fvec(pose[:3, 3])), fptr(colmajor(pose[:3, :3])), fptr(colmajor(cam.kinv)), w, h,
                                         C.c_void_p(eng.table.data_ptr()), C.c_void_p(eng.occ.data_ptr()), C.c_void_p(dvert.data_ptr()),
                                         C.c_void_p(mirror.data_ptr()), None, eng.fastdiv, eng.stream), "raycast_mirrored")
    torch.cuda.synchronize()
    assert_bits_equal(dvert.cpu().numpy(), want, "device vertex map")
    assert_bits_equal(mirror.numpy(), want, "mirrored vertex map")
    # level 2: pinned result buffers take the mirrored path, pageable ones the copy
    outs = []
    for pinned in (True, False):
        hv = torch.empty((h * w, 3), dtype=torch.float32)
        hn = torch.empty((h * w, 3), dtype=torch.float32)
        if pinned:
            hv, hn = hv.pin_memory(), hn.pin_memory()
        check(lib.tsdf_b200_volume_raycast(vol._h, w, h, fptr(colmajor(cam.pose)), fptr(colmajor(cam.kinv)), hv.numpy().ctypes.data,
                                           hn.numpy().ctypes.data))
        outs.append((hv.numpy().copy(), hn.numpy().copy()))
    assert_bits_equal(outs[0][0].reshape(-1), want, "level-2 vertices, pinned buffer")
    assert_bits_equal(outs[1][0].reshape(-1), want, "level-2 vertices, pageable buffer")
    assert_bits_equal(outs[0][1], outs[1][1], "level-2 normals")
    assert int((~np.isnan(want.reshape(-1, 3)[:, 0])).sum()) > 5000
    vol.close()


def test_fused_raycast_normals_and_mirrors(G):
    """tsdf_b200_raycast_fused: vertex and normal maps with the normals computed tile by tile as their
    inputs complete, and both maps mirrored into pinned host memory while the march runs — bit-equal to tsdf_b200_raycast_ex +
    tsdf_b200_normals, with and without the per-tile counters, on images of 8x4-tile and ragged size."""
    import ctypes as C
    import torch
    from tsdf_b200 import scenes, sharded
    from tsdf_b200.capi import lib, check, fptr, fvec, colmajor
    n, phys = (128, 128, 128), (3000.0, 3000.0, 3000.0)
    eng = sharded.ShardedEngine(n, phys)
    for w, h, scale in ((320, 240, 0.5), (640, 480, 1.0), (322, 241, 0.5)):
        eng.clear()
        cams = []
        for frame in (0, 2, 5, 7):
            cam = scenes.orbit_camera(frame, 12)
            k = cam.k.copy(); k[:2] *= scale
            cam.k = k
            cam.kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
            eng.integrate(torch.from_numpy(scenes.render_depth(cam, w, h)).cuda(), cam)
            cams.append(cam)
        for cam in cams[1:]:
            eng.raycast(w, h, cam)
            want_v, want_n = eng.vertices.cpu().numpy(), eng.normals.cpu().numpy()
            pose = np.asarray(cam.pose, np.float32)
            smin = eng.offset.copy()
            smax = (eng.offset + eng.physical).astype(np.float32)
            aligned = w % 8 == 0 and h % 4 == 0
            tiles = torch.zeros(lib.tsdf_b200_raycast_tile_counters(w, h), dtype=torch.int32, device="cuda")
            for with_counters in (True, False):
                for with_mirrors in ((True, False) if aligned else (False,)):
                    dv = torch.full((w * h * 3,), -3.0, dtype=torch.float32, device="cuda")
                    dn = torch.full((w * h * 3,), -3.0, dtype=torch.float32, device="cuda")
                    mv = torch.full((w * h * 3,), -7.0, dtype=torch.float32).pin_memory()
                    mn = torch.full((w * h * 3,), -7.0, dtype=torch.float32).pin_memory()
                    check(lib.tsdf_b200_raycast_fused(
                        C.c_void_p(eng.dist.data_ptr()), *n, fptr(eng.voxel), fptr(smin), fptr(smax), eng.trunc,
                        fptr(fvec(pose[:3, 3])), fptr(colmajor(pose[:3, :3])), fptr(colmajor(cam.kinv)), w, h,
                        C.c_void_p(eng.table.data_ptr()), C.c_void_p(eng.occ.data_ptr()), C.c_void_p(dv.data_ptr()),
                        C.c_void_p(dn.data_ptr()), C.c_void_p(mv.data_ptr()) if with_mirrors else None,
                        C.c_void_p(mn.data_ptr()) if with_mirrors else None,
                        C.c_void_p(tiles.data_ptr()) if with_counters else None, None, eng.fastdiv, eng.stream), "raycast_fused")
                    torch.cuda.synchronize()
                    tag = f"{w}x{h} counters={with_counters} mirrors={with_mirrors}"
                    assert_bits_equal(dv.cpu().numpy(), want_v, tag + ": device vertices")
                    assert_bits_equal(dn.cpu().numpy(), want_n, tag + ": device normals")
                    if with_mirrors:
                        assert_bits_equal(mv.numpy(), want_v, tag + ": mirrored vertices")
                        assert_bits_equal(mn.numpy(), want_n, tag + ": mirrored normals")
                    assert int(tiles.abs().sum().item()) == 0, tag + ": tile counters not left zero"
            assert int((~np.isnan(want_v.reshape(-1, 3)[:, 0])).sum()) > 3000
    eng.close()


def test_brick_distance_grid_equals_chessboard_transform(G):
    """The brick distance grid every raycast derives from the flags (distance_bits_kernel: dilations of 64-bit rows, one
    launch) equals the capped Chebyshev distance transform computed by scipy, for grids of one and two words per row, rows
    that are not a multiple of 16 bricks, a single slice, and more slices than one block produces."""
    import ctypes as C
    import torch
    from scipy import ndimage
    from tsdf_b200.capi import lib, check, fptr, fvec, colmajor
    from tsdf_b200 import scenes, capi
    rng = np.random.default_rng(5)
    cam = scenes.orbit_camera(0, 12)
    pose = np.asarray(cam.pose, np.float32)
    for n, density in (((512, 512, 512), 2e-4), ((512, 512, 512), 0.0), ((1024, 1024, 264), 1e-4), ((104, 72, 40), 0.02),
                       ((520, 96, 200), 1e-3), ((64, 64, 8), 0.05), ((128, 128, 128), 1.0)):
        nb = tuple((v + 7) // 8 for v in n)
        nbr = nb[0] * nb[1] * nb[2]
        phys = fvec([3000.0 * v / max(n) for v in n])
        voxel, trunc = capi.volume_params(n, phys)
        dist = torch.full((n[0] * n[1] * n[2],), float(trunc), dtype=torch.float32, device="cuda")
        occ = torch.zeros(lib.tsdf_b200_occupancy_bytes(*n), dtype=torch.uint8, device="cuda")
        flags = (rng.random(nbr) < density).astype(np.uint8)
        occ[:nbr] = torch.from_numpy(flags).cuda()
        table = torch.empty(4416, dtype=torch.float32, device="cuda")
        check(lib.tsdf_b200_ray_table(trunc, C.c_void_p(table.data_ptr()), None))
        V = torch.empty(8 * 4 * 3, dtype=torch.float32, device="cuda")
        check(lib.tsdf_b200_raycast_ex(C.c_void_p(dist.data_ptr()), *n, fptr(voxel), fptr(fvec([0, 0, 0])), fptr(phys), trunc,
                                       fptr(fvec(pose[:3, 3])), fptr(colmajor(pose[:3, :3])), fptr(colmajor(cam.kinv)), 8, 4,
                                       C.c_void_p(table.data_ptr()), C.c_void_p(occ.data_ptr()), C.c_void_p(V.data_ptr()), None, None,
                                       0, None), "raycast")
        torch.cuda.synchronize()
        got = occ[nbr:2 * nbr].cpu().numpy().reshape(nb[2], nb[1], nb[0])
        f3 = flags.reshape(nb[2], nb[1], nb[0])
        if f3.any():
            want = np.minimum(ndimage.distance_transform_cdt(f3 == 0, metric="chessboard"), 16).astype(np.uint8)
        else:
            want = np.full(f3.shape, 16, np.uint8)
        assert np.array_equal(got, want), f"grid {nb}, density {density}: {int((got != want).sum())} bricks differ"
        assert np.array_equal(occ[:nbr].cpu().numpy(), flags), "flags modified"
        del dist


def test_depth_pyramid_equals_max_pooling(G):
    """tsdf_b200_depth_stage: level l of the culling pyramid holds the largest depth of every 2^l x 2^l pixel tile (levels 3 up
    to the single-entry top), for image sizes with and without whole tiles and rows that are not 16-byte multiples."""
    import ctypes as C
    import torch
    from tsdf_b200.capi import lib, check
    rng = np.random.default_rng(11)
    for w, h in ((640, 480), (322, 241), (64, 48), (8, 8), (1000, 37)):
        depth = rng.integers(0, 65536, size=(h, w), dtype=np.uint16)
        depth[rng.random((h, w)) < 0.3] = 0
        d = torch.from_numpy(depth).cuda()
        st = torch.zeros((lib.tsdf_b200_depth_staged_bytes(w, h) + 1) // 2, dtype=torch.int16, device="cuda")
        check(lib.tsdf_b200_depth_stage(C.c_void_p(d.data_ptr()), w, h, C.c_void_p(st.data_ptr()), None), "depth_stage")
        torch.cuda.synchronize()
        got = st.cpu().numpy().view(np.uint16)
        off, level = 0, 3
        while True:
            t = 1 << level
            wl, hl = (w + t - 1) // t, (h + t - 1) // t
            pad = np.zeros((hl * t, wl * t), np.uint16)
            pad[:h, :w] = depth
            want = pad.reshape(hl, t, wl, t).max(axis=(1, 3))
            assert np.array_equal(got[off:off + wl * hl].reshape(hl, wl), want), f"{w}x{h}: pyramid level {level} differs"
            off += wl * hl
            if wl == 1 and hl == 1:
                break
            level += 1
