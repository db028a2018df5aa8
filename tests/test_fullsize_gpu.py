"""Full-size parity: BASELINE.json's configurations at their real sizes.

* config 2 geometry (256^3, 640x480, orbit): three frames fused and one raycast, GPU against the CPU oracle bit for bit
  (the oracle needs ~20 s of host time for this).
* config 3 geometry (512^3): size-independent properties the integration offers, checked GPU against GPU where the oracle
  would take minutes: the rigid kernel equals the general kernel; integrating Z-ranges one after another equals integrating
  the whole volume; the culling pyramid changes nothing; the voxels-rewritten counter equals the number of weights that
  moved; and a raycast with empty-space skipping equals one without it.
"""
import numpy as np
import pytest

from helpers import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gpu_util
    return gpu_util


def test_config2_256_orbit_against_oracle(G):
    from oracle import oracle
    from tsdf_b200 import scenes
    n, phys = (256,) * 3, (3000.0,) * 3
    dv = G.DeviceVolume(n, phys)
    ov = oracle.OracleVolume(n, phys)
    cams = [scenes.orbit_camera(i, 200) for i in (0, 37, 111)]
    for cam in cams:
        depth = scenes.render_depth(cam)
        assert dv.integrate(depth, cam.inv_pose, cam.k, cam.kinv) == ov.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
    assert_bits_equal(dv.dist.cpu().numpy(), ov.dist, "dist")
    assert_bits_equal(dv.weight.cpu().numpy(), ov.weight, "weight")
    cam = cams[1]
    V, N, kh, samples = dv.raycast(640, 480, cam.pose, cam.kinv, fastdiv=True)
    Vo, No, kho, marched = ov.raycast(640, 480, cam.pose, cam.kinv)
    assert np.array_equal(kh, kho)
    assert_bits_equal(V, Vo, "vertices")
    assert_bits_equal(N, No, "normals")
    hit = kh >= 0
    assert hit.sum() > 50000 and samples < marched / 20
    vox_g = np.floor((V[hit] - dv.offset) / dv.voxel).astype(np.int64)
    vox_o = np.floor((Vo[hit] - dv.offset) / dv.voxel).astype(np.int64)
    assert np.array_equal(vox_g, vox_o)                     # hit voxel indices


def test_config3_512_properties(G):
    import torch
    from tsdf_b200 import scenes
    n, phys = (512,) * 3, (3000.0,) * 3
    cams = [scenes.orbit_camera(i, 1000) for i in (3, 260, 640)]
    depths = [scenes.render_depth(c) for c in cams]

    a = G.DeviceVolume(n, phys)
    counts = [a.integrate(d, c.inv_pose, c.k, c.kinv) for c, d in zip(cams, depths)]
    w = a.weight.clone()
    assert int((w > 0).sum().item()) <= sum(counts)
    assert int(w.sum(dtype=torch.float64).item()) == sum(counts)      # every rewrite adds exactly 1 to one weight

    # (1) Z-ranges compose, (2) without the culling pyramid
    b = G.DeviceVolume(n, phys)
    for c, d in zip(cams, depths):
        got = 0
        for z0, z1 in ((0, 100), (100, 101), (101, 384), (384, 512)):
            got += b.integrate(d, c.inv_pose, c.k, c.kinv, z_begin=z0, z_end=z1, staged=(z0 != 101))
        assert got == counts[cams.index(c)]
    assert torch.equal(a.dist.view(torch.int32), b.dist.view(torch.int32))
    assert torch.equal(a.weight.view(torch.int32), b.weight.view(torch.int32))
    assert torch.equal(a.occ[: a.occ.numel() // 3], b.occ[: b.occ.numel() // 3])      # brick flags
    del b

    # (3) the general kernel (any matrices, no fast path) gives the same bits
    g = G.DeviceVolume(n, phys)
    G.lib.tsdf_b200_debug_force_generic_integrate(1)
    try:
        for c, d in zip(cams, depths):
            assert g.integrate(d, c.inv_pose, c.k, c.kinv) == counts[cams.index(c)]
    finally:
        G.lib.tsdf_b200_debug_force_generic_integrate(0)
    assert torch.equal(a.dist.view(torch.int32), g.dist.view(torch.int32))
    assert torch.equal(a.weight.view(torch.int32), g.weight.view(torch.int32))
    assert torch.equal(a.occ[: a.occ.numel() // 3], g.occ[: g.occ.numel() // 3])
    del g

    # (3b) the other staging variants of the rigid kernel — TMA boxes + mbarriers, and the two-pass work list — give the same bits
    for variant in (1, 2):
        g = G.DeviceVolume(n, phys)
        G.lib.tsdf_b200_debug_integrate_variant(variant)
        try:
            for c, d in zip(cams, depths):
                assert g.integrate(d, c.inv_pose, c.k, c.kinv) == counts[cams.index(c)]
        finally:
            G.lib.tsdf_b200_debug_integrate_variant(0)
        assert torch.equal(a.dist.view(torch.int32), g.dist.view(torch.int32)), f"variant {variant}: dist"
        assert torch.equal(a.weight.view(torch.int32), g.weight.view(torch.int32)), f"variant {variant}: weight"
        assert torch.equal(a.occ[: a.occ.numel() // 3], g.occ[: g.occ.numel() // 3]), f"variant {variant}: brick flags"
        del g

    # (4) raycast: skipping changes nothing but the number of samples evaluated
    cam = cams[1]
    V1, N1, k1, s1 = a.raycast(640, 480, cam.pose, cam.kinv, skip=True, fastdiv=True)
    V0, N0, k0, s0 = a.raycast(640, 480, cam.pose, cam.kinv, skip=False, fastdiv=False)
    assert np.array_equal(k1, k0)
    assert_bits_equal(V1, V0, "vertices")
    assert_bits_equal(N1, N0, "normals")
    assert (k1 >= 0).sum() > 50000 and s1 < s0 / 20
