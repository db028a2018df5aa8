"""Marching cubes on the GPU against the oracle: same vertices, same order, bit for bit (tsdf_b200_mc_extract vs
oracle_mc_extract, both restating src/MarchingCubes/MarkAndSweepMC.cu), plus Z-shard concatenation."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_bits_equal

pytestmark = pytest.mark.gpu


def gpu_mc(lib, check, fptr, d_dist, n, z_base, cz0, cz1, vox, off):
    import torch
    out = C.c_void_p()
    count = C.c_ulonglong()
    check(lib.tsdf_b200_mc_extract(C.c_void_p(d_dist.data_ptr()), n[0], n[1], n[2], z_base, cz0, cz1, fptr(vox), fptr(off),
                                   C.byref(out), C.byref(count), None), "mc_extract")
    v = np.empty((count.value, 3), np.float32)
    if count.value:
        check(lib.tsdf_b200_copy_to_host(v.ctypes.data, out, v.nbytes), "copy_to_host")
        lib.tsdf_b200_device_free(out)
    torch.cuda.synchronize()
    return v


@pytest.mark.parametrize("n", [(40, 36, 44), (7, 5, 3), (2, 2, 2), (65, 33, 20)])
def test_mc_matches_oracle(built, n):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import oracle
    from tsdf_b200.capi import lib, check, fptr
    rng = np.random.default_rng(n[0] * 1000 + n[1])
    vox = np.array([12.5, 9.75, 11.0], np.float32)
    off = np.array([3.0, -40.5, 1000.0], np.float32)
    z, y, x = np.meshgrid(*(np.arange(m) + 0.5 for m in (n[2], n[1], n[0])), indexing="ij")
    c = np.array([n[0] * vox[0], n[1] * vox[1], n[2] * vox[2]]) * 0.5
    d = np.sqrt((x * vox[0] - c[0]) ** 2 + (y * vox[1] - c[1]) ** 2 + (z * vox[2] - c[2]) ** 2) - 0.3 * c.min() * 2
    d = (d + rng.normal(scale=2.0, size=d.shape)).astype(np.float32).reshape(-1)     # noisy: many cube types
    d[rng.integers(0, d.size, size=5)] = 0.0                                         # exact zeros are "outside"
    want = oracle.mc_extract(d, n, vox, off)
    d_dist = torch.from_numpy(d).cuda()
    got = gpu_mc(lib, check, fptr, d_dist, n, 0, 0, n[2] - 1, vox, off)
    assert got.shape == want.shape
    assert_bits_equal(got, want, "mesh vertices")
    if n[2] >= 8:
        # Z-shards: rank r extracts the cubes based in its planes from a slab with one halo plane; concatenated in rank
        # order they reproduce the whole mesh
        cut = n[2] // 2
        plane = n[0] * n[1]
        lo = gpu_mc(lib, check, fptr, d_dist[: plane * (cut + 1)].contiguous(), (n[0], n[1], cut + 1), 0, 0, cut, vox, off)
        hi = gpu_mc(lib, check, fptr, d_dist[plane * cut:].contiguous(), (n[0], n[1], n[2] - cut), cut, 0, n[2] - cut - 1, vox, off)
        assert_bits_equal(np.concatenate([lo, hi]), want, "sharded mesh vertices")


def test_mc_empty_volume(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tsdf_b200.capi import lib, check, fptr
    n = (16, 16, 16)
    d_dist = torch.full((16 ** 3,), 5.0, dtype=torch.float32, device="cuda")
    one = np.ones(3, np.float32)
    assert gpu_mc(lib, check, fptr, d_dist, n, 0, 0, 15, one, one).shape == (0, 3)


def test_engine_mesh_per_shard_concatenates(built):
    """ShardedEngine.extract_mesh: the meshes of 3 emulated Z-shards (each from its own slab + halo plane), concatenated in
    rank order, are the single-volume mesh bit for bit."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tsdf_b200 import scenes, sharded
    n, phys = (96, 80, 112), (3000.0, 2500.0, 3000.0)
    whole = sharded.ShardedEngine(n, phys)
    ranks = [sharded.ShardedEngine(n, phys, rank=r, world=3) for r in range(3)]
    for frame in (0, 4, 8):
        cam = scenes.orbit_camera(frame, 12)
        k = cam.k.copy(); k[:2] *= 0.5
        cam.k = k
        cam.kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
        depth = torch.from_numpy(scenes.render_depth(cam, 320, 240)).cuda()
        for e in [whole] + ranks:
            e.integrate(depth, cam)
    want = whole.extract_mesh().cpu().numpy()
    got = np.concatenate([e.extract_mesh().cpu().numpy() for e in ranks])
    assert want.shape[0] > 3000 and want.shape[0] % 3 == 0
    assert_bits_equal(got, want, "sharded mesh")
