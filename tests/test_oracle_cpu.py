"""CPU-only tests: the oracle against the reference's golden vectors and its stated semantics."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle
from tsdf_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "t_100_2000_50.json")
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def golden():
    return json.load(open(GOLDEN))


def test_truncation_distance_matches_reference_fixture(golden):
    # TestData/t_100_2000_50.tsdf: 100^3 over 2000 mm -> trunc = 1.1f*|voxel| = 38.10512 (bit-exact)
    vox, trunc = oracle.volume_params(golden["size"], golden["physical"])
    assert vox.tolist() == [20.0, 20.0, 20.0]
    assert int(np.float32(trunc).view(np.uint32)) == golden["trunc_bits"]


def test_clear_matches_reference_fixture(golden):
    # The fixture was saved right after construction: set_size -> clear() ran with offset (0,0,0)
    # (the offset (50,50,50) in the header was set afterwards), SURVEY.md §4.
    v = oracle.OracleVolume(golden["size"], golden["physical"], with_deformation=True)
    assert sha(v.dist) == golden["dist_sha256"]
    assert sha(v.weight) == golden["weight_sha256"]
    assert sha(v.deform) == golden["deform_sha256"]
    for i, node in golden["deform_samples"].items():
        assert v.deform[6 * int(i):6 * int(i) + 6].tolist() == node


def test_voxel_centre_semantics():
    # Test_TSDFMetrics.cpp:97-108 (stale API, but pins the formula): 3x4x5 voxels over 3000 mm,
    # voxel (0,0,0) has centre (500, 375, 300).
    v = oracle.OracleVolume((3, 4, 5), (3000, 3000, 3000), with_deformation=True)
    assert v.deform[0:3].tolist() == [500.0, 375.0, 300.0]
    assert v.deform[6 * (3 * 4 * 5 - 1):6 * (3 * 4 * 5 - 1) + 3].tolist() == [2500.0, 2625.0, 2700.0]


def test_integrate_analytic_grid_equals_stored_grid():
    # Reading the deformation array clear() wrote and recomputing it analytically are the same bits.
    cam = scenes.orbit_camera(3, 20)
    depth = scenes.render_depth(cam, 160, 120)
    k = cam.k.copy(); k[:2] *= 0.25
    kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
    a = oracle.OracleVolume((48, 40, 32), (3000, 3000, 3000), with_deformation=True)
    b = oracle.OracleVolume((48, 40, 32), (3000, 3000, 3000), with_deformation=False)
    for v in (a, b):
        v.offset[:] = (10.5, -3.25, 7.0)      # offset applied after clear(): added on top of the grid
    na = a.integrate(depth, cam.inv_pose, k, kinv)
    nb = b.integrate(depth, cam.inv_pose, k, kinv)
    assert na == nb and na > 0
    assert sha(a.dist) == sha(b.dist) and sha(a.weight) == sha(b.weight)


def test_integrate_semantics_wall():
    # A flat wall 1000 mm in front of an identity camera: voxels in front get +trunc-clamped
    # positive sdf, voxels further than trunc behind it are untouched (TSDFVolume.cu:365).
    cam = scenes.PinholeCamera()
    cam.move_to(1500, 1500, -1000)
    depth = np.full((480, 640), 2000, np.uint16)       # wall at world z = 1000
    v = oracle.OracleVolume((32, 32, 32), (3000, 3000, 3000))
    n = v.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
    d = v.dist.reshape(32, 32, 32)   # [z][y][x]
    w = v.weight.reshape(32, 32, 32)
    zc = (np.arange(32) + 0.5) * v.voxel[2]
    col_d, col_w = d[:, 16, 16], w[:, 16, 16]
    assert n == int(w.sum())
    for z in range(32):
        sdf = 1000.0 - zc[z]
        if sdf >= -v.trunc:
            assert col_w[z] == 1.0
            assert col_d[z] == np.float32(min(sdf, v.trunc))
        else:
            assert col_w[z] == 0.0 and col_d[z] == v.trunc
    # weights are not clamped at max_weight = 15 (clamp commented out, TSDFVolume.cu:378)
    for _ in range(19):
        v.integrate(depth, cam.inv_pose, cam.k, cam.kinv)
    assert w.max() == 20.0


def test_raycast_wall_hit_and_normal():
    # Intent of the commented-out reference test (Test_TSDF_RayCast.cpp:307-342): a wall is hit
    # at its plane and the vertex-map normal points back at the camera.
    n = 64
    v = oracle.OracleVolume((n, n, n), (3000, 3000, 3000))
    zc = (np.arange(n, dtype=np.float32) + 0.5) * v.voxel[2]
    sdf = np.clip(1200.0 - zc, -v.trunc, v.trunc).astype(np.float32)
    v.dist[:] = np.repeat(sdf, n * n)
    cam = scenes.PinholeCamera()
    cam.move_to(1500, 1500, -1500)
    V, N, kh, ns = v.raycast(640, 480, cam.pose, cam.kinv)
    c = 240 * 640 + 320
    assert kh[c] >= 0
    assert abs(V[c, 2] - 1200.0) < 0.05 * v.voxel[2]
    assert abs(N[c, 2] + 1.0) < 1e-3 and abs(N[c, 0]) < 1e-3 and abs(N[c, 1]) < 1e-3
    # last row / column of the normal map are zero (GPURaycaster.cu:405-413)
    Nm = N.reshape(480, 640, 3)
    assert not Nm[-1].any() and not Nm[:, -1].any()
    # the direction is not normalised, so t is camera-z: samples to reach the wall = 2700/step-ish
    step = np.float32(np.float64(v.trunc) * 0.05)
    assert ns > 0 and kh.max() <= 4401


def test_raycast_sample_cap():
    # An all-positive volume never hits; a ray crossing > 4402 steps stops at the cap
    # (GPURaycaster.cu:369) and every pixel is NaN.
    v = oracle.OracleVolume((512, 4, 4), (3000, 3000, 3000))   # voxel.x small -> trunc small -> tiny step
    cam = scenes.PinholeCamera()
    cam.move_to(-10.0, 1500.0, 1500.0)
    cam.look_at(3000.0, 1500.0, 1500.0)
    V, N, kh, ns = v.raycast(64, 48, cam.pose, cam.kinv)
    assert np.isnan(V).all() and (kh == -1).all()
    assert ns <= 64 * 48 * 4402


def test_threads_do_not_change_results():
    cam = scenes.fixed_pose_camera()
    depth = scenes.render_depth(cam, 320, 240)
    k = cam.k.copy(); k[:2] *= 0.5
    kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
    out = []
    for t in (1, 4):
        oracle.set_threads(t)
        v = oracle.OracleVolume((40, 40, 40), (3000, 3000, 3000))
        v.integrate(depth, cam.inv_pose, k, kinv)
        V, N, kh, ns = v.raycast(320, 240, cam.pose, kinv)
        out.append((sha(v.dist), sha(V), sha(kh), ns))
    oracle.set_threads(os.cpu_count())
    assert out[0] == out[1]
