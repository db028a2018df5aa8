"""The level-2 volume sharded over several GPUs INSIDE ONE PROCESS (tsdf_b200_volume_create_sharded / TSDF_NGPUS, csrc/multi.cu —
the C++ coordinator under the TSDFVolume class): integrate, raycast, read-back, marching cubes, set_distance_data, clear and
save must give the bits of the same calls on a single-GPU volume.  Needs at least two GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import assert_bits_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("n,want_gpus", [((128, 128, 128), 8), ((96, 80, 112), 3), ((64, 64, 24), 2)])
def test_sharded_volume_equals_single_gpu(built, tmp_path, n, want_gpus):
    have = _gpus()
    if have < 2:
        pytest.skip("needs 2 GPUs")
    from tsdf_b200 import Volume, scenes
    from test_parity_gpu import sphere_sdf
    phys = (3000.0, 2500.0, 3000.0)
    one = Volume(n, phys)
    many = Volume(n, phys, gpus=want_gpus)
    assert one.gpus == 1 and 2 <= many.gpus <= min(want_gpus, have)
    w, h = 320, 240
    for step, f in enumerate((0, 2, 5, 7)):
        cam = scenes.orbit_camera(f, 12)
        k = cam.k.copy(); k[:2] *= 0.5
        kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
        depth = scenes.render_depth(scenes_cam(cam, k, kinv), w, h)
        one.integrate(depth, cam.inv_pose, k, kinv)
        many.integrate(depth, cam.inv_pose, k, kinv)
        V1, N1 = one.raycast(w, h, cam.pose, kinv)
        V2, N2 = many.raycast(w, h, cam.pose, kinv)
        assert_bits_equal(V2, V1, f"vertices after frame {f}")
        assert_bits_equal(N2, N1, f"normals after frame {f}")
        assert one.stats()[0] == many.stats()[0] > 0            # voxels rewritten add up over the slabs (halo planes are not counted)
    assert int((~np.isnan(V1[:, 0])).sum()) > 3000
    d1, w1 = one.read()
    d2, w2 = many.read()
    assert_bits_equal(d2, d1, "dist"); assert_bits_equal(w2, w1, "weight")
    m1, m2 = one.extract_mesh(), many.extract_mesh()
    assert m1.shape[0] > 1000 and m1.shape == m2.shape
    assert_bits_equal(m2, m1, "mesh vertices (slab order = cube order)")
    # save from the sharded volume == save from the single one (header, dist, weight, colours, deformation grid)
    p1, p2 = str(tmp_path / "one.tsdf"), str(tmp_path / "many.tsdf")
    one.save(p1); many.save(p2)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    # set_distance_data scatters slabs + halo planes and rebuilds every slab's occupancy grid
    sdf = sphere_sdf(n, phys, one.trunc, (1500, 1200, 1400), 700)
    one.set_distance_data(sdf); many.set_distance_data(sdf)
    cam = scenes.orbit_camera(3, 12)
    V1, N1 = one.raycast(w, h, cam.pose, kinv)
    V2, N2 = many.raycast(w, h, cam.pose, kinv)
    assert_bits_equal(V2, V1, "vertices of the uploaded sphere"); assert_bits_equal(N2, N1, "normals of the uploaded sphere")
    one.clear(); many.clear()
    V2, _ = many.raycast(w, h, cam.pose, kinv)
    assert np.isnan(V2).all()
    one.close(); many.close()


def scenes_cam(cam, k, kinv):
    """The orbit camera with scaled intrinsics (render_depth reads cam.kinv)."""
    cam.k, cam.kinv = k, kinv
    return cam


def test_kinfu_unchanged_on_two_gpus(built, tmp_path):
    """The reference's own kinfu driver, compiled unchanged, with TSDF_NGPUS=2: same progress lines and the same mesh size as
    on one GPU (the volume object behind its TSDFVolume is sharded; nothing in kinfu.cpp knows)."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    kinfu = os.path.join(ROOT, "build", "kinfu")
    if not os.path.exists(kinfu):
        pytest.skip("build/kinfu not prebuilt (needs the reference tree at build time)")
    tum = tmp_path / "tum"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_tum_dir.py"), str(tum), "--frames", "3"],
                          stdout=subprocess.DEVNULL)
    sizes = []
    for gpus in ("1", "2"):
        env = dict(os.environ, TSDF_NGPUS=gpus,
                   LD_LIBRARY_PATH=os.path.join(ROOT, "tsdf_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        out = subprocess.run([kinfu, "-m", "3", "-d", str(tum)], capture_output=True, text=True, env=env, timeout=300)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
        assert out.stdout.count("Integrating frame") == 3
        line = [l for l in out.stdout.splitlines() if l.startswith("Writing ") and "vertices" in l][-1]
        sizes.append(int(line.split()[1]))
    assert sizes[0] > 1000 and sizes[0] == sizes[1]
