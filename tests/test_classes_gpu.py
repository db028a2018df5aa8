"""The drop-in C++ class layer on the GPU: TSDFVolume / GPURaycaster / extract_surface / save+load through the classes
(tests/cpp/class_tests.cpp --gpu), and — when the prebuilt binary travelled with the snapshot — the reference's own
kinfu driver, compiled unchanged against tsdf_b200/include, run on a synthetic TUM-format directory."""
import os
import subprocess
import sys

import pytest

from test_classes_cpu import BUILD, ROOT, build_class_tests

pytestmark = pytest.mark.gpu


def test_classes_on_gpu(built, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = build_class_tests()
    out = subprocess.run([exe, "--gpu", str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "0 failures (with GPU part)" in out.stdout


def test_reference_kinfu_runs_on_synthetic_tum_directory(built, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    kinfu = os.path.join(BUILD, "kinfu")
    if not os.path.exists(kinfu):
        pytest.skip("build/kinfu not prebuilt (needs the reference tree at build time)")
    tum = tmp_path / "tum"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_tum_dir.py"), str(tum), "--frames", "3"],
                          stdout=subprocess.DEVNULL)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "tsdf_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([kinfu, "-m", "3", "-d", str(tum)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    # the driver's own progress lines (kinfu.cpp:34,176,209,212): three frames fused, a raycast, a non-empty mesh
    assert out.stdout.count("Integrating frame") == 3
    assert "Raycasting" in out.stdout and "Extracting ISO surface" in out.stdout
    line = [l for l in out.stdout.splitlines() if l.startswith("Writing ") and "vertices" in l][-1]
    assert int(line.split()[1]) > 1000
