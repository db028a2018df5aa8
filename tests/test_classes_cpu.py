"""The drop-in C++ class layer, CPU part: builds libtsdf_b200_classes.so, compiles tests/cpp/class_tests.cpp against
tsdf_b200/include and runs its Camera / PNG / TUM loader / PLY known-answer tests (no GPU work).  When the reference
tree is present it also compiles the reference's own src/Tools/kinfu.cpp UNCHANGED against these headers (the drop-in
claim of the north star) — compile and link only; running it needs a GPU (tests/test_classes_gpu.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")


def build_class_tests():
    subprocess.check_call(["make", "-C", ROOT, "-j8", "classes"], stdout=subprocess.DEVNULL)
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "class_tests")
    subprocess.check_call(["g++", "-O1", "-std=c++14", "-Wall", "-DTSDF_B200_PINNED_EIGEN", "-I" + os.path.join(ROOT, "tsdf_b200", "compat"),
                           "-I/usr/local/cuda/include", "-o", exe, os.path.join(ROOT, "tests", "cpp", "class_tests.cpp"),
                           "-L" + os.path.join(ROOT, "tsdf_b200"), "-ltsdf_b200_classes", "-ltsdf_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "tsdf_b200")])
    return exe


def test_class_layer_known_answers(built, tmp_path):
    exe = build_class_tests()
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout


def test_every_reference_header_kinfu_needs_exists():
    """kinfu.cpp includes these by relative path (src/Tools/kinfu.cpp:4-10) plus what they pull in."""
    inc = os.path.join(ROOT, "tsdf_b200", "include")
    for name in ["TSDFVolume.hpp", "PngWrapper.hpp", "DepthMapUtilities.hpp", "RenderUtilities.hpp", "ply.hpp",
                 "TUMDataLoader.hpp", "MarkAndSweepMC.hpp", "Camera.hpp", "DepthImage.hpp", "Raycaster.hpp",
                 "GPURaycaster.hpp", "Definitions.hpp"]:
        assert os.path.exists(os.path.join(inc, name)), name


@pytest.mark.skipif(not os.path.exists("/root/reference/src/Tools/kinfu.cpp"), reason="reference tree not present")
def test_reference_kinfu_compiles_unchanged(built):
    subprocess.check_call(["make", "-C", ROOT, "kinfu"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(BUILD, "kinfu"))
    # the translation unit really is the reference's file, not a copy
    assert os.path.realpath(os.path.join(BUILD, "dropin", "Tools", "kinfu.cpp")) == "/root/reference/src/Tools/kinfu.cpp"
