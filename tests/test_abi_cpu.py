"""CPU-only: the C-ABI library loads and exports every symbol include/tsdf_b200.h declares."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tsdf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsdf_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    from tsdf_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/tsdf_b200.h but not exported"
    # and the Python binding covers all of them
    assert set(names) == set(capi.SIGNATURES)


def test_host_only_entry_points(built):
    from tsdf_b200 import capi
    assert b"sm_100a" in capi.lib.tsdf_b200_version()
    assert capi.lib.tsdf_b200_strerror(-1) == b"invalid argument"
    vox, trunc = capi.volume_params((100, 100, 100), (2000, 2000, 2000))
    assert vox.tolist() == [20.0, 20.0, 20.0]
    assert int(np.float32(trunc).view(np.uint32)) == 1108896677        # 38.10512, reference fixture
    assert capi.lib.tsdf_b200_occupancy_bytes(512, 512, 512) == 3 * 64 ** 3
    assert capi.lib.tsdf_b200_occupancy_bytes(100, 9, 1) == 3 * 13 * 2 * 1
    # argument validation happens before any CUDA call
    assert capi.lib.tsdf_b200_integrate(None, None, None, 1, 1, 1, None, None, None, 1.0, None, None, None,
                                        1, 1, None, None, 0, 1, 0, None, None, None) == -1
    assert capi.lib.tsdf_b200_normals(0, 0, None, None, None) == -1


def test_params_match_oracle(built):
    from tsdf_b200 import capi
    from oracle import oracle
    rng = np.random.default_rng(7)
    for _ in range(200):
        n = tuple(int(x) for x in rng.integers(1, 1025, size=3))
        phys = rng.uniform(10, 5000, size=3).astype(np.float32)
        v0, t0 = capi.volume_params(n, phys)
        v1, t1 = oracle.volume_params(n, phys)
        assert v0.tobytes() == v1.tobytes() and np.float32(t0).tobytes() == np.float32(t1).tobytes()


def test_product_does_not_import_oracle():
    # the product package must never reach into oracle/
    pkg = os.path.join(ROOT, "tsdf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("CPU oracle", "").lower() or f == "scenes.py", (dirpath, f)
