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


def test_exchange_entry_points_validate_arguments(built):
    """The multi-GPU exchange calls reject bad arguments before any CUDA call (no device here)."""
    import ctypes as C
    from tsdf_b200 import capi
    lib = capi.lib
    one = (C.c_void_p * 1)(C.c_void_p(16))
    none = (C.c_void_p * 1)(None)
    # bricks_push: null source / zero destinations / too many / slab not a multiple of the brick / rank >= world / null dst
    assert lib.tsdf_b200_bricks_push(None, 8, 8, 8, 8, 1, 0, C.c_void_p(16), 1, one, None, None) == -1
    assert lib.tsdf_b200_bricks_push(C.c_void_p(16), 8, 8, 8, 8, 1, 0, C.c_void_p(16), 0, one, None, None) == -1
    assert lib.tsdf_b200_bricks_push(C.c_void_p(16), 8, 8, 8, 8, 1, 0, C.c_void_p(16), 17, one, None, None) == -1
    assert lib.tsdf_b200_bricks_push(C.c_void_p(16), 8, 8, 8, 12, 1, 0, C.c_void_p(16), 1, one, None, None) == -1
    assert lib.tsdf_b200_bricks_push(C.c_void_p(16), 8, 8, 8, 8, 2, 2, C.c_void_p(16), 1, one, None, None) == -1
    assert lib.tsdf_b200_bricks_push(C.c_void_p(16), 8, 8, 8, 8, 1, 0, C.c_void_p(16), 1, none, None, None) == -1
    f3 = capi.fptr(np.ones(3, np.float32))
    f9 = capi.fptr(np.eye(3, dtype=np.float32).reshape(-1))
    # raycast_tiles: rank >= world, no outputs; raycast_mirrored: image not a whole number of 8x4 tiles, misaligned mirror
    assert lib.tsdf_b200_raycast_tiles(C.c_void_p(16), 8, 8, 8, f3, f3, f3, 1.0, f3, f9, f9, 8, 4, C.c_void_p(16), None,
                                       2, 2, 1, one, None, 0, None) == -1
    assert lib.tsdf_b200_raycast_tiles(C.c_void_p(16), 8, 8, 8, f3, f3, f3, 1.0, f3, f9, f9, 8, 4, C.c_void_p(16), None,
                                       1, 0, 0, one, None, 0, None) == -1
    assert lib.tsdf_b200_raycast_mirrored(C.c_void_p(16), 8, 8, 8, f3, f3, f3, 1.0, f3, f9, f9, 10, 4, C.c_void_p(16), None,
                                          C.c_void_p(16), C.c_void_p(32), None, 0, None) == -1
    assert lib.tsdf_b200_raycast_mirrored(C.c_void_p(16), 8, 8, 8, f3, f3, f3, 1.0, f3, f9, f9, 8, 4, C.c_void_p(16), None,
                                          C.c_void_p(16), C.c_void_p(36), None, 0, None) == -1
    p = C.c_void_p()
    assert lib.tsdf_b200_peer_open(None, C.byref(p)) == -1
    assert lib.tsdf_b200_peer_alloc(0, C.byref(p), None) == -1
    assert lib.tsdf_b200_peer_close(None) == 0 and lib.tsdf_b200_peer_free(None) == 0


def test_host_pool_without_a_device(built):
    """tsdf_b200_host_alloc / _free: small blocks and — on a box without a CUDA device — large ones come from malloc; blocks are
    writable, free accepts NULL; tsdf_b200_raycast_fused validates its arguments like tsdf_b200_raycast_mirrored."""
    import ctypes as C
    from tsdf_b200 import capi
    lib = capi.lib
    for size in (1, 1000, 300 * 1024, 4 << 20):
        p = lib.tsdf_b200_host_alloc(size)
        assert p
        C.memset(p, 0x5a, size)
        assert C.string_at(p + size - 1, 1) == b"\x5a"
        lib.tsdf_b200_host_free(C.c_void_p(p))
    lib.tsdf_b200_host_free(None)
    assert lib.tsdf_b200_raycast_tile_counters(640, 480) == 2 * 80 * 120
    assert lib.tsdf_b200_raycast_tile_counters(9, 5) == 2 * 2 * 2
    f3 = capi.fptr(np.ones(3, np.float32)); f9 = capi.fptr(np.eye(3, dtype=np.float32).reshape(-1))
    d = C.c_void_p(16)
    # no normal map / ragged image with mirrors / misaligned mirror
    assert lib.tsdf_b200_raycast_fused(d, 8, 8, 8, f3, f3, f3, 1.0, f3, f9, f9, 8, 4, d, None, d, None, None, None, None, None, 0, None) == -1
    assert lib.tsdf_b200_raycast_fused(d, 8, 8, 8, f3, f3, f3, 1.0, f3, f9, f9, 10, 4, d, None, d, d, d, d, None, None, 0, None) == -1
    assert lib.tsdf_b200_raycast_fused(d, 8, 8, 8, f3, f3, f3, 1.0, f3, f9, f9, 8, 4, d, None, d, d, C.c_void_p(24), d, None, None, 0, None) == -1
