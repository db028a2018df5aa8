"""ctypes wrapper of oracle/liboracle.so (the CPU restatement in oracle/tsdf_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the tsdf_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "liboracle.so")


def build(force=False):
    src = os.path.join(_DIR, "tsdf_oracle.c")
    tables = os.path.join(_DIR, "..", "tsdf_b200", "csrc", "mc_tables.h")
    newest = max(os.path.getmtime(src), os.path.getmtime(tables))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
                               "-Wall", "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH


build()
_lib = C.CDLL(LIB_PATH)
_f = C.POINTER(C.c_float)
_u32 = C.c_uint32
_vp = C.c_void_p

_lib.oracle_set_threads.restype = C.c_int
_lib.oracle_set_threads.argtypes = [C.c_int]
_lib.oracle_volume_params.restype = None
_lib.oracle_volume_params.argtypes = [_u32, _u32, _u32, _f, _f, _f]
_lib.oracle_clear.restype = None
_lib.oracle_clear.argtypes = [_vp, _vp, _vp, _u32, _u32, _u32, _f, _f, C.c_float]
_lib.oracle_integrate.restype = C.c_uint64
_lib.oracle_integrate.argtypes = [_vp, _vp, _vp, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f, _u32, _u32, _vp, _u32, _u32]
_lib.oracle_raycast.restype = C.c_uint64
_lib.oracle_raycast.argtypes = [_vp, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f, _u32, _u32, _vp, _vp, _u32, _u32]
_lib.oracle_raycast_slab.restype = None
_lib.oracle_raycast_slab.argtypes = [_vp, _u32, _u32, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f, _u32, _u32, _vp]
_lib.oracle_resolve.restype = None
_lib.oracle_resolve.argtypes = [_vp, _f, _f, C.c_float, _f, _f, _f, _u32, _u32, _vp, _vp]
_lib.oracle_normals.restype = None
_lib.oracle_normals.argtypes = [_u32, _u32, _vp, _vp]
_lib.oracle_mc_extract.restype = C.c_uint64
_lib.oracle_mc_extract.argtypes = [_vp, _u32, _u32, _u32, _f, _f, _vp]
_lib.oracle_bilateral.restype = None
_lib.oracle_bilateral.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, _f, C.c_int, _f, C.c_int]
_lib.oracle_hit_voxels.restype = None
_lib.oracle_hit_voxels.argtypes = [_u32, _vp, _f, _f, _u32, _u32, _vp]


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f)


def _fv(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))


def _cm(m):
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T.reshape(-1))


def set_threads(n):
    return _lib.oracle_set_threads(int(n))


def volume_params(n, physical):
    vox = np.zeros(3, np.float32)
    tr = C.c_float()
    _lib.oracle_volume_params(n[0], n[1], n[2], _fp(_fv(physical)), _fp(vox), C.byref(tr))
    return vox, np.float32(tr.value)


class OracleVolume:
    """Host-memory twin of the reference TSDFVolume state touched by the hot path."""

    def __init__(self, n, physical, with_deformation=False):
        self.size = tuple(int(x) for x in n)
        self.physical = _fv(physical)
        self.voxel, self.trunc = volume_params(self.size, self.physical)
        self.offset = np.zeros(3, np.float32)
        self.offset_at_clear = np.zeros(3, np.float32)
        nv = self.size[0] * self.size[1] * self.size[2]
        self.dist = np.empty(nv, np.float32)
        self.weight = np.empty(nv, np.float32)
        self.deform = np.empty(nv * 6, np.float32) if with_deformation else None
        self.clear()

    def clear(self):
        self.offset_at_clear = self.offset.copy()
        _lib.oracle_clear(self.dist.ctypes.data, self.weight.ctypes.data,
                          self.deform.ctypes.data if self.deform is not None else None,
                          *self.size, _fp(self.voxel), _fp(self.offset_at_clear), self.trunc)

    def integrate(self, depth, inv_pose, k, kinv, z_begin=0, z_end=None):
        h, w = depth.shape
        assert depth.dtype == np.uint16 and depth.flags["C_CONTIGUOUS"]
        z_end = self.size[2] if z_end is None else z_end
        return int(_lib.oracle_integrate(self.dist.ctypes.data, self.weight.ctypes.data,
                                         self.deform.ctypes.data if self.deform is not None else None,
                                         *self.size, _fp(self.voxel), _fp(self.offset_at_clear), _fp(self.offset),
                                         self.trunc, _fp(_cm(inv_pose)), _fp(_cm(k)), _fp(_cm(kinv)), w, h,
                                         depth.ctypes.data, z_begin, z_end))

    def raycast(self, w, h, pose, kinv, want_khit=True, y_begin=0, y_step=1, want_normals=True):
        pose = np.asarray(pose, np.float32)
        vertices = np.empty((h * w, 3), np.float32)
        khit = np.empty(h * w, np.int32) if want_khit else None
        smin = self.offset.copy()
        smax = (self.offset + self.physical).astype(np.float32)
        origin = _fv(pose[:3, 3])
        n = _lib.oracle_raycast(self.dist.ctypes.data, *self.size, _fp(self.voxel), _fp(smin), _fp(smax), self.trunc,
                                _fp(origin), _fp(_cm(pose[:3, :3])), _fp(_cm(kinv)), w, h, vertices.ctypes.data,
                                khit.ctypes.data if want_khit else None, y_begin, y_step)
        normals = None
        if want_normals:
            normals = np.empty((h * w, 3), np.float32)
            _lib.oracle_normals(w, h, vertices.ctypes.data, normals.ctypes.data)
        return vertices, normals, khit, int(n)


def shard_ranges(nz, world, brick=8):
    """Z-slab ownership used by the multi-GPU path: whole bricks, rank order (mirrors tsdf_b200.sharded)."""
    bricks = (nz + brick - 1) // brick
    per = (bricks + world - 1) // world
    return [(min(r * per * brick, nz), min((r + 1) * per * brick, nz)) for r in range(world)]


def raycast_slab_keys(vol, w, h, pose, kinv, z_lo, z_hi):
    pose = np.asarray(pose, np.float32)
    keys = np.empty(h * w, np.int64)
    smin = vol.offset.copy()
    smax = (vol.offset + vol.physical).astype(np.float32)
    _lib.oracle_raycast_slab(vol.dist.ctypes.data, *vol.size, z_lo, z_hi, _fp(vol.voxel), _fp(smin), _fp(smax), vol.trunc,
                             _fp(_fv(pose[:3, 3])), _fp(_cm(pose[:3, :3])), _fp(_cm(kinv)), w, h, keys.ctypes.data)
    return keys


def resolve_keys(vol, keys, w, h, pose, kinv):
    pose = np.asarray(pose, np.float32)
    keys = np.ascontiguousarray(keys, np.int64)
    V = np.empty((h * w, 3), np.float32)
    kh = np.empty(h * w, np.int32)
    smin = vol.offset.copy()
    smax = (vol.offset + vol.physical).astype(np.float32)
    _lib.oracle_resolve(keys.ctypes.data, _fp(smin), _fp(smax), vol.trunc, _fp(_fv(pose[:3, 3])), _fp(_cm(pose[:3, :3])),
                        _fp(_cm(kinv)), w, h, V.ctypes.data, kh.ctypes.data)
    return V, kh


def normals(w, h, vertices):
    vertices = np.ascontiguousarray(vertices, np.float32)
    out = np.empty((h * w, 3), np.float32)
    _lib.oracle_normals(w, h, vertices.ctypes.data, out.ctypes.data)
    return out


def hit_voxels(vertices, space_min, voxel, nx, ny):
    vertices = np.ascontiguousarray(vertices, np.float32)
    out = np.empty(vertices.shape[0], np.int64)
    _lib.oracle_hit_voxels(vertices.shape[0], vertices.ctypes.data, _fp(_fv(space_min)), _fp(_fv(voxel)), nx, ny,
                           out.ctypes.data)
    return out


def mc_extract(dist, n, voxel, offset):
    """Marching cubes of a host distance array (x fastest): (n_vertices, 3) float32 in the reference's order."""
    dist = np.ascontiguousarray(dist, np.float32)
    vox, off = _fv(voxel), _fv(offset)
    count = _lib.oracle_mc_extract(dist.ctypes.data, n[0], n[1], n[2], _fp(vox), _fp(off), None)
    out = np.empty((count, 3), np.float32)
    if count:
        got = _lib.oracle_mc_extract(dist.ctypes.data, n[0], n[1], n[2], _fp(vox), _fp(off), out.ctypes.data)
        assert got == count
    return out


def bilateral_tables(sigma_colour, sigma_space, n_similarity=256):
    """The look-up tables of BilateralFilter's constructor (reference src/BilateralFilter.cpp:15-42), float32 like the
    reference (std::exp on a float argument is expf)."""
    f = np.float32
    radius = int(np.ceil(f(sigma_space) * f(1.5)))
    size = 2 * radius + 1
    inv_c = f(1.0) / (f(sigma_colour) * f(sigma_colour))
    inv_s = f(1.0) / (f(sigma_space) * f(sigma_space))
    kernel = np.empty(size * size, np.float32)
    i = 0
    for x in range(-radius, radius + 1):
        for y in range(-radius, radius + 1):
            kernel[i] = _expf(-(f(x * x + y * y)) * inv_s)
            i += 1
    similarity = np.array([_expf(-(f(d)) * inv_c) for d in range(n_similarity)], np.float32)
    return kernel, similarity


_libm = C.CDLL("libm.so.6")
_libm.expf.restype = C.c_float
_libm.expf.argtypes = [C.c_float]


def _expf(x):
    return np.float32(_libm.expf(C.c_float(float(x))))


def bilateral(image, kernel, similarity):
    """Filtered copy of a uint8 / uint16 (H, W) image with the given tables."""
    image = np.ascontiguousarray(image)
    bits = {np.dtype(np.uint8): 8, np.dtype(np.uint16): 16}[image.dtype]
    out = np.empty_like(image)
    size = int(round(np.sqrt(kernel.size)))
    _lib.oracle_bilateral(image.ctypes.data, out.ctypes.data, bits, image.shape[1], image.shape[0], _fp(kernel), size,
                          _fp(similarity), similarity.size)
    return out
