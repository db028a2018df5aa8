"""ctypes wrapper of oracle/_ref/libref_cuda_{O3,G}.so — the REFERENCE's own CUDA kernels and host classes,
compiled from /root/reference/src by oracle/build_ref.sh.  TEST INFRASTRUCTURE ONLY (GPU box)."""
import ctypes as C
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_f = C.POINTER(C.c_float)
_vp = C.c_void_p
_u32 = C.c_uint32


def available(tag="O3"):
    return os.path.exists(os.path.join(_DIR, f"libref_cuda_{tag}.so"))


def _cm(m):
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T.reshape(-1))


def _fp(a):
    return a.ctypes.data_as(_f)


class RefLib:
    def __init__(self, tag="O3"):
        self.lib = lib = C.CDLL(os.path.join(_DIR, f"libref_cuda_{tag}.so"))
        lib.ref_volume_create.restype = _vp
        lib.ref_volume_create.argtypes = [_u32, _u32, _u32, C.c_float, C.c_float, C.c_float]
        lib.ref_volume_destroy.argtypes = [_vp]
        lib.ref_volume_offset.argtypes = [_vp, C.c_float, C.c_float, C.c_float]
        lib.ref_volume_clear.argtypes = [_vp]
        lib.ref_volume_trunc.restype = C.c_float
        lib.ref_volume_trunc.argtypes = [_vp]
        lib.ref_volume_voxel.argtypes = [_vp, _f]
        lib.ref_camera_matrices.argtypes = [_f, _f, _f, _f]
        lib.ref_volume_integrate.argtypes = [_vp, _vp, _u32, _u32, _f, _f]
        lib.ref_volume_raycast.argtypes = [_vp, _u32, _u32, _f, _f, _vp, _vp]
        lib.ref_volume_read.argtypes = [_vp, _vp, _vp]
        lib.ref_volume_set_distance_data.argtypes = [_vp, _vp]
        lib.ref_volume_read_deformation.argtypes = [_vp, _vp]
        lib.ref_volume_save.argtypes = [_vp, C.c_char_p]
        lib.ref_volume_load.restype = _vp
        lib.ref_volume_load.argtypes = [C.c_char_p]

    def camera_matrices(self, k, pose):
        """(kinv 3x3, inv_pose 4x4) exactly as the reference Camera derives them."""
        kinv = np.zeros(9, np.float32)
        ip = np.zeros(16, np.float32)
        self.lib.ref_camera_matrices(_fp(_cm(k)), _fp(_cm(pose)), _fp(kinv), _fp(ip))
        return kinv.reshape(3, 3).T.copy(), ip.reshape(4, 4).T.copy()


class RefVolume:
    def __init__(self, reflib, n, physical, handle=None):
        self.r = reflib
        self.n = tuple(int(x) for x in n) if n is not None else None
        self.h = handle if handle is not None else reflib.lib.ref_volume_create(*self.n, *[float(p) for p in physical])
        assert self.h, "reference TSDFVolume construction failed"

    @property
    def trunc(self):
        return np.float32(self.r.lib.ref_volume_trunc(self.h))

    @property
    def voxel(self):
        v = np.zeros(3, np.float32)
        self.r.lib.ref_volume_voxel(self.h, _fp(v))
        return v

    def offset(self, ox, oy, oz):
        self.r.lib.ref_volume_offset(self.h, ox, oy, oz)

    def clear(self):
        self.r.lib.ref_volume_clear(self.h)

    def integrate(self, depth, k, pose):
        h, w = depth.shape
        self.r.lib.ref_volume_integrate(self.h, depth.ctypes.data, w, h, _fp(_cm(k)), _fp(_cm(pose)))

    def raycast(self, w, h, k, pose):
        V = np.empty((h * w, 3), np.float32)
        N = np.empty((h * w, 3), np.float32)
        self.r.lib.ref_volume_raycast(self.h, w, h, _fp(_cm(k)), _fp(_cm(pose)), V.ctypes.data, N.ctypes.data)
        return V, N

    def read(self):
        nv = self.n[0] * self.n[1] * self.n[2]
        d, w = np.empty(nv, np.float32), np.empty(nv, np.float32)
        assert self.r.lib.ref_volume_read(self.h, d.ctypes.data, w.ctypes.data) == 0
        return d, w

    def read_deformation(self):
        nv = self.n[0] * self.n[1] * self.n[2]
        out = np.empty(nv * 6, np.float32)
        assert self.r.lib.ref_volume_read_deformation(self.h, out.ctypes.data) == 0
        return out

    def set_distance_data(self, d):
        d = np.ascontiguousarray(d, np.float32)
        self.r.lib.ref_volume_set_distance_data(self.h, d.ctypes.data)

    def save(self, path):
        assert self.r.lib.ref_volume_save(self.h, os.fsencode(path)) == 0

    def close(self):
        if self.h:
            self.r.lib.ref_volume_destroy(self.h)
            self.h = None
