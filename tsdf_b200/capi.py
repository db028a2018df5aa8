"""ctypes binding of include/tsdf_b200.h.  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.environ.get("TSDF_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtsdf_b200.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first "
        "(python -c 'import __graft_entry__ as g; g.build()' or `make lib`). "
        "tsdf_b200 has no CPU fallback.")
lib = C.CDLL(LIB_PATH)

_f = C.POINTER(C.c_float)
_u8 = C.POINTER(C.c_uint8)
_u16 = C.POINTER(C.c_uint16)
_u32 = C.c_uint32
_ull = C.POINTER(C.c_ulonglong)
_vp = C.c_void_p

# name -> (restype, argtypes); also the list of symbols include/tsdf_b200.h declares.
SIGNATURES = {
    "tsdf_b200_version": (C.c_char_p, []),
    "tsdf_b200_strerror": (C.c_char_p, [C.c_int]),
    "tsdf_b200_volume_params": (C.c_int, [_u32, _u32, _u32, _f, _f, _f]),
    "tsdf_b200_clear": (C.c_int, [_vp, _vp, _u32, _u32, _u32, C.c_float, _vp, _vp]),
    "tsdf_b200_init_deformation": (C.c_int, [_vp, _u32, _u32, _u32, _f, _f, _vp]),
    "tsdf_b200_depth_staged_bytes": (C.c_size_t, [_u32, _u32]),
    "tsdf_b200_depth_stage": (C.c_int, [_vp, _u32, _u32, _vp, _vp]),
    "tsdf_b200_integrate": (C.c_int, [_vp, _vp, _vp, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f,
                                      _u32, _u32, _vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "tsdf_b200_debug_force_generic_integrate": (None, [C.c_int]),
    "tsdf_b200_debug_integrate_variant": (None, [C.c_int]),
    "tsdf_b200_occupancy_bytes": (C.c_size_t, [_u32, _u32, _u32]),
    "tsdf_b200_occupancy_rebuild": (C.c_int, [_vp, _u32, _u32, _u32, C.c_float, _vp, _vp]),
    "tsdf_b200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "tsdf_b200_host_free": (None, [_vp]),
    "tsdf_b200_ray_table": (C.c_int, [C.c_float, _vp, _vp]),
    "tsdf_b200_raycast": (C.c_int, [_vp, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f, _u32, _u32,
                                    _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsdf_b200_raycast_ex": (C.c_int, [_vp, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f, _u32, _u32,
                                       _vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "tsdf_b200_raycast_mirrored": (C.c_int, [_vp, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f, _u32, _u32, _vp, _vp,
                                             _vp, _vp, _vp, C.c_int, _vp]),
    "tsdf_b200_raycast_fused": (C.c_int, [_vp, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f, _u32, _u32, _vp, _vp,
                                          _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "tsdf_b200_raycast_tile_counters": (C.c_size_t, [_u32, _u32]),
    "tsdf_b200_raycast_slab": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f,
                                         _u32, _u32, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "tsdf_b200_raycast_slab_min": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f,
                                             _u32, _u32, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "tsdf_b200_raycast_resolve_reset": (C.c_int, [_vp, _f, _f, C.c_float, _f, _f, _f, _u32, _u32, _vp, _vp, _vp, _vp]),
    "tsdf_b200_fill_i64": (C.c_int, [_vp, C.c_size_t, C.c_longlong, _vp]),
    "tsdf_b200_raycast_interleaved": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f,
                                                _u32, _u32, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "tsdf_b200_bricks_push": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _u32, C.POINTER(_vp), _vp, _vp]),
    "tsdf_b200_raycast_tiles": (C.c_int, [_vp, _u32, _u32, _u32, _f, _f, _f, C.c_float, _f, _f, _f, _u32, _u32, _vp, _vp,
                                          _u32, _u32, _u32, C.POINTER(_vp), _vp, C.c_int, _vp]),
    "tsdf_b200_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp), C.c_char_p]),
    "tsdf_b200_peer_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "tsdf_b200_peer_close": (C.c_int, [_vp]),
    "tsdf_b200_peer_free": (C.c_int, [_vp]),
    "tsdf_b200_fill_f32": (C.c_int, [_vp, C.c_size_t, C.c_float, _vp]),
    "tsdf_b200_raycast_resolve": (C.c_int, [_vp, _f, _f, C.c_float, _f, _f, _f, _u32, _u32, _vp, _vp, _vp, _vp]),
    "tsdf_b200_normals": (C.c_int, [_u32, _u32, _vp, _vp, _vp]),
    "tsdf_b200_selftest_division": (C.c_int, [C.c_float, _ull]),
    "tsdf_b200_bilateral_u8": (C.c_int, [_vp, _vp, _u32, _u32, _vp, _u32, _vp, _u32, _vp]),
    "tsdf_b200_bilateral_u16": (C.c_int, [_vp, _vp, _u32, _u32, _vp, _u32, _vp, _u32, _vp]),
    "tsdf_b200_bilateral_host": (C.c_int, [_vp, C.c_int, _u32, _u32, _f, _u32, _f, _u32]),
    "tsdf_b200_mc_extract": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _u32, _u32, _f, _f, C.POINTER(_vp), _ull, _vp]),
    "tsdf_b200_device_free": (None, [_vp]),
    "tsdf_b200_copy_to_host": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tsdf_b200_volume_create": (C.c_int, [_u32, _u32, _u32, C.c_float, C.c_float, C.c_float, C.POINTER(_vp)]),
    "tsdf_b200_volume_create_sharded": (C.c_int, [_u32, _u32, _u32, C.c_float, C.c_float, C.c_float, C.c_int, C.POINTER(_vp)]),
    "tsdf_b200_volume_gpus": (C.c_int, [_vp]),
    "tsdf_b200_volume_extract_mesh": (C.c_int, [_vp, C.POINTER(_vp), _ull]),
    "tsdf_b200_volume_load": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "tsdf_b200_volume_destroy": (None, [_vp]),
    "tsdf_b200_volume_get": (C.c_int, [_vp, C.POINTER(_u32), _f, _f, _f, _f, _f]),
    "tsdf_b200_volume_get_global": (C.c_int, [_vp, _f, _f]),
    "tsdf_b200_volume_set_offset": (C.c_int, [_vp, C.c_float, C.c_float, C.c_float]),
    "tsdf_b200_volume_clear": (C.c_int, [_vp]),
    "tsdf_b200_volume_distance_data": (_vp, [_vp]),
    "tsdf_b200_volume_weight_data": (_vp, [_vp]),
    "tsdf_b200_volume_deformation": (_vp, [_vp]),
    "tsdf_b200_volume_set_distance_data": (C.c_int, [_vp, _vp]),
    "tsdf_b200_volume_set_weight_data": (C.c_int, [_vp, _vp]),
    "tsdf_b200_volume_set_deformation": (C.c_int, [_vp, _vp]),
    "tsdf_b200_volume_read": (C.c_int, [_vp, _vp, _vp]),
    "tsdf_b200_volume_integrate": (C.c_int, [_vp, _vp, _u32, _u32, _f, _f, _f]),
    "tsdf_b200_volume_raycast": (C.c_int, [_vp, _u32, _u32, _f, _f, _vp, _vp]),
    "tsdf_b200_volume_save": (C.c_int, [_vp, C.c_char_p]),
    "tsdf_b200_volume_stats": (C.c_int, [_vp, _ull, _ull]),
    "tsdf_b200_volume_set_skipping": (C.c_int, [_vp, C.c_int]),
}
for _name, (_res, _args) in SIGNATURES.items():
    if os.environ.get("TSDF_B200_LIB") and not hasattr(lib, _name):
        continue                       # bisecting with an older build of the library (tuning aid)
    _fn = getattr(lib, _name)          # AttributeError here = symbol missing from the library
    _fn.restype = _res
    _fn.argtypes = _args


class TsdfError(RuntimeError):
    pass


def check(code, what=""):
    if code != 0:
        raise TsdfError(f"{what}: tsdf_b200 error {code}: {lib.tsdf_b200_strerror(code).decode()}")


def fptr(a):
    """float* of a contiguous float32 numpy array."""
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f)


def fvec(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))


def colmajor(m):
    """Column-major float32 flattening of a matrix (Eigen's .data())."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T.reshape(-1))


def volume_params(n, physical):
    vox = np.zeros(3, np.float32)
    tr = C.c_float()
    check(lib.tsdf_b200_volume_params(n[0], n[1], n[2], fptr(fvec(physical)), fptr(vox), C.byref(tr)), "volume_params")
    return vox, np.float32(tr.value)


class Volume:
    """Level-2 handle: the C-ABI equivalent of the reference's TSDFVolume object (host buffers)."""

    def __init__(self, n, physical, handle=None, gpus=None):
        self._h = _vp()
        if handle is not None:
            self._h = handle
        elif gpus is not None:
            check(lib.tsdf_b200_volume_create_sharded(n[0], n[1], n[2], physical[0], physical[1], physical[2], int(gpus),
                                                      C.byref(self._h)), "volume_create_sharded")
        else:
            check(lib.tsdf_b200_volume_create(n[0], n[1], n[2], physical[0], physical[1], physical[2],
                                              C.byref(self._h)), "volume_create")
        size = (_u32 * 3)()
        phys, vox, off = (np.zeros(3, np.float32) for _ in range(3))
        tr, mw = C.c_float(), C.c_float()
        check(lib.tsdf_b200_volume_get(self._h, size, fptr(phys), fptr(vox), fptr(off), C.byref(tr), C.byref(mw)))
        self.size = tuple(int(s) for s in size)
        self.physical, self.voxel = phys, vox
        self.trunc, self.max_weight = np.float32(tr.value), np.float32(mw.value)

    @classmethod
    def load(cls, path):
        h = _vp()
        check(lib.tsdf_b200_volume_load(os.fsencode(path), C.byref(h)), "volume_load")
        return cls(None, None, handle=h)

    @property
    def offset(self):
        off = np.zeros(3, np.float32)
        check(lib.tsdf_b200_volume_get(self._h, None, None, None, fptr(off), None, None))
        return off

    def set_offset(self, ox, oy, oz):
        check(lib.tsdf_b200_volume_set_offset(self._h, ox, oy, oz))

    def close(self):
        if self._h:
            lib.tsdf_b200_volume_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nvox(self):
        return self.size[0] * self.size[1] * self.size[2]

    def clear(self):
        check(lib.tsdf_b200_volume_clear(self._h), "clear")

    def integrate(self, depth, inv_pose, k, kinv):
        """depth: (H, W) uint16 host array (or an int address with shape given via .shape)."""
        h, w = depth.shape
        assert depth.dtype == np.uint16 and depth.flags["C_CONTIGUOUS"]
        check(lib.tsdf_b200_volume_integrate(self._h, depth.ctypes.data, w, h, fptr(colmajor(inv_pose)),
                                             fptr(colmajor(k)), fptr(colmajor(kinv))), "integrate")

    def raycast(self, w, h, pose, kinv, vertices=None, normals=None):
        if vertices is None:
            vertices = np.empty((h * w, 3), np.float32)
        if normals is None:
            normals = np.empty((h * w, 3), np.float32)
        check(lib.tsdf_b200_volume_raycast(self._h, w, h, fptr(colmajor(pose)), fptr(colmajor(kinv)),
                                           vertices.ctypes.data, normals.ctypes.data), "raycast")
        return vertices, normals

    def read(self):
        d = np.empty(self.nvox, np.float32)
        w = np.empty(self.nvox, np.float32)
        check(lib.tsdf_b200_volume_read(self._h, d.ctypes.data, w.ctypes.data), "read")
        return d, w

    def set_distance_data(self, d):
        d = np.ascontiguousarray(d, np.float32)
        assert d.size == self.nvox
        check(lib.tsdf_b200_volume_set_distance_data(self._h, d.ctypes.data))

    def set_weight_data(self, w):
        w = np.ascontiguousarray(w, np.float32)
        assert w.size == self.nvox
        check(lib.tsdf_b200_volume_set_weight_data(self._h, w.ctypes.data))

    def set_deformation(self, nodes):
        nodes = np.ascontiguousarray(nodes, np.float32)
        assert nodes.size == self.nvox * 6
        check(lib.tsdf_b200_volume_set_deformation(self._h, nodes.ctypes.data))

    def set_skipping(self, enabled):
        check(lib.tsdf_b200_volume_set_skipping(self._h, int(enabled)))

    def stats(self):
        a, b = C.c_ulonglong(), C.c_ulonglong()
        check(lib.tsdf_b200_volume_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def gpus(self):
        return lib.tsdf_b200_volume_gpus(self._h)

    def extract_mesh(self):
        """(n, 3) float32 mesh vertices (three per triangle) of extract_surface, copied to the host."""
        out = _vp()
        count = C.c_ulonglong()
        check(lib.tsdf_b200_volume_extract_mesh(self._h, C.byref(out), C.byref(count)), "extract_mesh")
        v = np.empty((count.value, 3), np.float32)
        if count.value:
            check(lib.tsdf_b200_copy_to_host(v.ctypes.data, out, v.nbytes), "copy_to_host")
            lib.tsdf_b200_device_free(out)
        return v

    def save(self, path):
        check(lib.tsdf_b200_volume_save(self._h, os.fsencode(path)), "save")

    @property
    def distance_ptr(self):
        return lib.tsdf_b200_volume_distance_data(self._h)

    @property
    def weight_ptr(self):
        return lib.tsdf_b200_volume_weight_data(self._h)
