"""Kernel-level (level-1 C-ABI) drivers for the GPU tests: torch only provides device memory."""
import ctypes as C

import numpy as np
import torch

from tsdf_b200 import capi
from tsdf_b200.capi import lib, check, fptr, fvec, colmajor


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class DeviceVolume:
    """Raw device arrays driven through the level-1 entry points (what the level-2 object does inside)."""

    def __init__(self, n, physical, offset=(0, 0, 0), with_deformation=False, with_occ=True):
        self.n = tuple(int(x) for x in n)
        self.physical = fvec(physical)
        self.voxel, self.trunc = capi.volume_params(self.n, self.physical)
        self.offset = fvec(offset)
        self.offset_at_clear = fvec(offset)
        nv = self.n[0] * self.n[1] * self.n[2]
        self.dist = torch.empty(nv, dtype=torch.float32, device="cuda")
        self.weight = torch.empty(nv, dtype=torch.float32, device="cuda")
        self.occ = torch.empty(lib.tsdf_b200_occupancy_bytes(*self.n), dtype=torch.uint8, device="cuda") if with_occ else None
        self.deform = torch.empty(nv * 6, dtype=torch.float32, device="cuda") if with_deformation else None
        self.table = torch.empty(4416, dtype=torch.float32, device="cuda")
        self.counter = torch.zeros(2, dtype=torch.int64, device="cuda")
        check(lib.tsdf_b200_ray_table(self.trunc, ptr(self.table), None))
        self.clear()

    def clear(self):
        check(lib.tsdf_b200_clear(ptr(self.dist), ptr(self.weight), *self.n, self.trunc, ptr(self.occ), None))
        if self.deform is not None:
            check(lib.tsdf_b200_init_deformation(ptr(self.deform), *self.n, fptr(self.voxel), fptr(self.offset_at_clear), None))
        torch.cuda.synchronize()

    def integrate(self, depth, inv_pose, k, kinv, z_begin=0, z_end=None, count=True, staged=True):
        h, w = depth.shape
        d = depth if isinstance(depth, torch.Tensor) else dev(depth)
        z_end = self.n[2] if z_end is None else z_end
        self.counter[0] = 0
        st = None
        if staged:
            st = torch.empty((lib.tsdf_b200_depth_staged_bytes(w, h) + 3) // 4, dtype=torch.float32, device="cuda")
            check(lib.tsdf_b200_depth_stage(ptr(d), w, h, ptr(st), None), "depth_stage")
        check(lib.tsdf_b200_integrate(ptr(self.dist), ptr(self.weight), ptr(self.deform), *self.n, fptr(self.voxel),
                                      fptr(self.offset_at_clear), fptr(self.offset), self.trunc, fptr(colmajor(inv_pose)),
                                      fptr(colmajor(k)), fptr(colmajor(kinv)), w, h, ptr(d), ptr(st), z_begin, z_end, 0, ptr(self.occ),
                                      C.c_void_p(self.counter.data_ptr()) if count else None, None), "integrate")
        torch.cuda.synchronize()
        return int(self.counter[0].item())

    def raycast(self, w, h, pose, kinv, skip=True, fastdiv=False):
        pose = np.asarray(pose, np.float32)
        V = torch.empty(h * w * 3, dtype=torch.float32, device="cuda")
        N = torch.empty(h * w * 3, dtype=torch.float32, device="cuda")
        kh = torch.empty(h * w, dtype=torch.int32, device="cuda")
        self.counter[1] = 0
        smin = self.offset.copy()
        smax = (self.offset + self.physical).astype(np.float32)
        fn = lib.tsdf_b200_raycast_ex
        check(fn(ptr(self.dist), *self.n, fptr(self.voxel), fptr(smin), fptr(smax), self.trunc, fptr(fvec(pose[:3, 3])),
                 fptr(colmajor(pose[:3, :3])), fptr(colmajor(kinv)), w, h, ptr(self.table),
                 ptr(self.occ) if skip else None, ptr(V), ptr(kh), C.c_void_p(self.counter.data_ptr() + 8),
                 int(fastdiv), None), "raycast")
        check(lib.tsdf_b200_normals(w, h, ptr(V), ptr(N), None), "normals")
        torch.cuda.synchronize()
        return (V.cpu().numpy().reshape(-1, 3), N.cpu().numpy().reshape(-1, 3), kh.cpu().numpy(),
                int(self.counter[1].item()))

    def upload_dist(self, d):
        self.dist.copy_(torch.from_numpy(np.ascontiguousarray(d, np.float32)))
        if self.occ is not None:
            check(lib.tsdf_b200_occupancy_rebuild(ptr(self.dist), *self.n, self.trunc, ptr(self.occ), None))
        torch.cuda.synchronize()


def raycast_sharded_on_one_gpu(dv, w, h, pose, kinv, world, skip=True, fastdiv=True):
    """Emulates the multi-GPU raycast on one device: the volume is cut into `world` Z-slabs (own planes + halo),
    each slab is marched by tsdf_b200_raycast_slab from its own copy and its own occupancy grid, the keys are
    min-reduced (what the NCCL all-reduce does) and resolved."""
    from tsdf_b200.sharded import shard_ranges
    pose = np.asarray(pose, np.float32)
    nx, ny, nz = dv.n
    smin = dv.offset.copy()
    smax = (dv.offset + dv.physical).astype(np.float32)
    keys = None
    total = 0
    full = dv.dist.view(nz, ny * nx)
    for z0, z1 in shard_ranges(nz, world):
        if z1 <= z0:
            continue
        zs1 = min(z1 + 1, nz)
        slab = full[z0:zs1].contiguous().view(-1)
        occ = torch.zeros(lib.tsdf_b200_occupancy_bytes(nx, ny, zs1 - z0), dtype=torch.uint8, device="cuda")
        check(lib.tsdf_b200_occupancy_rebuild(ptr(slab), nx, ny, zs1 - z0, dv.trunc, ptr(occ), None))
        k = torch.empty(h * w, dtype=torch.int64, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        check(lib.tsdf_b200_raycast_slab(ptr(slab), nx, ny, nz, z0, zs1 - z0, z0, z1, fptr(dv.voxel), fptr(smin), fptr(smax),
                                         dv.trunc, fptr(fvec(pose[:3, 3])), fptr(colmajor(pose[:3, :3])), fptr(colmajor(kinv)),
                                         w, h, ptr(dv.table), ptr(occ) if skip else None, ptr(k), ptr(cnt), int(fastdiv), None),
              "raycast_slab")
        torch.cuda.synchronize()
        total += int(cnt.item())
        keys = k if keys is None else torch.minimum(keys, k)
    V = torch.empty(h * w * 3, dtype=torch.float32, device="cuda")
    kh = torch.empty(h * w, dtype=torch.int32, device="cuda")
    check(lib.tsdf_b200_raycast_resolve(ptr(keys), fptr(smin), fptr(smax), dv.trunc, fptr(fvec(pose[:3, 3])),
                                        fptr(colmajor(pose[:3, :3])), fptr(colmajor(kinv)), w, h, ptr(dv.table), ptr(V), ptr(kh), None),
          "resolve")
    torch.cuda.synchronize()
    return V.cpu().numpy().reshape(-1, 3), kh.cpu().numpy(), total
