"""Device-resident engine over the level-1 C-ABI: one process per GPU, the volume sharded along Z.

world == 1: the whole volume lives on this GPU; integrate and raycast are single kernel launches.
world  > 1: rank r owns the Z-slab [z0, z1) (multiples of the 8-voxel brick) plus one redundant halo
plane z1, fused by the same integrate kernel (every voxel depends only on itself and the frame, so the
halo is bit-identical to its owner's copy without communication).  Raycast: every rank marches all
rays through the samples whose interpolation cell starts in its slab and emits, per pixel, the key
(k_hit << 32 | float_bits(sample)); ONE all-reduce(min) over NVLink picks the first hit along each ray
(t_k is ray-independent, so the winning (k, sample) reproduces the single-GPU vertex bit for bit);
every rank then resolves keys to vertices and normals.

layout="replica" (world > 1) shards the IMAGE instead: slabs are dealt round robin, every rank pushes the surface bricks
it owns into a full-size copy of the distance volume on every GPU (peer stores over NVLink, tsdf_b200_bricks_push), then
marches its own pixel tiles with the single-GPU kernel and stores the vertices into every GPU's vertex map
(tsdf_b200_raycast_tiles).  Collectives: a max-reduce of the brick flags and two barriers per frame.

PyTorch provides device memory, the stream and torch.distributed — nothing else.
"""
import ctypes as C
import os

import numpy as np
import torch

from .capi import lib, check, fptr, fvec, colmajor, volume_params

BRICK = 8


def shard_ranges(nz, world, brick=BRICK):
    """Z-slab [z0, z1) owned by each rank: whole 8-voxel bricks, in rank order."""
    bricks = (nz + brick - 1) // brick
    per = (bricks + world - 1) // world
    return [(min(r * per * brick, nz), min((r + 1) * per * brick, nz)) for r in range(world)]


def interleaved_slabs(nz, world, rank, slab=16):
    """Global slabs [z0, z1) of `slab` planes dealt round robin: the ones rank `rank` owns, in ascending order."""
    n_slabs = (nz + slab - 1) // slab
    return [(s * slab, min((s + 1) * slab, nz)) for s in range(n_slabs) if s % world == rank]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class _RawF32:
    """A raw device pointer dressed for torch.as_tensor (peer-shareable cudaMalloc blocks are not torch allocations)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _peer_alloc(count):
    p = C.c_void_p()
    h = C.create_string_buffer(64)
    check(lib.tsdf_b200_peer_alloc(count * 4, C.byref(p), h), "peer_alloc")
    return p.value, h.raw


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


def _on_stream(method):
    """Runs a ShardedEngine method with the engine's stream as torch's current stream, so that the torch-side operations
    inside it (counter resets, collectives, .item() reads, copies) are ordered against the C-ABI kernel launches."""
    import functools

    @functools.wraps(method)
    def wrapped(self, *a, **kw):
        if torch.cuda.current_stream().cuda_stream == self._tstream.cuda_stream:
            return method(self, *a, **kw)
        with torch.cuda.stream(self._tstream):
            return method(self, *a, **kw)
    return wrapped


class ShardedEngine:
    def __init__(self, n, physical, rank=0, world=1, stream=None, skipping=True, stage_depth=True, layout="contiguous", slab=16,
                 exchange="allreduce"):
        """layout (world > 1): "contiguous" — one slab per rank (default); "interleaved" — global slabs of `slab` planes
        dealt round robin so that surfaces spread over the ranks.  Interleaving is exact (tests) but measured slower at
        512^3 on 2-8 GPUs: the integrate turns into many sub-wave launches and every rank walks
        every ray end to end.
        exchange (contiguous layout): "allreduce" — every rank writes a key map, NCCL min-reduces it, every rank resolves
        (every rank ends up with the vertex map); "peer" — the march min-merges its hits straight into rank 0's key map
        with atomics over NVLink peer memory (tsdf_b200_raycast_slab_min, CUDA IPC mapping), one 4-byte all-reduce serves
        as the barrier, and only rank 0 resolves (only rank 0 holds the vertex and normal maps)."""
        self.n = tuple(int(x) for x in n)
        self.physical = fvec(physical)
        self.rank, self.world = rank, world
        self.voxel, self.trunc = volume_params(self.n, self.physical)
        self.offset = np.zeros(3, np.float32)
        self.offset_at_clear = np.zeros(3, np.float32)
        # Every torch-side operation of the engine (counter resets, collectives, D2H reads) must be ordered against the
        # C-ABI launches: they all run with `self._tstream` current, which IS the stream the kernels are launched on.
        # `stream`: None (torch's current stream), a torch.cuda.Stream, or a raw cudaStream_t as an int (0 = legacy default).
        if stream is None:
            self._tstream = torch.cuda.current_stream()
        elif isinstance(stream, torch.cuda.Stream):
            self._tstream = stream
        elif int(stream) == 0:
            self._tstream = torch.cuda.default_stream()
        else:
            self._tstream = torch.cuda.ExternalStream(int(stream))
        self.stream = C.c_void_p(self._tstream.cuda_stream)
        self.skipping = skipping
        self.stage_depth = stage_depth
        self._staged = None
        nz = self.n[2]
        self.layout = layout if world > 1 else "contiguous"
        self.slab = slab
        if self.layout in ("interleaved", "replica"):
            assert slab % BRICK == 0
            self.slabs = interleaved_slabs(nz, world, rank, slab)          # [(z0, z1)] owned, each stored with a halo plane
            planes = max(len(self.slabs), 1) * (slab + 1)
            self.z0, self.z1 = (self.slabs[0][0], self.slabs[-1][1]) if self.slabs else (0, 0)
            self.zs1 = self.z1
            occ_bytes = lib.tsdf_b200_occupancy_bytes(*self.n)           # whole-volume brick grid, only own bricks get flagged
        else:
            self.z0, self.z1 = shard_ranges(nz, world)[rank]
            self.zs1 = min(self.z1 + 1, nz) if world > 1 else nz        # stored planes [z0, zs1)
            planes = max(self.zs1 - self.z0, 1)
            occ_bytes = lib.tsdf_b200_occupancy_bytes(self.n[0], self.n[1], planes)
        nvl = self.n[0] * self.n[1] * planes
        self.local_n = (self.n[0], self.n[1], planes)
        self.dist = torch.empty(nvl, dtype=torch.float32, device="cuda")
        self.weight = torch.empty(nvl, dtype=torch.float32, device="cuda")
        self.occ = torch.empty(occ_bytes, dtype=torch.uint8, device="cuda")
        self.table = torch.empty(4416, dtype=torch.float32, device="cuda")
        self.counters = torch.zeros(2, dtype=torch.int64, device="cuda")
        self._pix = 0
        check(lib.tsdf_b200_ray_table(self.trunc, _ptr(self.table), self.stream), "ray_table")
        self.fastdiv = 1
        for b in sorted(set(float(x) for x in self.voxel)):
            bad = C.c_ulonglong(1)
            check(lib.tsdf_b200_selftest_division(np.float32(b), C.byref(bad)), "selftest_division")
            if bad.value:
                self.fastdiv = 0
        self.replica = None
        self.exchange = exchange if (world > 1 and self.layout == "contiguous") else "allreduce"
        self._keymaps = None              # peer exchange: rank 0's two key maps (alternating frames) as seen from this process
        self._frame = 0
        self._opened = []
        if self.layout == "replica":
            # full-size copy of the distance volume; only surface bricks are ever written (by every rank's push) or read
            self.replica, self.replica_handle = _peer_alloc(self.n[0] * self.n[1] * self.n[2])
            check(lib.tsdf_b200_fill_f32(C.c_void_p(self.replica), self.n[0] * self.n[1] * self.n[2], self.trunc, self.stream), "fill")
            self.replicas = None          # every rank's replica as seen from this process, in rank order (connect())
            self.vmaps = None
            self.vmap = None
            self._token = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.clear()
        # kernels per step: depth pyramid (one cluster launch) + integrate + brick distances (one launch from bit rows, two byte
        # kernels for grids of 128 x 128 bricks per slice) + march (which finishes its set-aside rays itself) + normals;
        # sharded: + halo integrate + resolve (the barrier is NCCL's)
        wpr, rpl = (self.local_n[0] // BRICK + 63) // 64, (self.local_n[1] // BRICK + 31) // 32
        dist_launches = 1 if (wpr <= 2 and rpl <= 4 and not (wpr == 2 and rpl > 2)) else 2
        self.launches_per_step = (3 if world == 1 else 5) + dist_launches + (1 if stage_depth else 0)
        if self.layout in ("interleaved", "replica"):
            self.launches_per_step += 2 * (len(self.slabs) - 1)          # one integrate + one halo integrate per owned slab

    # ------------------------------------------------------------------------------------------
    @_on_stream
    def clear(self):
        if self.layout in ("interleaved", "replica"):
            check(lib.tsdf_b200_clear(_ptr(self.dist), _ptr(self.weight), *self.local_n, self.trunc, None, self.stream), "clear")
            self.occ.zero_()
            if self.replica is not None:
                # the replica keeps whatever was pushed before the clear, and after it those bricks are no longer flagged,
                # hence no longer pushed — yet the march still reads some of them (the low-edge layer; apron voxels of
                # unflagged neighbours of a flagged brick): back to the fill value.  Ordering against the peers' pushes:
                # their last push landed before the barrier that ended the previous raycast on this rank, and their next
                # one follows the flag reduction of the next raycast, which this rank enters after this fill (same stream).
                nv = self.n[0] * self.n[1] * self.n[2]
                check(lib.tsdf_b200_fill_f32(C.c_void_p(self.replica), nv, self.trunc, self.stream), "fill")
        else:
            check(lib.tsdf_b200_clear(_ptr(self.dist), _ptr(self.weight), *self.local_n, self.trunc, _ptr(self.occ),
                                      self.stream), "clear")
        self.offset_at_clear = self.offset.copy()

    def _buffers(self, w, h):
        if self._pix != w * h:
            self._pix = w * h
            self.vertices = torch.empty(w * h * 3, dtype=torch.float32, device="cuda")
            self.normals = torch.empty(w * h * 3, dtype=torch.float32, device="cuda")
            self.keys = torch.empty(w * h, dtype=torch.int64, device="cuda")
            # per-tile counters of the fused raycast (zero between launches)
            self.tile_counters = torch.zeros(lib.tsdf_b200_raycast_tile_counters(w, h), dtype=torch.int32, device="cuda")

    @staticmethod
    def _mats(cam):
        """Column-major float* views of the camera matrices, cached on the camera object."""
        m = getattr(cam, "_cabi", None)
        if m is None:
            pose = np.asarray(cam.pose, np.float32)
            arrs = [colmajor(cam.inv_pose), colmajor(cam.k), colmajor(cam.kinv), fvec(pose[:3, 3]), colmajor(pose[:3, :3])]
            m = tuple(fptr(a) for a in arrs) + (arrs,)      # keep the arrays alive
            cam._cabi = m
        return m

    def stage(self, d_depth):
        """Build the culling pyramid of a depth frame (tsdf_b200_depth_stage); integrate() does it unless restage=False."""
        h, w = d_depth.shape
        need = lib.tsdf_b200_depth_staged_bytes(w, h)
        if self._staged is None or self._staged.numel() * 4 < need:
            self._staged = torch.empty((need + 3) // 4, dtype=torch.float32, device="cuda")
        check(lib.tsdf_b200_depth_stage(_ptr(d_depth), w, h, _ptr(self._staged), self.stream), "depth_stage")

    @_on_stream
    def integrate(self, d_depth, cam, count=False, restage=True):
        """d_depth: (H, W) uint16 CUDA tensor.  Returns voxels rewritten (owned planes only) when count.
        restage=False reuses the staged frame of the previous call (same depth frame; kernel-timing aid)."""
        h, w = d_depth.shape
        if count:
            self.counters[0] = 0
        mats = self._mats(cam)[:3]
        staged = None
        if self.stage_depth:
            if restage:
                self.stage(d_depth)
            staged = _ptr(self._staged)
        cnt = C.c_void_p(self.counters.data_ptr()) if count else None
        if self.layout in ("interleaved", "replica"):
            nx, ny, nz = self.n
            plane = nx * ny
            bricks_per_layer = ((nx + BRICK - 1) // BRICK) * ((ny + BRICK - 1) // BRICK)
            for j, (z0, z1) in enumerate(self.slabs):
                d = C.c_void_p(self.dist.data_ptr() + 4 * plane * j * (self.slab + 1))
                wgt = C.c_void_p(self.weight.data_ptr() + 4 * plane * j * (self.slab + 1))
                occ = C.c_void_p(self.occ.data_ptr() + bricks_per_layer * (z0 // BRICK))
                own = z1 - z0
                stored = own + (1 if z1 < nz else 0)
                # the slab's array holds `stored` planes; plane 0 is global plane z0
                check(lib.tsdf_b200_integrate(d, wgt, None, nx, ny, stored, fptr(self.voxel), fptr(self.offset_at_clear),
                                              fptr(self.offset), self.trunc, *mats, w, h, _ptr(d_depth), staged, 0, own, z0, occ, cnt,
                                              self.stream), "integrate")
                if stored > own:      # redundant halo plane, not counted
                    check(lib.tsdf_b200_integrate(d, wgt, None, nx, ny, stored, fptr(self.voxel), fptr(self.offset_at_clear),
                                                  fptr(self.offset), self.trunc, *mats, w, h, _ptr(d_depth), staged, own, stored, z0, occ,
                                                  None, self.stream), "integrate halo")
            if count:
                return int(self.counters[0].item())
            return None
        own, stored = self.z1 - self.z0, self.zs1 - self.z0
        if own > 0:
            check(lib.tsdf_b200_integrate(_ptr(self.dist), _ptr(self.weight), None, *self.local_n, fptr(self.voxel),
                                          fptr(self.offset_at_clear), fptr(self.offset), self.trunc, *mats,
                                          w, h, _ptr(d_depth), staged, 0, own, self.z0, _ptr(self.occ), cnt, self.stream),
                  "integrate")
        if stored > own:      # redundant halo plane, not counted
            check(lib.tsdf_b200_integrate(_ptr(self.dist), _ptr(self.weight), None, *self.local_n, fptr(self.voxel),
                                          fptr(self.offset_at_clear), fptr(self.offset), self.trunc, *mats,
                                          w, h, _ptr(d_depth), staged, own, stored, self.z0, _ptr(self.occ), None, self.stream),
                  "integrate halo")
        if count:
            return int(self.counters[0].item())
        return None

    @_on_stream
    def march(self, w, h, cam, count=False):
        """Sharded raycast, phase 1: this rank's keys (k_hit << 32 | sample bits, INT64_MAX = no hit in my planes)."""
        self._buffers(w, h)
        _, _, kinv_p, origin_p, rot_p, _ = self._mats(cam)
        smin = self.offset.copy()
        smax = (self.offset + self.physical).astype(np.float32)
        cnt = C.c_void_p(self.counters.data_ptr() + 8) if count else None
        occ = _ptr(self.occ) if self.skipping else None
        if self.layout == "interleaved":
            check(lib.tsdf_b200_raycast_interleaved(_ptr(self.dist), *self.n, self.slab, self.world, self.rank,
                                                    fptr(self.voxel), fptr(smin), fptr(smax), self.trunc, origin_p, rot_p, kinv_p,
                                                    w, h, _ptr(self.table), occ, _ptr(self.keys), cnt, self.fastdiv, self.stream),
                  "raycast_interleaved")
        else:
            check(lib.tsdf_b200_raycast_slab(_ptr(self.dist), *self.n, self.z0, self.local_n[2], self.z0, self.z1,
                                             fptr(self.voxel), fptr(smin), fptr(smax), self.trunc, origin_p, rot_p, kinv_p,
                                             w, h, _ptr(self.table), occ, _ptr(self.keys), cnt, self.fastdiv, self.stream),
                  "raycast_slab")
        return self.keys

    @_on_stream
    def resolve(self, w, h, cam):
        """Sharded raycast, phase 2: min-reduced keys (in self.keys) -> vertices and normals."""
        _, _, kinv_p, origin_p, rot_p, _ = self._mats(cam)
        smin = self.offset.copy()
        smax = (self.offset + self.physical).astype(np.float32)
        check(lib.tsdf_b200_raycast_resolve(_ptr(self.keys), fptr(smin), fptr(smax), self.trunc, origin_p, rot_p,
                                            kinv_p, w, h, _ptr(self.table), _ptr(self.vertices), None, self.stream), "resolve")
        check(lib.tsdf_b200_normals(w, h, _ptr(self.vertices), _ptr(self.normals), self.stream), "normals")

    # ---- layout="replica": image-sharded raycast over peer memory ------------------------------------------------------
    @_on_stream
    def connect(self, w, h, peers=None):
        """Allocates this rank's shareable vertex map and learns every rank's replica and vertex map.  peers: the engines
        of all ranks when they live in this process (tests emulate the ranks on one GPU); otherwise the CUDA IPC handles
        travel through torch.distributed (collective call)."""
        if self.vmap is None or self._pix != w * h:
            self._pix = w * h
            self.vmap, self.vmap_handle = _peer_alloc(w * h * 3)
            self.vertices = torch.as_tensor(_RawF32(self.vmap, w * h * 3), device="cuda")
            self.normals = torch.empty(w * h * 3, dtype=torch.float32, device="cuda")
            self.replicas = None
        if self.replicas is not None:
            return
        if peers is not None:
            if any(e.vmap is None or e._pix != w * h for e in peers):
                return                                      # the last engine to allocate completes everyone's lists
            for e in peers:
                e.replicas = [q.replica for q in peers]
                e.vmaps = [q.vmap for q in peers]
            return
        import torch.distributed as dist
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.replica_handle, self.vmap_handle))
        self.replicas, self.vmaps = [], []
        for r, (hr, hv) in enumerate(handles):
            if r == self.rank:
                self.replicas.append(self.replica)
                self.vmaps.append(self.vmap)
                continue
            for hnd, out in ((hr, self.replicas), (hv, self.vmaps)):
                p = C.c_void_p()
                check(lib.tsdf_b200_peer_open(hnd, C.byref(p)), "peer_open")
                self._opened.append(p.value)
                out.append(p.value)

    def flags(self):
        """The brick flags of the whole volume (first third of the occupancy buffer): what the ranks max-reduce."""
        nb = ((self.n[0] + BRICK - 1) // BRICK) * ((self.n[1] + BRICK - 1) // BRICK) * ((self.n[2] + BRICK - 1) // BRICK)
        return self.occ[:nb]

    @_on_stream
    def push(self, count=False):
        """Copies the owned surface bricks into every rank's replica (flags must be merged first)."""
        if count:
            self.counters[1] = 0
        cnt = C.c_void_p(self.counters.data_ptr() + 8) if count else None
        check(lib.tsdf_b200_bricks_push(_ptr(self.dist), *self.n, self.slab, self.world, self.rank, _ptr(self.occ),
                                        len(self.replicas), _ptr_array(self.replicas), cnt, self.stream), "bricks_push")
        if count:
            return int(self.counters[1].item())
        return None

    @_on_stream
    def march_tiles(self, w, h, cam, count=False):
        """Marches this rank's pixel tiles against its replica and stores the vertices into every rank's vertex map."""
        _, _, kinv_p, origin_p, rot_p, _ = self._mats(cam)
        smin = self.offset.copy()
        smax = (self.offset + self.physical).astype(np.float32)
        cnt = C.c_void_p(self.counters.data_ptr() + 8) if count else None
        check(lib.tsdf_b200_raycast_tiles(C.c_void_p(self.replica), *self.n, fptr(self.voxel), fptr(smin), fptr(smax), self.trunc,
                                          origin_p, rot_p, kinv_p, w, h, _ptr(self.table), _ptr(self.occ) if self.skipping else None,
                                          self.world, self.rank, len(self.vmaps), _ptr_array(self.vmaps), cnt, self.fastdiv,
                                          self.stream), "raycast_tiles")

    def finish(self, w, h):
        check(lib.tsdf_b200_normals(w, h, _ptr(self.vertices), _ptr(self.normals), self.stream), "normals")

    def _barrier(self):
        import torch.distributed as dist
        dist.all_reduce(self._token)          # stream-ordered; the host does not wait

    def _connect_keys(self, w, h):
        """Peer exchange: rank 0 allocates two shareable key maps (INT64_MAX everywhere), the others map them (collective)."""
        import torch.distributed as dist
        if self._keymaps is not None and self._keypix == w * h:
            return
        self._keypix = w * h
        handles = [None, None]
        if self.rank == 0:
            own = []
            for i in range(2):
                p, hnd = _peer_alloc(w * h * 2)          # count is in 4-byte units: 8 bytes per key
                check(lib.tsdf_b200_fill_i64(C.c_void_p(p), w * h, 0x7fffffffffffffff, self.stream), "fill_i64")
                own.append(p)
                handles[i] = hnd
            self._own_keymaps = own
            torch.cuda.synchronize()
        box = [handles]
        dist.broadcast_object_list(box, src=0)
        if self.rank == 0:
            self._keymaps = self._own_keymaps
        else:
            self._keymaps = []
            for hnd in box[0]:
                p = C.c_void_p()
                check(lib.tsdf_b200_peer_open(hnd, C.byref(p)), "peer_open")
                self._opened.append(p.value)
                self._keymaps.append(p.value)
        self._token = torch.zeros(1, dtype=torch.int32, device="cuda")

    def _raycast_peer(self, w, h, cam, count):
        """One frame of the peer exchange.  Key map f & 1 is written by every rank's march of frame f and read (and reset)
        by rank 0's resolve of frame f; the march of frame f + 2 follows the barrier of frame f + 1, which rank 0 enters
        after that resolve (stream order) — so no rank can write a map that is still being read."""
        import torch.distributed as dist
        self._connect_keys(w, h)
        self._buffers(w, h)
        _, _, kinv_p, origin_p, rot_p, _ = self._mats(cam)
        smin = self.offset.copy()
        smax = (self.offset + self.physical).astype(np.float32)
        cnt = C.c_void_p(self.counters.data_ptr() + 8) if count else None
        occ = _ptr(self.occ) if self.skipping else None
        keys = C.c_void_p(self._keymaps[self._frame & 1])
        self._frame += 1
        check(lib.tsdf_b200_raycast_slab_min(_ptr(self.dist), *self.n, self.z0, self.local_n[2], self.z0, self.z1,
                                             fptr(self.voxel), fptr(smin), fptr(smax), self.trunc, origin_p, rot_p, kinv_p,
                                             w, h, _ptr(self.table), occ, keys, cnt, self.fastdiv, self.stream), "raycast_slab_min")
        dist.all_reduce(self._token)          # barrier on the stream: every rank's atomics have landed
        if self.rank == 0:
            check(lib.tsdf_b200_raycast_resolve_reset(keys, fptr(smin), fptr(smax), self.trunc, origin_p, rot_p, kinv_p, w, h,
                                                      _ptr(self.table), _ptr(self.vertices), None, self.stream), "resolve_reset")
            check(lib.tsdf_b200_normals(w, h, _ptr(self.vertices), _ptr(self.normals), self.stream), "normals")

    @_on_stream
    def raycast(self, w, h, cam, count=False):
        if self.layout == "replica":
            import torch.distributed as dist
            self.connect(w, h)
            # every rank has left the previous frame's march (it reads the replicas) before anyone's push can start:
            # the reduction cannot complete on a rank before all ranks have entered it
            dist.all_reduce(self.flags(), op=dist.ReduceOp.MAX)
            self.push()
            self._barrier()                   # all pushes have landed
            if count:
                self.counters[1] = 0
            self.march_tiles(w, h, cam, count)
            self._barrier()                   # all vertex tiles have landed
            self.finish(w, h)
            if count:
                return int(self.counters[1].item())
            return None
        self._buffers(w, h)
        if count:
            self.counters[1] = 0
        if self.world == 1:
            _, _, kinv_p, origin_p, rot_p, _ = self._mats(cam)
            smin = self.offset.copy()
            smax = (self.offset + self.physical).astype(np.float32)
            cnt = C.c_void_p(self.counters.data_ptr() + 8) if count else None
            occ = _ptr(self.occ) if self.skipping else None
            if os.environ.get("TSDF_B200_FUSED", "0") == "1":
                # march + normals in one launch sequence (tsdf_b200_raycast_fused: what the level-2 volume uses to stream both
                # maps into pinned host buffers while the march runs).  With the maps staying on the device it is SLOWER than
                # the march followed by the normals kernel (292 against 275 us on the bench frames: every tile ends with a
                # fence and an atomic round trip), hence not the default here.
                check(lib.tsdf_b200_raycast_fused(_ptr(self.dist), *self.n, fptr(self.voxel), fptr(smin), fptr(smax), self.trunc,
                                                  origin_p, rot_p, kinv_p, w, h, _ptr(self.table), occ, _ptr(self.vertices),
                                                  _ptr(self.normals), None, None, _ptr(self.tile_counters), cnt, self.fastdiv,
                                                  self.stream), "raycast_fused")
            else:
                check(lib.tsdf_b200_raycast_ex(_ptr(self.dist), *self.n, fptr(self.voxel), fptr(smin), fptr(smax), self.trunc,
                                               origin_p, rot_p, kinv_p, w, h, _ptr(self.table), occ, _ptr(self.vertices), None,
                                               cnt, self.fastdiv, self.stream), "raycast")
                check(lib.tsdf_b200_normals(w, h, _ptr(self.vertices), _ptr(self.normals), self.stream), "normals")
        elif self.exchange == "peer":
            self._raycast_peer(w, h, cam, count)
        else:
            import torch.distributed as dist
            self.march(w, h, cam, count)
            dist.all_reduce(self.keys, op=dist.ReduceOp.MIN)      # the one exchange: first hit along every ray
            self.resolve(w, h, cam)
        if count:
            return int(self.counters[1].item())
        return None

    @_on_stream
    def extract_mesh(self):
        """Marching cubes of this rank's owned planes (tsdf_b200_mc_extract; the halo plane closes the cubes at the slab's
        upper face).  Returns a (n, 3) float32 CUDA tensor: three consecutive vertices per triangle, in the reference's
        order; the whole mesh is the concatenation over the ranks in rank order.  Contiguous layout only."""
        assert self.layout == "contiguous"
        own = self.z1 - self.z0
        out = C.c_void_p()
        count = C.c_ulonglong()
        # cubes based in the owned planes; the call clamps the range to the stored planes minus one (a cube needs plane z+1)
        check(lib.tsdf_b200_mc_extract(_ptr(self.dist), *self.local_n, self.z0, 0, own, fptr(self.voxel),
                                       fptr(self.offset), C.byref(out), C.byref(count), self.stream), "mc_extract")
        mesh = torch.empty((count.value, 3), dtype=torch.float32, device="cuda")
        if count.value:
            mesh.copy_(torch.as_tensor(_RawF32(out.value, count.value * 3), device="cuda").view(-1, 3))
            torch.cuda.synchronize()
            lib.tsdf_b200_device_free(out)
        return mesh

    @_on_stream
    def last_ray_stats(self):
        """Hit pixels / NaN pixels of the last raycast (diagnostics for the bench line)."""
        if self._pix == 0:
            return None
        v = self.vertices.view(-1, 3)[:, 0]
        hits = int((~torch.isnan(v)).sum().item())
        return {"hit_pixels": hits, "pixels": self._pix}

    @_on_stream
    def e2e(self, frames, cams, warmup, steps, w, h):
        """Host-buffer loop for the sharded case: pinned depth H2D on every rank, result D2H on rank 0."""
        import time
        import torch.distributed as dist
        pin = [torch.from_numpy(f).pin_memory() for f in frames]
        d = torch.empty((h, w), dtype=torch.uint16, device="cuda")
        hv = torch.empty(w * h * 3, dtype=torch.float32).pin_memory()
        hn = torch.empty(w * h * 3, dtype=torch.float32).pin_memory()

        def one(i):
            d.copy_(pin[i], non_blocking=True)
            self.integrate(d, cams[i])
            self.raycast(w, h, cams[i])
            if self.rank == 0:
                hv.copy_(self.vertices, non_blocking=True)
                hn.copy_(self.normals, non_blocking=True)
            torch.cuda.synchronize()

        for i in range(warmup):
            one(i)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(steps):
            one(warmup + s)
        dist.barrier()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        return {"value": 1e3 / ms, "unit": "frames/s", "h2d_bytes_per_step": w * h * 2 * self.world,
                "d2h_bytes_per_step": 2 * w * h * 3 * 4, "ms_per_step": ms,
                "api": "ShardedEngine: pinned depth H2D on every rank, integrate+raycast, vertex+normal D2H on rank 0"}

    @_on_stream
    def read_local(self):
        torch.cuda.synchronize()
        return self.dist.cpu().numpy(), self.weight.cpu().numpy()

    def close(self):
        torch.cuda.synchronize()
        if self._opened or getattr(self, "_own_keymaps", None):
            # (collective: every rank of a layout / exchange that shares memory comes through here, owner or not)
            import torch.distributed as dist
            for p in self._opened:
                lib.tsdf_b200_peer_close(C.c_void_p(p))
            self._opened = []
            if dist.is_initialized():
                dist.barrier()                # nobody frees a block that a peer still has mapped
        if getattr(self, "_own_keymaps", None):
            for p in self._own_keymaps:
                lib.tsdf_b200_peer_free(C.c_void_p(p))
            self._own_keymaps = None
        if self.replica is not None:
            self.vertices = None
            lib.tsdf_b200_peer_free(C.c_void_p(self.replica))
            if self.vmap is not None:
                lib.tsdf_b200_peer_free(C.c_void_p(self.vmap))
            self.replica = self.vmap = None
        for name in ("dist", "weight", "occ", "vertices", "normals", "keys"):
            if hasattr(self, name):
                setattr(self, name, None)
        torch.cuda.empty_cache()
