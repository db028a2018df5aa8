"""CPU-only tests of the multi-GPU decomposition (no CUDA): slab ownership, and a world_size-2 gloo run in which
each rank marches its slab with the CPU oracle, the keys are min-reduced with torch.distributed and resolved —
the result must equal the undivided raycast bit for bit."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_volume_in_whole_bricks():
    from oracle import oracle
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharded_src", os.path.join(ROOT, "tsdf_b200", "sharded.py"))
    src = open(os.path.join(ROOT, "tsdf_b200", "sharded.py")).read()
    ns = {}
    exec(src[src.index("def shard_ranges"):src.index("def _ptr")], {"BRICK": 8}, ns)     # host logic only, no CUDA import
    for nz in (1, 7, 8, 9, 64, 100, 512, 1024):
        for world in (1, 2, 3, 4, 8):
            r = ns["shard_ranges"](nz, world)
            assert r == oracle.shard_ranges(nz, world)
            assert len(r) == world and r[0][0] == 0 and max(b for _, b in r) == nz
            for (a0, a1), (b0, b1) in zip(r, r[1:]):
                assert a1 == b0 or (a1 == nz and b0 == nz)
            assert all(a % 8 == 0 for a, b in r if b > a)


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["TSDF_ROOT"])
from oracle import oracle
from tsdf_b200 import scenes
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
n = (40, 40, 48)
ov = oracle.OracleVolume(n, (3000,) * 3)
oracle.set_threads(2)
for f in (0, 4, 9):
    cam = scenes.orbit_camera(f, 12)
    k = cam.k.copy(); k[:2] *= 0.25
    kinv = np.linalg.inv(k.astype(np.float64)).astype(np.float32)
    depth = scenes.render_depth(cam, 160, 120)
    # every rank fuses only its own slab (+ halo plane), like the GPU path
    z0, z1 = oracle.shard_ranges(n[2], 2)[rank]
    ov.integrate(depth, cam.inv_pose, k, kinv, z0, min(z1 + 1, n[2]))
keys = torch.from_numpy(oracle.raycast_slab_keys(ov, 160, 120, cam.pose, kinv, z0, z1))
dist.all_reduce(keys, op=dist.ReduceOp.MIN)
V, kh = oracle.resolve_keys(ov, keys.numpy(), 160, 120, cam.pose, kinv)
if rank == 0:
    whole = oracle.OracleVolume(n, (3000,) * 3)
    for f in (0, 4, 9):
        c2 = scenes.orbit_camera(f, 12)
        whole.integrate(scenes.render_depth(c2, 160, 120), c2.inv_pose, k, kinv)
    Vo, No, ko, so = whole.raycast(160, 120, cam.pose, kinv)
    same = ((V.view(np.uint32) == Vo.view(np.uint32)) | (np.isnan(V) & np.isnan(Vo))).all()
    assert same and np.array_equal(kh, ko) and (ko >= 0).sum() > 500, (same, (ko >= 0).sum())
    print("SHARDED_OK", int((ko >= 0).sum()))
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharded_raycast(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TSDF_ROOT=ROOT, PORT=str(29500 + os.getpid() % 2000), OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARDED_OK" in outs[0]


def test_interleaved_slabs_partition_the_volume():
    from tsdf_b200.sharded import interleaved_slabs
    for nz, world, slab in [(512, 8, 16), (112, 3, 8), (100, 4, 16), (16, 8, 16)]:
        seen = []
        for r in range(world):
            own = interleaved_slabs(nz, world, r, slab)
            assert all((z0 // slab) % world == r and z0 % slab == 0 and z1 - z0 <= slab for z0, z1 in own)
            seen += [z for z0, z1 in own for z in range(z0, z1)]
        assert sorted(seen) == list(range(nz))
