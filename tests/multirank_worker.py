"""Worker of tests/test_multirank_gpu.py: run under torch.distributed.run with one rank per GPU (NCCL).

Every rank fuses the same orbit frames into (a) a whole volume of its own (the single-GPU result) and (b) its shard of
a ShardedEngine spanning all ranks, for each multi-GPU layout, and compares — bit for bit — its slab's planes with the
whole volume's planes and the sharded raycast's vertex and normal maps with the single-GPU maps.  Exit code 0 = equal.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=60))
    from tsdf_b200 import scenes, sharded
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    layouts = sys.argv[2].split(",") if len(sys.argv) > 2 else ["contiguous", "peer", "interleaved", "replica"]
    n, phys, w, h = (size,) * 3, (3000.0,) * 3, 640, 480
    frames = [0, 90, 333, 610]
    cams = [scenes.orbit_camera(f, 1000) for f in frames]
    depths = [torch.from_numpy(scenes.render_depth(c, w, h)).cuda() for c in cams]
    whole = sharded.ShardedEngine(n, phys)
    want = []
    for cam, d in zip(cams, depths):
        whole.integrate(d, cam)
        whole.raycast(w, h, cam)
        torch.cuda.synchronize()
        want.append((whole.vertices.clone(), whole.normals.clone()))
    wd = whole.dist.view(size, -1).view(torch.int32)
    ww = whole.weight.view(size, -1).view(torch.int32)
    failures = []

    def same(a, b):
        a, b = a.view(torch.int32), b.view(torch.int32)
        nan = torch.isnan(a.view(torch.float32)) & torch.isnan(b.view(torch.float32))
        return bool(((a == b) | nan).all().item())

    for layout in layouts:
        # "peer": the contiguous layout with the key exchange fused into the march (atomics into rank 0's map over NVLink);
        # only rank 0 holds the result there
        exchange = "peer" if layout == "peer" else "allreduce"
        compare_here = layout != "peer" or rank == 0
        layout = "contiguous" if layout == "peer" else layout
        eng = sharded.ShardedEngine(n, phys, rank, world, layout=layout, slab=16, exchange=exchange)
        for i, (cam, d) in enumerate(zip(cams, depths)):
            eng.integrate(d, cam)
            eng.raycast(w, h, cam)
            torch.cuda.synchronize()
            dist.barrier()
            if compare_here and not same(eng.vertices, want[i][0]):
                failures.append(f"{layout}/{exchange}: vertices of frame {frames[i]} differ on rank {rank}")
            if compare_here and not same(eng.normals, want[i][1]):
                failures.append(f"{layout}/{exchange}: normals of frame {frames[i]} differ on rank {rank}")
        ld = eng.dist.view(-1, size * size).view(torch.int32)
        lw = eng.weight.view(-1, size * size).view(torch.int32)
        if layout == "contiguous":
            stored = eng.zs1 - eng.z0
            if not (torch.equal(ld[:stored], wd[eng.z0:eng.zs1]) and torch.equal(lw[:stored], ww[eng.z0:eng.zs1])):
                failures.append(f"{layout}: slab planes [{eng.z0}, {eng.zs1}) differ on rank {rank}")
        else:
            for j, (z0, z1) in enumerate(eng.slabs):
                st = (z1 - z0) + (1 if z1 < size else 0)
                if not (torch.equal(ld[j * (eng.slab + 1): j * (eng.slab + 1) + st], wd[z0:z0 + st]) and
                        torch.equal(lw[j * (eng.slab + 1): j * (eng.slab + 1) + st], ww[z0:z0 + st])):
                    failures.append(f"{layout}: slab [{z0}, {z1}) differs on rank {rank}")
        # clear + one more frame: the replica must not keep pre-clear voxels
        eng.clear()
        whole2 = sharded.ShardedEngine(n, phys)
        whole2.integrate(depths[2], cams[2]); whole2.raycast(w, h, cams[2])
        eng.integrate(depths[2], cams[2]); eng.raycast(w, h, cams[2])
        torch.cuda.synchronize()
        dist.barrier()
        if compare_here and not same(eng.vertices, whole2.vertices):
            failures.append(f"{layout}/{exchange}: vertices after clear differ on rank {rank}")
        whole2.close()
        torch.cuda.synchronize()
        dist.barrier()
        eng.close()
        dist.barrier()
    hits = int((~torch.isnan(want[-1][0].view(-1, 3)[:, 0])).sum().item())
    if hits < 20000:
        failures.append(f"only {hits} ray hits: the comparison is vacuous")
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for f in failures:
        print("FAIL", f, flush=True)
    if rank == 0:
        print(f"multirank parity: world {world}, size {size}, layouts {layouts}: {int(flag.item())} failures, {hits} hits", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
