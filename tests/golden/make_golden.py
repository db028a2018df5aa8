"""Generates tests/golden/t_100_2000_50.json from the reference's only binary fixture,
/root/reference/TestData/t_100_2000_50.tsdf (35 MB: a freshly cleared 100^3 / 2000 mm volume saved
by the reference's own TSDFVolume::save_to_file).  Run in the build container (the reference tree is
not present on the GPU box); the JSON it writes is what the tests read.

Layout (reference src/TSDF/TSDFVolume.cu:995-1013): 68-byte header, float dist[N], float weight[N],
uchar3 colour[N], DeformationNode{float3 translation; float3 rotation}[N].
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/TestData/t_100_2000_50.tsdf"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "t_100_2000_50.json")

raw = open(SRC, "rb").read()
hdr = raw[:68]
nx, ny, nz = struct.unpack_from("<3I", hdr, 0)
phys = struct.unpack_from("<3f", hdr, 12)
off = struct.unpack_from("<3f", hdr, 24)
trunc, maxw = struct.unpack_from("<2f", hdr, 36)
gt = struct.unpack_from("<3f", hdr, 44)
gr = struct.unpack_from("<3f", hdr, 56)
n = nx * ny * nz
assert len(raw) == 68 + n * (4 + 4 + 3 + 24), len(raw)
o = 68
dist = np.frombuffer(raw, "<f4", n, o); o += 4 * n
weight = np.frombuffer(raw, "<f4", n, o); o += 4 * n
colour = np.frombuffer(raw, "u1", 3 * n, o); o += 3 * n
deform = np.frombuffer(raw, "<f4", 6 * n, o)

sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
sample_idx = [0, 1, 99, 100, 9999, 10000, 123456, 500000, n - 1]
golden = {
    "source": "TestData/t_100_2000_50.tsdf",
    "file_bytes": len(raw),
    "header_hex": hdr.hex(),
    "size": [nx, ny, nz], "physical": phys, "offset": off,
    "trunc_bits": struct.unpack("<I", struct.pack("<f", trunc))[0], "trunc": trunc,
    "max_weight": maxw, "global_translation": gt, "global_rotation": gr,
    "dist_sha256": sha(dist), "weight_sha256": sha(weight), "deform_sha256": sha(deform),
    "dist_unique_bits": sorted(set(int(x) for x in np.unique(dist.view("<u4")))),
    "weight_unique_bits": sorted(set(int(x) for x in np.unique(weight.view("<u4")))),
    "colour_nonzero": int(np.count_nonzero(colour)),
    "deform_samples": {str(i): [float(x) for x in deform[6 * i:6 * i + 6]] for i in sample_idx},
}
json.dump(golden, open(OUT, "w"), indent=1)
print("wrote", OUT)
