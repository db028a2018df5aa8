#!/usr/bin/env python
"""Write the synthetic sphere + wall scene as a TUM RGB-D style directory, so that the reference's own driver
(`kinfu -m N -d DIR`, src/Tools/kinfu.cpp) runs unchanged on it:

    DIR/ground_truth.txt      "timestamp tx ty tz qx qy qz qw"   (metres, unit quaternion; TUMDataLoader.cpp:111-128)
    DIR/depth/<timestamp>.png 16-bit greyscale, 5000 units per metre (scaled x0.2 to mm by TUMDataLoader.cpp:96-100)
"""
import argparse
import os
import struct
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsdf_b200 import scenes  # noqa: E402


def write_png16(path, img):
    h, w = img.shape
    raw = b"".join(b"\x00" + img[y].astype(">u2").tobytes() for y in range(h))

    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 1)) + chunk(b"IEND", b""))


def quaternion(rot):
    """Unit quaternion (x, y, z, w) of a rotation matrix (Shepperd's method)."""
    m = np.asarray(rot, np.float64)
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        return (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s
    i = int(np.argmax(np.diag(m)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(1.0 + m[i, i] - m[j, j] - m[k, k]) * 2
    q = [0.0, 0.0, 0.0, 0.0]
    q[i] = 0.25 * s
    q[j] = (m[j, i] + m[i, j]) / s
    q[k] = (m[k, i] + m[i, k]) / s
    q[3] = (m[k, j] - m[j, k]) / s
    return tuple(q)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("directory")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--orbit", type=int, default=40, help="frames in a full orbit")
    args = ap.parse_args()
    os.makedirs(os.path.join(args.directory, "depth"), exist_ok=True)
    with open(os.path.join(args.directory, "ground_truth.txt"), "w") as gt:
        gt.write("# timestamp tx ty tz qx qy qz qw\n")
        for i in range(args.frames):
            cam = scenes.orbit_camera(i, args.orbit)
            depth_mm = scenes.render_depth(cam)
            stamp = f"{1000 + i}.000000"
            # multiples of 5 survive the x0.2 (float) scaling exactly
            write_png16(os.path.join(args.directory, "depth", stamp + ".png"), depth_mm.astype(np.uint32) * 5 % 65536)
            t = cam.pose[:3, 3].astype(np.float64) / 1000.0
            q = quaternion(cam.pose[:3, :3])
            gt.write(f"{stamp} {t[0]:.9f} {t[1]:.9f} {t[2]:.9f} {q[0]:.9f} {q[1]:.9f} {q[2]:.9f} {q[3]:.9f}\n")
    print(f"wrote {args.frames} frames to {args.directory}")


if __name__ == "__main__":
    main()
