import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle, refcuda
from tsdf_b200 import scenes
lib = refcuda.RefLib("O3")
n = (64, 64, 64); w, h = 320, 240; s = 0.5
rv = refcuda.RefVolume(lib, n, (3000, 3000, 3000))
ov = oracle.OracleVolume(n, (3000, 3000, 3000))
for f in (0, 3, 7):
    cam = scenes.orbit_camera(f, 12)
    k = cam.k.copy(); k[:2] *= s
    kinv, inv_pose = lib.camera_matrices(k, cam.pose)
    depth = scenes.render_depth(cam, w, h)
    rv.integrate(depth, k, cam.pose); ov.integrate(depth, inv_pose, k, kinv)
Vr, Nr = rv.raycast(w, h, k, cam.pose)
Vo, No, ko, so = ov.raycast(w, h, cam.pose, kinv)
bad = np.flatnonzero((Vr.view(np.uint32) != Vo.view(np.uint32)).any(axis=1) & ~(np.isnan(Vr).all(axis=1) & np.isnan(Vo).all(axis=1)))
print("mismatching pixels", bad.size, "of", w * h, "oracle hits", (ko >= 0).sum(), "ref hits", (~np.isnan(Vr[:, 0])).sum())
ys, xs = bad // w, bad % w
print("x range", xs.min(), xs.max(), "y range", ys.min(), ys.max())
print("khit of oracle at mismatches: min", ko[bad].min(), "max", ko[bad].max(), "hist", np.histogram(ko[bad], bins=8)[0])
print("ref NaN at mismatches:", np.isnan(Vr[bad, 0]).sum(), " oracle NaN at mismatches:", np.isnan(Vo[bad, 0]).sum())
for i in bad[:8]:
    print(i % w, i // w, "ref", Vr[i], "oracle", Vo[i], "k", ko[i])
# where both hit: max abs diff
both = ~np.isnan(Vr[:, 0]) & ~np.isnan(Vo[:, 0])
print("both hit:", both.sum(), "max |diff|", np.abs(Vr[both] - Vo[both]).max() if both.any() else None)
img = np.zeros((h, w), np.uint8); img[ys, xs] = 1
for r in range(0, h, 12):
    print("".join("#" if img[r:r+12, c:c+8].any() else "." for c in range(0, w, 8)))
