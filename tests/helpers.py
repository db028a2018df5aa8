"""Shared helpers for the parity tests: seeded scenes and bit-level comparisons."""
import numpy as np

from tsdf_b200 import scenes


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    """Bit-exact float comparison; NaNs compare equal to NaNs (payloads are not observable)."""
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    b = np.ascontiguousarray(b, np.float32).reshape(-1)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    both_nan = np.isnan(a) & np.isnan(b)
    bad = (bits(a) != bits(b)) & ~both_nan
    if bad.any():
        i = np.flatnonzero(bad)
        raise AssertionError(f"{what}: {i.size} of {a.size} values differ; first at {i[0]}: {a[i[0]]!r} vs {b[i[0]]!r}")


def random_rigid_pose(rng, centre=(1500.0, 1500.0, 1500.0), radius=(2500.0, 5000.0)):
    """A camera somewhere around the volume looking roughly at its centre (float32 4x4)."""
    cam = scenes.PinholeCamera()
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    r = rng.uniform(*radius)
    pos = np.asarray(centre) + r * d
    cam.move_to(*pos)
    tgt = np.asarray(centre) + rng.uniform(-300, 300, size=3)
    cam.look_at(*tgt)
    return cam


def random_depth(rng, w, h, lo=500, hi=6000, holes=0.1):
    d = rng.integers(lo, hi, size=(h, w)).astype(np.uint16)
    d[rng.random((h, w)) < holes] = 0
    return np.ascontiguousarray(d)
