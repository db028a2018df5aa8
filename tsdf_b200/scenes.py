"""Deterministic synthetic inputs for tests and bench.

Scene: a sphere (centre (1500,1500,1500), r = 800 mm) in front of a back wall z = 2600 mm inside
the reference's default 3000 mm cube — the analytic counterparts of make_sphere_depth_map /
make_wall_depth_map (reference src/Tests/TestTSDF/TestHelpers.cpp:145-209).  Depth frames are
rendered in float64 and stored as uint16 millimetres of camera-space z, 0 = no measurement,
exactly the format TSDFVolume::integrate consumes (src/TSDF/TSDFVolume.cu:861).

The pin-hole camera below mirrors the host maths of the reference Camera (src/Camera.cpp:20-191)
in float32 numpy; it only produces kernel INPUTS (the same matrices go to the oracle and to the
CUDA path), the drop-in C++ Camera lives in tsdf_b200/include/Camera.hpp.
"""
import numpy as np

F32 = np.float32
DEFAULT_INTRINSICS = (591.1, 590.1, 331.0, 234.6)   # Camera::default_depth_camera(), Camera.hpp:41-44
CENTRE = np.array([1500.0, 1500.0, 1500.0])
SPHERE_R = 800.0
WALL_Z = 2600.0


class PinholeCamera:
    def __init__(self, fx=DEFAULT_INTRINSICS[0], fy=DEFAULT_INTRINSICS[1], cx=DEFAULT_INTRINSICS[2], cy=DEFAULT_INTRINSICS[3]):
        fx, fy, cx, cy = F32(fx), F32(fy), F32(cx), F32(cy)
        self.k = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], F32)
        one = F32(1)
        self.kinv = np.array([[one / fx, 0, -cx / fx], [0, one / fy, -cy / fy], [0, 0, 1]], F32)
        self.set_pose(np.eye(4, dtype=F32))

    def set_pose(self, pose):
        self.pose = np.array(pose, F32)
        # rigid poses are inverted in float64 and rounded once; anything else goes through LAPACK
        self.inv_pose = np.linalg.inv(self.pose.astype(np.float64)).astype(F32)

    def move_to(self, x, y, z):                      # Camera.cpp:128-134
        p = self.pose.copy()
        p[0, 3], p[1, 3], p[2, 3] = x, y, z
        self.set_pose(p)

    def look_at(self, x, y, z):                      # Camera.cpp:142-191 (gluLookAt with +Y up)
        pos = self.pose[:3, 3].astype(F32)
        fwd = (np.array([x, y, z], F32) - pos).astype(F32)
        fwd = (fwd / F32(np.sqrt(F32(np.dot(fwd, fwd))))).astype(F32)
        eps = 1e-6
        if abs(fwd[0]) < eps and abs(fwd[2]) < eps:
            up = np.array([0, 0, 1 if fwd[1] < 0 else -1], F32)
        else:
            up = np.array([0, 1, 0], F32)
        left = np.cross(up, fwd).astype(F32)
        left = (left / F32(np.sqrt(F32(np.dot(left, left))))).astype(F32)
        up = np.cross(fwd, left).astype(F32)
        up = (up / F32(np.sqrt(F32(np.dot(up, up))))).astype(F32)
        p = self.pose.copy()
        p[:3, 0], p[:3, 1], p[:3, 2] = left, up, fwd
        p[3, :] = (0, 0, 0, 1)
        self.set_pose(p)

    @property
    def position(self):
        return self.pose[:3, 3].copy()

    @property
    def rot(self):
        return self.pose[:3, :3].copy()


def fixed_pose_camera():
    """Config 1: identity rotation at (1500,1500,-2500), looking down +z at the volume's front face."""
    cam = PinholeCamera()
    cam.move_to(1500.0, 1500.0, -2500.0)
    return cam


def orbit_camera(i, n_frames, radius=4000.0):
    """Configs 2-5: orbit in the XZ plane about the volume centre, theta_i = 2*pi*i/n."""
    th = 2.0 * np.pi * i / n_frames
    cam = PinholeCamera()
    cam.move_to(CENTRE[0] + radius * np.sin(th), CENTRE[1], CENTRE[2] - radius * np.cos(th))
    cam.look_at(*CENTRE)
    return cam


def xorshift32(seed, n):
    """n values of Marsaglia's xorshift32 stream (13, 17, 5) seeded with `seed`."""
    out = np.empty(n, np.uint32)
    x = np.uint32(seed)
    # vectorised over independent lanes would change the stream; n is small (one frame) so loop in chunks
    x = int(x)
    for i in range(n):
        x ^= (x << 13) & 0xFFFFFFFF
        x ^= x >> 17
        x ^= (x << 5) & 0xFFFFFFFF
        out[i] = x
    return out


def render_depth(cam, width=640, height=480, sphere=True, wall=True, noise_seed=None):
    """Analytic depth (uint16 mm, camera-space z) of the sphere + wall scene seen from `cam`."""
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    kinv = cam.kinv.astype(np.float64)
    dc = np.stack([kinv[0, 0] * u + kinv[0, 1] * v + kinv[0, 2],
                   kinv[1, 0] * u + kinv[1, 1] * v + kinv[1, 2],
                   np.ones_like(u)], axis=-1)                       # camera-space ray, z = 1
    R = cam.pose[:3, :3].astype(np.float64)
    o = cam.pose[:3, 3].astype(np.float64)
    d = dc @ R.T                                                    # world-space ray
    best = np.full(u.shape, np.inf)
    if sphere:
        oc = o - CENTRE
        a = np.sum(d * d, axis=-1)
        b = 2.0 * (d @ oc)
        c = float(oc @ oc) - SPHERE_R ** 2
        disc = b * b - 4 * a * c
        ok = disc >= 0
        s = np.where(ok, (-b - np.sqrt(np.where(ok, disc, 0.0))) / (2 * a), np.inf)
        s = np.where(s > 0, s, np.inf)
        best = np.minimum(best, s)
    if wall:
        with np.errstate(divide="ignore", invalid="ignore"):
            s = (WALL_Z - o[2]) / d[..., 2]
        s = np.where(np.isfinite(s) & (s > 0), s, np.inf)
        best = np.minimum(best, s)
    depth = np.where(np.isfinite(best) & (best < 65535.0), np.rint(best), 0.0)
    if noise_seed is not None:
        r = xorshift32(noise_seed, width * height).reshape(height, width)
        depth = np.where(depth > 0, np.clip(depth + (r % 5).astype(np.float64) - 2.0, 1, 65535), 0.0)
    return np.ascontiguousarray(depth.astype(np.uint16))
