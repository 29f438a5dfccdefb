"""The BASELINE.json configurations as C-ABI descriptors + synthetic scenes.

Host-side matrix glue goes through the PRODUCT's own host functions (dfb_interpolate_transform,
dfb_prerotate_z90, dfb_fused_intrinsics — pure host math, no GPU needed), exactly the way
FusionNode builds them (fusion_node.cpp:155-165, 768-777).
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .ctypes_defs import CameraDesc, DfsCamera, DfsScene, FuseDesc, Intrinsics, MAX_CAMERAS

DEPTH_SCALE = 0.001  # D455 depth unit, fusion_node/config/config_default.yaml:15
MIN_DEPTH = 0.5      # config_default.yaml:21-22
MAX_DEPTH = 3.0
SEED_BASE = 0xD4550000

_MASK64 = (1 << 64) - 1


def _mix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _MASK64
    return x ^ (x >> 31)


def _unit(h):
    return (h & 0xFFFFFFFF) / 4294967296.0


def rot_x(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def rot_y(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def rot_z(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def affine(R, t):
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def integer_intrinsics(width, height):
    """Integer-valued intrinsics, as cameraInfo2Intrinsics truncates them (fusion_node.cpp:574-577)."""
    table = {(1280, 720): 631.0, (848, 480): 418.0, (3840, 2160): 1893.0}
    f = table.get((width, height), float(int(631.0 * width / 1280.0)))
    intr = Intrinsics()
    intr.width, intr.height = width, height
    intr.ppx, intr.ppy = float(width // 2), float(height // 2)
    intr.fx = intr.fy = f
    intr.model = 4  # DISTORTION_BROWN_CONRADY, fusion_node.cpp:578
    return intr


# right->left registration transforms. P0 "aligned", P1 "upside-down" (registration_node.cpp:154-160)
def transform_profile(profile, cam_index=1):
    if cam_index == 1:
        R = rot_x(2.0) @ rot_y(-8.0)
        t = np.array([0.35, 0.01, -0.02])
    elif cam_index == 2:
        R = rot_x(-1.5) @ rot_y(7.0)
        t = np.array([-0.33, -0.015, 0.01])
    else:
        R = rot_x(-6.0) @ rot_y(-1.0)
        t = np.array([0.02, -0.30, -0.01])
    if profile == "P1" and cam_index == 1:
        R = rot_z(180.0) @ R
    return affine(R, t)


@dataclass
class FusionConfig:
    name: str
    width: int
    height: int
    n_cameras: int
    profile: str
    density: str
    vertical: bool
    mirror: bool
    desc: FuseDesc
    cams: list
    cam_transforms: list  # 4x4 float64 per camera (identity for camera 0)
    fused_transform: np.ndarray
    hole_fraction: float = 0.0
    noise_lsb: int = 2
    extra: dict = field(default_factory=dict)

    @property
    def in_pixels(self):
        return self.width * self.height

    @property
    def out_width(self):
        return self.desc.out_intr.width

    @property
    def out_height(self):
        return self.desc.out_intr.height

    @property
    def out_pixels(self):
        return self.out_width * self.out_height

    @property
    def algorithmic_bytes(self):
        """every input pixel read once (2 B depth + 3 B colour), every output pixel written once."""
        return self.n_cameras * self.in_pixels * 5 + self.out_pixels * 3

    @property
    def mpix_in(self):
        return self.n_cameras * self.in_pixels / 1e6

    def scene(self, frameset_index):
        """Per-frameset scene (App. B.2): tilted back wall + sphere, sphere jittered per frameset."""
        seed = (SEED_BASE + int(frameset_index)) & _MASK64
        sc = DfsScene()
        wall_z = 2.4 if self.density == "dense" else 3.6
        n = rot_y(5.0) @ np.array([0.0, 0.0, 1.0])
        h1, h2, h3 = _mix64(seed ^ 0x11), _mix64(seed ^ 0x22), _mix64(seed ^ 0x33)
        c = (0.15 + 0.30 * (_unit(h1) - 0.5), 0.10 * (_unit(h2) - 0.5), 1.3 + 0.20 * (_unit(h3) - 0.5))
        for i in range(3):
            sc.wall_p0[i] = (0.0, 0.0, wall_z)[i]
            sc.wall_n[i] = float(n[i])
            sc.sphere_c[i] = c[i]
        sc.sphere_r = 0.35
        sc.inv_depth_unit = 1000.0
        sc.noise_lsb = self.noise_lsb
        sc.hole_q16 = int(round(self.hole_fraction * 65536.0))
        sc.seed = seed
        return sc

    def scenes(self, first, count):
        arr = (DfsScene * count)()
        for i in range(count):
            arr[i] = self.scene(first + i)
        return arr

    def cams_array(self):
        arr = (DfsCamera * self.n_cameras)()
        for i, c in enumerate(self.cams):
            arr[i] = c
        return arr


def make_config(width, height, n_cameras=2, profile="P0", density="dense", vertical=True, mirror=True,
                hole_fraction=0.0, noise_lsb=2, name=None, roi=None, min_depth=MIN_DEPTH, max_depth=MAX_DEPTH,
                depth_scale=DEPTH_SCALE):
    assert 1 <= n_cameras <= MAX_CAMERAS
    intr = integer_intrinsics(width, height)
    desc = FuseDesc()
    desc.n_cameras = n_cameras
    desc.mirror = int(mirror)
    cams, cam_T = [], []
    for c in range(n_cameras):
        T = np.eye(4) if c == 0 else transform_profile(profile, c)
        cam_T.append(T)
        cd = CameraDesc()
        cd.intr = intr
        cd.depth_scale = depth_scale
        cd.min_depth = min_depth
        cd.max_depth = max_depth
        r = roi[c] if roi is not None else ((0, 0, width, height) if c == 0 else (-1, -1, -1, -1))  # fusion_node.cpp:739-741
        for i in range(4):
            cd.roi[i] = r[i]
        cd.has_transform = int(c > 0)
        t16 = capi.affine_to_float16(T)
        for i in range(16):
            cd.to_ref[i] = float(t16[i])
        desc.cam[c] = cd
        dc = DfsCamera()
        dc.width, dc.height = width, height
        dc.ppx, dc.ppy, dc.fx, dc.fy = intr.ppx, intr.ppy, intr.fx, intr.fy
        for i in range(16):
            dc.cam_to_ref[i] = float(t16[i])
        cams.append(dc)
    # fused transform: interpolate(identity, T_right) then prerotate 90 deg about Z (fusion_node.cpp:768-777)
    right = cam_T[1] if n_cameras > 1 else np.eye(4)
    fused = capi.interpolate_transform(np.eye(4), right)
    if vertical:
        fused = capi.prerotate_z90(fused)
    f16 = capi.affine_to_float16(fused)
    for i in range(16):
        desc.fused_transform[i] = float(f16[i])
    desc.out_intr = capi.fused_intrinsics(intr, vertical)
    return FusionConfig(
        name=name or f"{n_cameras}x{width}x{height}-{profile}-{density}",
        width=width, height=height, n_cameras=n_cameras, profile=profile, density=density,
        vertical=vertical, mirror=mirror, desc=desc, cams=cams, cam_transforms=cam_T,
        fused_transform=fused, hole_fraction=hole_fraction, noise_lsb=noise_lsb,
    )


# BASELINE.json configs[0..4]
CONFIGS = {
    "C1": dict(width=848, height=480, n_cameras=2),
    "C2": dict(width=1280, height=720, n_cameras=2),
    "C3": dict(width=848, height=480, n_cameras=2),   # registration path: depth only
    "C4": dict(width=1280, height=720, n_cameras=4),
    "C5": dict(width=3840, height=2160, n_cameras=2),
}
