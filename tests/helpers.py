"""Helpers shared by CPU and GPU tests."""
import ctypes as C

import numpy as np

from dfb_b200.ctypes_defs import CameraDesc, FuseDesc, Intrinsics


def make_intrinsics(width, height, fx, fy, ppx, ppy):
    i = Intrinsics()
    i.width, i.height = width, height
    i.fx, i.fy, i.ppx, i.ppy = fx, fy, ppx, ppy
    i.model = 4
    return i


def identity16():
    return np.eye(4, dtype=np.float32).reshape(16)


def simple_desc(width, height, n_cameras=1, mirror=False, fx=None, out=None, min_depth=0.5, max_depth=3.0,
                scale=0.001, fused=None, cam_T=None, roi=None, ppx=None, ppy=None):
    """Single-resolution FuseDesc with explicit transforms (float32 row-major 4x4)."""
    fx = fx if fx is not None else float(int(631.0 * width / 1280.0) or 1)
    intr = make_intrinsics(width, height, fx, fx, float(width // 2) if ppx is None else ppx,
                           float(height // 2) if ppy is None else ppy)
    d = FuseDesc()
    d.n_cameras = n_cameras
    d.mirror = int(mirror)
    for c in range(n_cameras):
        cd = CameraDesc()
        cd.intr = intr
        cd.depth_scale, cd.min_depth, cd.max_depth = scale, min_depth, max_depth
        r = roi[c] if roi is not None else (-1, -1, -1, -1)
        for i in range(4):
            cd.roi[i] = r[i]
        T = identity16() if (cam_T is None or cam_T[c] is None) else np.asarray(cam_T[c], dtype=np.float32).reshape(16)
        cd.has_transform = int(cam_T is not None and cam_T[c] is not None)
        for i in range(16):
            cd.to_ref[i] = float(T[i])
        d.cam[c] = cd
    F = identity16() if fused is None else np.asarray(fused, dtype=np.float32).reshape(16)
    for i in range(16):
        d.fused_transform[i] = float(F[i])
    d.out_intr = out if out is not None else intr
    return d


def fma32(a, b, c):
    """fp32 fused multiply-add emulated exactly through float64 (products of two fp32 are exact in fp64;
    the sum of that product and an fp32 is rounded once to fp64 and once to fp32 — double rounding is
    possible in theory, so use only for tolerance checks, never for bit-exact pins)."""
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def random_rigid(rng, max_deg=15.0, max_t=0.4):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = np.radians(rng.uniform(-max_deg, max_deg))
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rng.uniform(-max_t, max_t, size=3)
    return T


def candidate_counts(oracle, cloud, intr, mirror):
    """Number of points landing on each output pixel (vectorised restatement of dfo_project_index,
    checked against it in test_oracle_kats)."""
    idx = project_indices_np(cloud, intr, mirror)
    ok = idx >= 0
    return np.bincount(idx[ok], minlength=intr.width * intr.height).reshape(intr.height, intr.width), idx


def project_indices_np(cloud, intr, mirror):
    X, Y, Z = (cloud[:, i].astype(np.float32) for i in range(3))
    with np.errstate(all="ignore"):
        ix = np.float32(intr.ppx) + (np.float32(intr.fx) * X) / Z
        iy = np.float32(intr.ppy) + (np.float32(intr.fy) * Y) / Z
        tx = ix + np.float32(0.5)
        ty = iy + np.float32(0.5)
        ok = (Z > 0) & (tx > -1) & (tx < np.float32(intr.width)) & (ty > -1) & (ty < np.float32(intr.height))
        px = np.where(ok, tx, 0).astype(np.int32)
        py = np.where(ok, ty, 0).astype(np.int32)
    col = (intr.width - 1 - px) if mirror else px
    return np.where(ok, py * intr.width + col, -1).astype(np.int64)
