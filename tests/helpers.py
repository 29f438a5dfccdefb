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


def align_case(rng, depth_model=4, color_model=2, coeff_scale=0.0, dw=640, dh=480, cw=848, ch=480):
    """A synthetic unaligned-depth case: depth camera (dw x dh) next to a colour camera (cw x ch), a tilted plane with
    holes as the depth image, small rotation + 15 mm baseline as depth->colour extrinsics. Returns
    (depth u16[dh][dw], depth_intr, color_intr, Extrinsics, depth_scale)."""
    from dfb_b200.ctypes_defs import Extrinsics

    di = make_intrinsics(dw, dh, 0.6 * dw, 0.6 * dw, dw / 2 - 0.5, dh / 2 + 0.25)
    ci = make_intrinsics(cw, ch, 0.75 * cw, 0.75 * cw, cw / 2 + 3.5, ch / 2 - 1.75)
    di.model, ci.model = depth_model, color_model
    for i in range(5):
        di.coeffs[i] = float(rng.uniform(-1, 1) * coeff_scale)
        ci.coeffs[i] = float(rng.uniform(-1, 1) * coeff_scale)
    T = random_rigid(rng, 1.5, 0.0)
    ex = Extrinsics()
    for r in range(3):
        for c in range(3):
            ex.rotation[c * 3 + r] = float(T[r, c])  # column-major, like rs2_extrinsics
    ex.translation[0], ex.translation[1], ex.translation[2] = 0.015, -0.0005, 0.0003
    v, u = np.mgrid[0:dh, 0:dw]
    z = 1.2 + 0.0009 * u + 0.0004 * v + 0.05 * np.sin(u / 37.0)
    depth = np.clip(z / 0.001 + rng.integers(-2, 3, size=z.shape), 0, 65535).astype(np.uint16)
    depth[rng.random(z.shape) < 0.1] = 0
    depth[:, : dw // 10][rng.random((dh, dw // 10)) < 0.5] = 300  # a near object: occlusion / min test
    return depth, di, ci, ex, np.float32(0.001)
