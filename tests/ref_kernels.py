"""ctypes wrapper of oracle/_ref/libref_kernels.so: the UNMODIFIED reference kernels.cu compiled in place
for sm_100a (oracle/Makefile `ref`) behind oracle/ref_driver.cu. TEST INFRASTRUCTURE ONLY; needs a GPU."""
import ctypes as C
import os

import numpy as np

from dfb_b200.ctypes_defs import Extrinsics, FuseDesc, Intrinsics

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATH = os.path.join(_REPO, "oracle", "_ref", "libref_kernels.so")
_LIB = None


def available():
    return os.path.exists(_PATH)


def load():
    global _LIB
    if _LIB is None:
        L = C.CDLL(_PATH)
        vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint32
        P = C.POINTER
        L.ref_init.restype = i32
        L.ref_init.argtypes = [i32, i32]
        L.ref_fuse_frameset.restype = i32
        L.ref_fuse_frameset.argtypes = [P(FuseDesc), P(vp), P(vp), vp, vp, P(u32), vp, P(vp), P(vp), P(vp), i32]
        L.ref_fuse_bench.restype = i32
        L.ref_fuse_bench.argtypes = [P(FuseDesc), vp, vp, vp, i32, i32, i32, P(C.c_float)]
        L.ref_transform.restype = i32
        L.ref_transform.argtypes = [vp, u32, P(C.c_float)]
        L.ref_project.restype = i32
        L.ref_project.argtypes = [vp, vp, vp, u32, P(Intrinsics), i32]
        L.ref_filter_color.restype = i32
        L.ref_filter_color.argtypes = [vp, i32, i32, i32]
        L.ref_align_depth.restype = i32
        L.ref_align_depth.argtypes = [vp, vp, P(Intrinsics), P(Intrinsics), P(Extrinsics), C.c_float]
        assert L.ref_init(0, 0) == 0
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fuse_frameset(desc, depths, colors, post_filter=0):
    """Runs the reference operator sequence on the GPU. Returns dict(image, cloud, counts, zbuf,
    depth_filtered[], mask[], indices[])."""
    nc = desc.n_cameras
    ds = [np.ascontiguousarray(d, dtype=np.uint16) for d in depths]
    cs = [np.ascontiguousarray(c, dtype=np.uint8) for c in colors]
    vp = C.c_void_p
    dp = (vp * nc)(*[d.ctypes.data for d in ds])
    cp = (vp * nc)(*[c.ctypes.data for c in cs])
    oh, ow = desc.out_intr.height, desc.out_intr.width
    img = np.zeros((oh, ow, 3), np.uint8)
    zbuf = np.zeros((oh, ow), np.float32)
    cap = sum(d.size for d in ds)
    cloud = np.zeros((cap, 4), np.float32)
    counts = (C.c_uint32 * nc)()
    dfs = [np.zeros_like(d) for d in ds]
    masks = [np.zeros(d.shape, np.uint8) for d in ds]
    idxs = [np.zeros(d.size, np.uint32) for d in ds]
    dfp = (vp * nc)(*[a.ctypes.data for a in dfs])
    mp = (vp * nc)(*[a.ctypes.data for a in masks])
    ip = (vp * nc)(*[a.ctypes.data for a in idxs])
    total = load().ref_fuse_frameset(C.byref(desc), dp, cp, _p(img), _p(cloud), counts, _p(zbuf), dfp, mp, ip, post_filter)
    if total < 0:
        raise RuntimeError("ref_fuse_frameset failed")
    cnts = [int(x) for x in counts]
    return dict(image=img, cloud=cloud[:total].copy(), counts=cnts, zbuf=zbuf, depth_filtered=dfs, mask=masks,
                indices=[i[:n].copy() for i, n in zip(idxs, cnts)], total=total)


def transform(points, T16):
    pts = np.ascontiguousarray(points, dtype=np.float32).copy()
    t = np.ascontiguousarray(np.asarray(T16, dtype=np.float32).reshape(16))
    assert load().ref_transform(_p(pts), len(pts), t.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return pts


def project(points, intr, mirror):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    img = np.zeros((intr.height, intr.width, 3), np.uint8)
    zbuf = np.zeros((intr.height, intr.width), np.float32)
    assert load().ref_project(_p(img), _p(zbuf), _p(pts), len(pts), C.byref(intr), int(mirror)) == 0
    return img, zbuf


def filter_color(image, median=False):
    img = np.ascontiguousarray(image, dtype=np.uint8).copy()
    h, w = img.shape[:2]
    assert load().ref_filter_color(_p(img), w, h, int(median)) == 0
    return img


def align_depth(depth, depth_intr, color_intr, extrinsics, depth_scale):
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    out = np.zeros((color_intr.height, color_intr.width), np.uint16)
    assert load().ref_align_depth(_p(out), _p(d), C.byref(depth_intr), C.byref(color_intr), C.byref(extrinsics), depth_scale) == 0
    return out


def fuse_bench(desc, d_depth, d_color, d_out, n_framesets, iters=3, post_filter=0):
    ms = C.c_float()
    rc = load().ref_fuse_bench(C.byref(desc), d_depth, d_color, d_out, n_framesets, iters, post_filter, C.byref(ms))
    if rc != 0:
        raise RuntimeError("ref_fuse_bench failed")
    return ms.value
