"""ctypes wrapper of oracle/liboracle.so — CPU ORACLE, TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs. The product never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, os.path.join(_REPO, "pointcloud-depthfusion_b200"))

from dfb_b200.ctypes_defs import CameraDesc, DfsCamera, DfsScene, Extrinsics, FuseDesc, Intrinsics  # noqa: E402

_LIBS = {}


def build():
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)


def load(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, name)
    sources = [os.path.join(_HERE, "dfb_oracle.c"), os.path.join(_HERE, "dfb_oracle.h"),
               os.path.join(_REPO, "include", "dfb", "dfb_synth.h"), os.path.join(_REPO, "include", "dfb", "dfb_types.h")]
    stale = not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in sources)
    if stale:  # the checker is rebuilt from source whenever it is older than its sources (gcc is in the image)
        build()
    L = C.CDLL(path)
    vp, i32, u32, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_float
    P = C.POINTER

    def proto(fname, restype, *argtypes):
        fn = getattr(L, fname)
        fn.restype = restype
        fn.argtypes = list(argtypes)

    proto("dfo_thresholds", None, f32, f32, f32, P(C.c_uint16), P(C.c_uint16))
    proto("dfo_resolve_roi", None, P(C.c_int32), i32, i32, P(C.c_int32))
    proto("dfo_filter_depth", u32, vp, vp, vp, i32, i32, f32, f32, f32, P(C.c_int32))
    proto("dfo_deproject_color", None, vp, vp, vp, P(Intrinsics), vp, u32, f32)
    proto("dfo_deproject", None, vp, vp, P(Intrinsics), vp, u32, f32)
    proto("dfo_transform", None, vp, u32, P(f32))
    proto("dfo_project_sequential", None, vp, vp, vp, u32, P(Intrinsics), i32)
    proto("dfo_project_index", C.c_int32, vp, P(Intrinsics), i32)
    proto("dfo_project_canonical", None, vp, vp, vp, vp, P(u32), i32, P(Intrinsics), i32)
    proto("dfo_fuse_frameset", C.c_int64, P(FuseDesc), P(vp), P(vp), vp, i32, vp, P(u32), vp)
    proto("dfo_fuse_batch", i32, P(FuseDesc), vp, vp, vp, i32, i32, i32)
    proto("dfo_registration_cloud", u32, P(CameraDesc), vp, vp)
    proto("dfo_align_depth_to_color", i32, vp, P(Intrinsics), P(Intrinsics), P(Extrinsics), f32, vp)
    proto("dfo_filter_color_gauss", None, vp, i32, i32)
    proto("dfo_filter_color_median", None, vp, i32, i32)
    proto("dfo_interpolate_transform", None, P(C.c_double), P(C.c_double), P(C.c_double))
    proto("dfo_prerotate_z90", None, P(C.c_double))
    proto("dfo_affine_to_float16", None, P(C.c_double), P(f32))
    proto("dfo_fused_intrinsics", None, P(Intrinsics), i32, P(Intrinsics))
    proto("dfo_voxel_approximate", u32, vp, u32, f32, vp)
    proto("dfo_voxel_exact", u32, vp, u32, f32, vp, vp)
    proto("dfo_synth_camera", None, P(DfsScene), P(DfsCamera), u32, vp, vp)
    proto("dfo_max_threads", i32)
    _LIBS[name] = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def thresholds(min_depth, max_depth, depth_scale):
    lt, gt = C.c_uint16(), C.c_uint16()
    load().dfo_thresholds(min_depth, max_depth, depth_scale, C.byref(lt), C.byref(gt))
    return lt.value, gt.value


def filter_depth(depth, width, height, min_depth, max_depth, depth_scale, roi=(-1, -1, -1, -1), mask=None):
    """Returns (filtered depth, mask, indices[:count], count). `mask` = previous mask state (stale outside ROI)."""
    d = np.ascontiguousarray(depth, dtype=np.uint16).reshape(-1).copy()
    m = np.zeros(width * height, dtype=np.uint8) if mask is None else np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1).copy()
    idx = np.zeros(width * height, dtype=np.uint32)
    r = (C.c_int32 * 4)(*roi)
    cnt = load().dfo_filter_depth(_ptr(d), _ptr(m), _ptr(idx), width, height, min_depth, max_depth, depth_scale, r)
    npos = int((m > 0).sum())
    return d.reshape(height, width), m.reshape(height, width), idx[:npos].copy(), int(cnt)


def deproject(depth, intr, indices, depth_scale, color=None):
    d = np.ascontiguousarray(depth, dtype=np.uint16).reshape(-1)
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    pts = np.zeros((len(idx), 4), dtype=np.float32)
    if color is not None:
        c = np.ascontiguousarray(color, dtype=np.uint8).reshape(-1)
        load().dfo_deproject_color(_ptr(pts), _ptr(d), _ptr(c), C.byref(intr), _ptr(idx), len(idx), depth_scale)
    else:
        load().dfo_deproject(_ptr(pts), _ptr(d), C.byref(intr), _ptr(idx), len(idx), depth_scale)
    return pts


def transform(points, T16):
    pts = np.ascontiguousarray(points, dtype=np.float32).copy()
    t = np.ascontiguousarray(np.asarray(T16, dtype=np.float32).reshape(16))
    load().dfo_transform(_ptr(pts), len(pts), t.ctypes.data_as(C.POINTER(C.c_float)))
    return pts


def project_sequential(points, intr, mirror):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    out = np.zeros((intr.height, intr.width, 3), dtype=np.uint8)
    zbuf = np.zeros((intr.height, intr.width), dtype=np.float32)
    load().dfo_project_sequential(_ptr(out), _ptr(zbuf), _ptr(pts), len(pts), C.byref(intr), int(mirror))
    return out, zbuf


def project_indices(points, intr, mirror):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    L = load()
    out = np.empty(len(pts), dtype=np.int32)
    base = pts.ctypes.data
    for i in range(len(pts)):
        out[i] = L.dfo_project_index(C.c_void_p(base + 16 * i), C.byref(intr), int(mirror))
    return out


def project_canonical(points, seg_counts, intr, mirror):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    offs = np.zeros(len(seg_counts) + 1, dtype=np.uint32)
    offs[1:] = np.cumsum(seg_counts)
    out = np.zeros((intr.height, intr.width, 3), dtype=np.uint8)
    keys = np.zeros((intr.height, intr.width), dtype=np.uint64)
    amb = np.zeros((intr.height, intr.width), dtype=np.uint8)
    load().dfo_project_canonical(_ptr(out), _ptr(keys), _ptr(amb), _ptr(pts), offs.ctypes.data_as(C.POINTER(C.c_uint32)),
                                 len(seg_counts), C.byref(intr), int(mirror))
    return out, keys, amb


def fuse_frameset(desc, depths, colors, mode=1, want_cloud=False, want_ambiguous=False, fast=False):
    """depths / colors: lists of per-camera arrays. Returns dict(image, total, counts[, cloud][, ambiguous])."""
    nc = desc.n_cameras
    ds = [np.ascontiguousarray(d, dtype=np.uint16) for d in depths]
    cs = [np.ascontiguousarray(c, dtype=np.uint8) for c in colors]
    dp = (C.c_void_p * nc)(*[d.ctypes.data for d in ds])
    cp = (C.c_void_p * nc)(*[c.ctypes.data for c in cs])
    ow, oh = desc.out_intr.width, desc.out_intr.height
    out = np.zeros((oh, ow, 3), dtype=np.uint8)
    cap = sum(d.size for d in ds)
    cloud = np.zeros((cap, 4), dtype=np.float32) if want_cloud else None
    counts = (C.c_uint32 * nc)()
    amb = np.zeros((oh, ow), dtype=np.uint8) if want_ambiguous else None
    total = load(fast).dfo_fuse_frameset(C.byref(desc), dp, cp, _ptr(out), int(mode),
                                         _ptr(cloud) if want_cloud else None, counts,
                                         _ptr(amb) if want_ambiguous else None)
    if total < 0:
        raise MemoryError("dfo_fuse_frameset failed")
    res = dict(image=out, total=int(total), counts=[int(x) for x in counts])
    if want_cloud:
        res["cloud"] = cloud[:total]
    if want_ambiguous:
        res["ambiguous"] = amb
    return res


def fuse_batch(desc, depth_pool, color_pool, n_framesets, mode=1, n_threads=None, fast=True):
    ow, oh = desc.out_intr.width, desc.out_intr.height
    out = np.zeros((n_framesets, oh, ow, 3), dtype=np.uint8)
    L = load(fast)
    if n_threads is None:
        n_threads = L.dfo_max_threads()
    rc = L.dfo_fuse_batch(C.byref(desc), _ptr(depth_pool), _ptr(color_pool), _ptr(out), n_framesets, int(mode), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"dfo_fuse_batch failed: {rc}")
    return out


def registration_cloud(cam_desc, depth):
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    pts = np.zeros((d.size, 4), dtype=np.float32)
    n = load().dfo_registration_cloud(C.byref(cam_desc), _ptr(d), _ptr(pts))
    return pts[:n].copy()


def align_depth_to_color(depth, depth_intr, color_intr, extrinsics, depth_scale):
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    out = np.zeros((color_intr.height, color_intr.width), dtype=np.uint16)
    rc = load().dfo_align_depth_to_color(_ptr(d), C.byref(depth_intr), C.byref(color_intr), C.byref(extrinsics), depth_scale, _ptr(out))
    if rc != 0:
        raise ValueError("alignment: unsupported distortion model or odd pixel count")
    return out


def filter_color(image, median=False):
    img = np.ascontiguousarray(image, dtype=np.uint8).copy()
    h, w = img.shape[:2]
    (load().dfo_filter_color_median if median else load().dfo_filter_color_gauss)(_ptr(img), w, h)
    return img


def interpolate_transform(left, right):
    la = np.ascontiguousarray(np.asarray(left, dtype=np.float64).reshape(16))
    ra = np.ascontiguousarray(np.asarray(right, dtype=np.float64).reshape(16))
    out = np.zeros(16, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    load().dfo_interpolate_transform(la.ctypes.data_as(dp), ra.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out.reshape(4, 4)


def prerotate_z90(T):
    a = np.ascontiguousarray(np.asarray(T, dtype=np.float64).reshape(16)).copy()
    load().dfo_prerotate_z90(a.ctypes.data_as(C.POINTER(C.c_double)))
    return a.reshape(4, 4)


def fused_intrinsics(color_left, vertical):
    out = Intrinsics()
    load().dfo_fused_intrinsics(C.byref(color_left), int(bool(vertical)), C.byref(out))
    return out


def voxel_approximate(points, leaf):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    out = np.zeros_like(pts)
    n = load().dfo_voxel_approximate(_ptr(pts), len(pts), leaf, _ptr(out))
    return out[:n].copy()


def voxel_exact(points, leaf):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    out = np.zeros_like(pts)
    ids = np.zeros((len(pts), 3), dtype=np.int32)
    n = load().dfo_voxel_exact(_ptr(pts), len(pts), leaf, _ptr(out), _ptr(ids))
    return out[:n].copy(), ids[:n].copy()


def synth_frameset(cfg, frameset_index):
    """CPU generator (include/dfb/dfb_synth.h) for all cameras of one frameset of a FusionConfig."""
    sc = cfg.scene(frameset_index)
    depths, colors = [], []
    for c, cam in enumerate(cfg.cams):
        d = np.zeros((cfg.height, cfg.width), dtype=np.uint16)
        col = np.zeros((cfg.height, cfg.width, 3), dtype=np.uint8)
        load(fast=True).dfo_synth_camera(C.byref(sc), C.byref(cam), c, _ptr(d), _ptr(col))
        depths.append(d)
        colors.append(col)
    return depths, colors


def synth_pool(cfg, first, count):
    depth = np.zeros((count, cfg.n_cameras, cfg.height, cfg.width), dtype=np.uint16)
    color = np.zeros((count, cfg.n_cameras, cfg.height, cfg.width, 3), dtype=np.uint8)
    for f in range(count):
        ds, cs = synth_frameset(cfg, first + f)
        for c in range(cfg.n_cameras):
            depth[f, c] = ds[c]
            color[f, c] = cs[c]
    return depth, color


def max_threads():
    """host threads the timed CPU arms use: the cores this process may run on (sched_getaffinity), NOT OpenMP's default —
    torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would silently pin the reference arm to one core.
    dfo_fuse_batch passes the count explicitly (num_threads clause), so the environment variable has no say."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)
