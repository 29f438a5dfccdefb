"""Loader + prototypes for libdfb_b200.so. Pointers are plain ints (e.g. torch.Tensor.data_ptr())."""
import ctypes as C
import os

import numpy as np

from .ctypes_defs import CameraDesc, DfsCamera, DfsScene, Extrinsics, FuseDesc, ImageMsg, Intrinsics

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class DfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdfb_b200 error {code}: {msg}")
        self.code = code


def lib_path():
    """the in-tree library; DFB_B200_LIB selects another build of the SAME library (kernel experiments, tools/variants.sh)"""
    return os.environ.get("DFB_B200_LIB") or os.path.join(_PKG_DIR, "libdfb_b200.so")


def load():
    """dlopen the in-tree library; fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: build it with `make -C pointcloud-depthfusion_b200` "
            "(or __graft_entry__.build()). There is no CPU fallback."
        )
    L = C.CDLL(path)
    vp, u64, i32, f32 = C.c_void_p, C.c_uint64, C.c_int, C.c_float
    P = C.POINTER

    experimental = bool(os.environ.get("DFB_B200_LIB"))

    def proto(name, restype, *argtypes):
        if experimental and not hasattr(L, name):
            return  # an older experiment build of the library: entry points added since are simply absent
        fn = getattr(L, name)
        fn.restype = restype
        fn.argtypes = list(argtypes)

    proto("dfb_init", i32, i32, i32)
    proto("dfb_last_error", C.c_char_p)
    proto("dfb_version", i32)
    proto("dfb_launch_count", u64)
    proto("dfb_host_alloc", i32, P(vp), C.c_size_t)
    proto("dfb_host_free", i32, vp)
    proto("dfb_device_alloc", i32, P(vp), C.c_size_t)
    proto("dfb_device_free", i32, vp)
    proto("dfb_ipc_export", i32, vp, C.c_char_p)
    proto("dfb_ipc_open", i32, C.c_char_p, P(vp))
    proto("dfb_ipc_close", i32, vp)
    proto("dfb_fuser_fuse_images", i32, vp, P(vp), P(vp), vp, vp)
    proto("dfb_interpolate_transform", None, P(C.c_double), P(C.c_double), P(C.c_double))
    proto("dfb_prerotate_z90", None, P(C.c_double))
    proto("dfb_affine_to_float16", None, P(C.c_double), P(f32))
    proto("dfb_fused_intrinsics", None, P(Intrinsics), i32, P(Intrinsics))
    proto("dfb_camera_info_to_intrinsics", None, C.c_uint32, C.c_uint32, P(C.c_double), P(C.c_double), P(Intrinsics))
    proto("dfb_virtual_camera_transform", None, P(C.c_double), P(C.c_double), P(C.c_double))
    proto("dfb_wire_unpack_depth", i32, P(ImageMsg), i32, i32, vp)
    proto("dfb_wire_unpack_color", i32, P(ImageMsg), i32, i32, vp)
    proto("dfb_wire_fused_image", i32, P(Intrinsics), vp, P(ImageMsg))
    proto("dfb_depth_thresholds", None, f32, f32, f32, P(C.c_uint16), P(C.c_uint16))
    proto("dfb_fuser_create", i32, P(vp), P(FuseDesc))
    proto("dfb_fuser_destroy", i32, vp)
    proto("dfb_fuser_update", i32, vp, P(FuseDesc))
    proto("dfb_fuser_set_group", i32, vp, i32)
    proto("dfb_fuser_get_group", i32, vp)
    proto("dfb_fuser_set_variant", i32, vp, i32)
    proto("dfb_fuser_set_post_filter", i32, vp, i32)
    proto("dfb_fuser_fuse_device", i32, vp, vp, vp, vp, i32, vp)
    proto("dfb_fuser_fuse_host", i32, vp, vp, vp, vp, i32)
    proto("dfb_fuser_profile_device", i32, vp, vp, vp, vp, i32, vp, P(f32), P(i32))
    proto("dfb_workspace_create", i32, P(vp))
    proto("dfb_workspace_destroy", i32, vp)
    proto("dfb_op_filter_depth", i32, vp, vp, vp, vp, i32, i32, f32, f32, f32, P(C.c_int32), P(C.c_uint32), vp)
    proto("dfb_op_filter_depth_begin", i32, vp, vp, vp, vp, i32, i32, f32, f32, f32, P(C.c_int32), vp)
    proto("dfb_op_filter_depth_count", i32, vp, P(C.c_uint32))
    proto("dfb_op_filter_depth_indices", i32, vp, vp, vp, i32, i32, vp)
    proto("dfb_op_deproject", i32, vp, vp, vp, P(Intrinsics), vp, C.c_uint32, f32, vp)
    proto("dfb_op_transform", i32, vp, C.c_uint32, P(f32), vp)
    proto("dfb_op_project", i32, vp, vp, vp, P(C.c_uint32), i32, P(Intrinsics), i32, vp)
    proto("dfb_op_keys_clear", i32, vp, C.c_size_t, vp)
    proto("dfb_op_filter_color", i32, vp, vp, i32, i32, i32, vp)
    proto("dfb_registration_cloud", i32, vp, P(CameraDesc), vp, vp, P(C.c_uint32), vp)
    proto("dfb_op_align_depth_to_color", i32, vp, vp, P(Intrinsics), P(Intrinsics), P(Extrinsics), f32, vp)
    proto("dfb_voxel_downsample", i32, vp, vp, C.c_uint32, f32, vp, P(C.c_uint32), vp)
    proto("dfb_voxel_downsample_approximate", i32, vp, vp, C.c_uint32, f32, vp, P(C.c_uint32), vp)
    proto("dfb_synth_generate", i32, P(DfsScene), i32, P(DfsCamera), i32, vp, vp, vp)
    proto("dfb_selftest_division", C.c_int64, u64, u64)
    proto("dfb_probe_atomics", f32, u64, u64, i32, C.c_uint32, i32)
    proto("dfb_selftest_exclusive_scan", i32, vp, C.c_uint32, C.POINTER(C.c_uint32), vp)
    _LIB = L
    return L


def lib():
    return load()


def check(rc):
    if rc != 0:
        raise DfbError(rc, load().dfb_last_error().decode(errors="replace"))
    return rc


def last_error():
    return (lib().dfb_last_error() or b"").decode()


def init(device=0, verbose=False):
    check(load().dfb_init(int(device), int(bool(verbose))))


def launch_count():
    return int(load().dfb_launch_count())


def _f64x16(a):
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(16))
    return arr, arr.ctypes.data_as(C.POINTER(C.c_double))


def interpolate_transform(left, right):
    """FusionNode::interpolateTransform (fusion_node.cpp:589-604); 4x4 row-major doubles."""
    la, lp = _f64x16(left)
    ra, rp = _f64x16(right)
    out = np.zeros(16, dtype=np.float64)
    load().dfb_interpolate_transform(lp, rp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out.reshape(4, 4)


def prerotate_z90(T):
    a, p = _f64x16(T)
    a = a.copy()
    load().dfb_prerotate_z90(a.ctypes.data_as(C.POINTER(C.c_double)))
    return a.reshape(4, 4)


def virtual_camera_transform(translation, rotation_deg):
    t = np.ascontiguousarray(np.asarray(translation, dtype=np.float64).reshape(3))
    r = np.ascontiguousarray(np.asarray(rotation_deg, dtype=np.float64).reshape(3))
    out = np.zeros(16, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    load().dfb_virtual_camera_transform(t.ctypes.data_as(dp), r.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out.reshape(4, 4)


def affine_to_float16(T):
    a, p = _f64x16(T)
    out = np.zeros(16, dtype=np.float32)
    load().dfb_affine_to_float16(p, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def fused_intrinsics(color_left, vertical):
    out = Intrinsics()
    load().dfb_fused_intrinsics(C.byref(color_left), int(bool(vertical)), C.byref(out))
    return out


def camera_info_to_intrinsics(width, height, K, D):
    k = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    d = np.ascontiguousarray(np.asarray(D, dtype=np.float64).reshape(5))
    out = Intrinsics()
    dp = C.POINTER(C.c_double)
    load().dfb_camera_info_to_intrinsics(int(width), int(height), k.ctypes.data_as(dp), d.ctypes.data_as(dp), C.byref(out))
    return out


def depth_thresholds(min_depth, max_depth, depth_scale):
    lt, gt = C.c_uint16(), C.c_uint16()
    load().dfb_depth_thresholds(min_depth, max_depth, depth_scale, C.byref(lt), C.byref(gt))
    return lt.value, gt.value


class Fuser:
    """dfb_fuser: the batched fused path (one FusionNode::processSyncedFrames per frameset)."""

    def __init__(self, desc: FuseDesc):
        self._h = C.c_void_p()
        self.desc = desc
        check(load().dfb_fuser_create(C.byref(self._h), C.byref(desc)))

    def close(self):
        if self._h:
            load().dfb_fuser_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, desc: FuseDesc):
        check(load().dfb_fuser_update(self._h, C.byref(desc)))
        self.desc = desc

    def set_group(self, g):
        check(load().dfb_fuser_set_group(self._h, int(g)))

    @property
    def group(self):
        return load().dfb_fuser_get_group(self._h)

    def set_post_filter(self, mode):
        check(load().dfb_fuser_set_post_filter(self._h, int(mode)))

    def set_variant(self, v):
        check(load().dfb_fuser_set_variant(self._h, int(v)))

    def fuse_device(self, d_depth, d_color, d_out, n_framesets, stream=0):
        check(load().dfb_fuser_fuse_device(self._h, d_depth, d_color, d_out, int(n_framesets), stream))

    def profile_device(self, d_depth, d_color, d_out, n_framesets, stream=0):
        """Returns dict(splat_ms, resolve_ms, splat_launches, resolve_launches) summed over the call."""
        ms = (C.c_float * 2)()
        n = (C.c_int * 2)()
        check(load().dfb_fuser_profile_device(self._h, d_depth, d_color, d_out, int(n_framesets), stream, ms, n))
        return dict(splat_ms=ms[0], resolve_ms=ms[1], splat_launches=n[0], resolve_launches=n[1])

    def fuse_host(self, h_depth, h_color, h_out, n_framesets):
        check(load().dfb_fuser_fuse_host(self._h, h_depth, h_color, h_out, int(n_framesets)))


class Workspace:
    def __init__(self):
        self._h = C.c_void_p()
        check(load().dfb_workspace_create(C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            load().dfb_workspace_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
