"""ctypes mirrors of include/dfb/dfb_types.h and include/dfb/dfb_synth.h (layout is ABI)."""
import ctypes as C

MAX_CAMERAS = 8


class Intrinsics(C.Structure):
    """struct Intrinsics, pointcloud_processing/include/pointcloud_processing/intrinsics.h:20-30 (48 bytes)."""

    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("ppx", C.c_float),
        ("ppy", C.c_float),
        ("fx", C.c_float),
        ("fy", C.c_float),
        ("model", C.c_int32),
        ("coeffs", C.c_float * 5),
    ]


class Extrinsics(C.Structure):
    """struct Extrinsics, intrinsics.h:35-39."""

    _fields_ = [("rotation", C.c_float * 9), ("translation", C.c_float * 3)]


class CameraDesc(C.Structure):
    _fields_ = [
        ("intr", Intrinsics),
        ("depth_scale", C.c_float),
        ("min_depth", C.c_float),
        ("max_depth", C.c_float),
        ("roi", C.c_int32 * 4),
        ("has_transform", C.c_int32),
        ("to_ref", C.c_float * 16),
    ]


class FuseDesc(C.Structure):
    _fields_ = [
        ("n_cameras", C.c_int32),
        ("mirror", C.c_int32),
        ("cam", CameraDesc * MAX_CAMERAS),
        ("fused_transform", C.c_float * 16),
        ("out_intr", Intrinsics),
    ]


class DfsCamera(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("ppx", C.c_float),
        ("ppy", C.c_float),
        ("fx", C.c_float),
        ("fy", C.c_float),
        ("cam_to_ref", C.c_float * 16),
    ]


class DfsScene(C.Structure):
    _fields_ = [
        ("wall_p0", C.c_double * 3),
        ("wall_n", C.c_double * 3),
        ("sphere_c", C.c_double * 3),
        ("sphere_r", C.c_double),
        ("inv_depth_unit", C.c_double),
        ("noise_lsb", C.c_int32),
        ("hole_q16", C.c_uint32),
        ("seed", C.c_uint64),
    ]


assert C.sizeof(Intrinsics) == 48
assert C.sizeof(Extrinsics) == 48


class ImageMsg(C.Structure):
    """dfb_image_msg: the sensor_msgs/Image fields of a DepthFrameset member (include/dfb/dfb_types.h)"""
    _fields_ = [("height", C.c_uint32), ("width", C.c_uint32), ("encoding", C.c_char_p), ("is_bigendian", C.c_uint8),
                ("step", C.c_uint32), ("data", C.c_void_p), ("data_size", C.c_size_t)]
