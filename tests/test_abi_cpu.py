"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/dfb/dfb.h
declares, the PODs have the reference's layout, the host-side glue (pure host math) agrees with the
oracle's restatement, and compute entry points fail LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import random_rigid


def declared_symbols(repo_root):
    text = open(os.path.join(repo_root, "include", "dfb", "dfb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"DFB_API\s+[\w\s\*]+?\b(dfb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(dfb, repo_root):
    syms = declared_symbols(repo_root)
    assert len(syms) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", dfb.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dfb_\w+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f"declared in dfb.h but not exported: {missing}"
    L = dfb.lib()
    for s in syms:
        assert getattr(L, s) is not None
    assert L.dfb_version() >= 100


def test_library_is_sm100a_only_and_has_no_torch_dependency(dfb):
    out = subprocess.run(["cuobjdump", "-lelf", dfb.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    ldd = subprocess.run(["ldd", dfb.lib_path()], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "libcudart" not in ldd  # static cudart, no framework dependency


def test_pod_layouts_match_reference(dfb):
    # intrinsics.h:20-30 — 48 bytes, ppx at 8, fx at 16; Extrinsics 12 floats
    I = dfb.Intrinsics
    assert C.sizeof(I) == 48 and I.ppx.offset == 8 and I.fx.offset == 16 and I.model.offset == 24 and I.coeffs.offset == 28
    assert C.sizeof(dfb.Extrinsics) == 48
    assert C.sizeof(dfb.CameraDesc) == 48 + 12 + 16 + 4 + 64
    assert dfb.FuseDesc.cam.offset == 8 and C.sizeof(dfb.FuseDesc) == 8 + 8 * 144 + 64 + 48


def test_no_cpu_fallback(dfb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on CPU-only hosts")
    L = dfb.lib()
    assert L.dfb_init(0, 0) == -3  # DFB_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.dfb_last_error()
    cfg = dfb.make_config(848, 480)
    with pytest.raises(dfb.DfbError) as e:
        dfb.Fuser(cfg.desc)
    assert e.value.code == -3
    with pytest.raises(dfb.DfbError):
        dfb.Workspace()
    assert L.dfb_selftest_division(10, 1) == -3


def test_depth_thresholds_match_oracle(dfb, oracle):
    from dfb_b200 import capi

    for args in [(0.5, 3.0, 0.001), (0.5, 2.0, 0.001), (0.5, 3.0, 0.00025), (0.0, 100.0, 0.001), (0.3, 0.2, 0.01)]:
        assert capi.depth_thresholds(*args) == oracle.thresholds(*args)


def test_fused_intrinsics_and_camera_info(dfb, oracle):
    from dfb_b200 import capi

    K = [631.7, 0, 640.9, 0, 632.2, 361.4, 0, 0, 1]
    intr = capi.camera_info_to_intrinsics(1280, 720, K, [0.1, 0.2, 0.3, 0.4, 0.5])
    # fusion_node.cpp:574-577 truncates through int
    assert (intr.fx, intr.fy, intr.ppx, intr.ppy, intr.model) == (631.0, 632.0, 640.0, 361.0, 4)
    assert list(intr.coeffs) == [np.float32(x) for x in (0.1, 0.2, 0.3, 0.4, 0.5)]
    for vertical in (False, True):
        a = capi.fused_intrinsics(intr, vertical)
        b = oracle.fused_intrinsics(intr, vertical)
        assert bytes(a) == bytes(b)
    v = capi.fused_intrinsics(intr, True)
    assert (v.width, v.height, v.fx, v.fy, v.ppx, v.ppy) == (720, 1280, 632.0, 631.0, 360.0, 640.0)


def test_interpolate_transform_matches_oracle_and_analytic(dfb, oracle):
    from dfb_b200 import capi, configs

    rng = np.random.default_rng(0)
    for _ in range(50):
        A, B = random_rigid(rng, 170), random_rigid(rng, 170)
        p, o = capi.interpolate_transform(A, B), oracle.interpolate_transform(A, B)
        np.testing.assert_allclose(p, o, rtol=0, atol=1e-14)
        np.testing.assert_allclose(capi.prerotate_z90(p), oracle.prerotate_z90(o), rtol=0, atol=1e-14)
    # analytic: halfway between identity and a rotation about Y by a, translation halved
    T = configs.affine(configs.rot_y(-8.0), [0.35, 0.01, -0.02])
    half = capi.interpolate_transform(np.eye(4), T)
    np.testing.assert_allclose(half[:3, :3], configs.rot_y(-4.0), atol=1e-12)
    np.testing.assert_allclose(half[:3, 3], [0.175, 0.005, -0.01], atol=1e-15)
    np.testing.assert_allclose(capi.interpolate_transform(np.eye(4), np.eye(4)), np.eye(4), atol=1e-15)
    # right camera upside down (registration_node.cpp:154-160): quaternion w < 0 -> inverse (fusion_node.cpp:603)
    up = configs.affine(configs.rot_z(180.0) @ configs.rot_x(2.0) @ configs.rot_y(-8.0), [0.35, 0.01, -0.02])
    h = capi.interpolate_transform(np.eye(4), up)
    np.testing.assert_allclose(h[:3, :3] @ h[:3, :3].T, np.eye(3), atol=1e-12)
    q_w_negative = oracle.interpolate_transform(np.eye(4), up)
    np.testing.assert_allclose(h, q_w_negative, atol=1e-14)
    # prerotate by exactly 90 degrees maps x->y, y->-x and rotates the translation too
    P = capi.prerotate_z90(configs.affine(np.eye(3), [1.0, 2.0, 3.0]))
    np.testing.assert_allclose(P, [[0, -1, 0, -2], [1, 0, 0, 1], [0, 0, 1, 3], [0, 0, 0, 1]], atol=1e-15)


def test_virtual_camera_transform(dfb):
    from dfb_b200 import capi, configs

    T = capi.virtual_camera_transform([0.1, -0.2, 0.3], [0, 0, 0])
    exp = configs.affine(configs.rot_z(90.0), [0, 0, 0])
    exp[:3, 3] += [-0.1, 0.2, -0.3]
    np.testing.assert_allclose(T, exp, atol=1e-15)
    T2 = capi.virtual_camera_transform([0, 0, 0], [10, 20, 30])
    R = configs.rot_x(10) @ configs.rot_y(20) @ configs.rot_z(30)
    np.testing.assert_allclose(T2[:3, :3], R.T @ configs.rot_z(90.0), atol=1e-12)


def test_config_descriptors(dfb):
    cfg = dfb.make_config(1280, 720)
    assert cfg.algorithmic_bytes == 11_980_800 and cfg.out_width == 720 and cfg.out_height == 1280
    assert dfb.make_config(848, 480).algorithmic_bytes == 5_291_520
    assert dfb.make_config(1280, 720, n_cameras=4).algorithmic_bytes == 21_196_800
    assert dfb.make_config(3840, 2160).algorithmic_bytes == 107_827_200
    d = cfg.desc
    assert d.cam[0].has_transform == 0 and d.cam[1].has_transform == 1 and d.mirror == 1
    assert list(d.cam[0].roi) == [0, 0, 1280, 720] and list(d.cam[1].roi) == [-1, -1, -1, -1]
    assert (d.out_intr.fx, d.out_intr.ppx, d.out_intr.ppy) == (631.0, 360.0, 640.0)


def test_wire_formats_of_the_depth_frameset_message(dfb):
    # camera_interfaces/msg/DepthFrameset.msg:1-7 as camera_node fills it (camera_node.cpp:191-209: mono16 LE step 2W, rgb8
    # step 3W) -> rows of the fuser's host pools; the fused sensor_msgs/Image header of fusion_node.cpp:799-808
    import ctypes as C

    from dfb_b200.ctypes_defs import ImageMsg, Intrinsics

    L = dfb.load()
    w, h = 20, 6
    rng = np.random.default_rng(4)
    depth = rng.integers(0, 65536, size=(h, w), dtype=np.uint16)
    color = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)

    def msg(payload, step, encoding, big=0):
        m = ImageMsg()
        m.height, m.width, m.encoding, m.is_bigendian, m.step = h, w, encoding, big, step
        m._keep = payload  # keep the buffer alive
        m.data, m.data_size = payload.ctypes.data, payload.nbytes
        return m

    # the camera node's own layout
    out_d = np.zeros((h, w), np.uint16)
    assert L.dfb_wire_unpack_depth(C.byref(msg(depth.view(np.uint8).reshape(h, -1).copy(), 2 * w, b"mono16")), w, h, out_d.ctypes.data) == 0
    assert np.array_equal(out_d, depth)
    out_c = np.zeros((h, w, 3), np.uint8)
    assert L.dfb_wire_unpack_color(C.byref(msg(color.reshape(h, -1).copy(), 3 * w, b"rgb8")), w, h, out_c.ctypes.data) == 0
    assert np.array_equal(out_c, color)
    # padded rows (step > width * bytes per pixel) and big-endian depth
    padded = np.full((h, 2 * w + 12), 0xAB, np.uint8)
    padded[:, : 2 * w] = depth.byteswap().view(np.uint8).reshape(h, -1)
    out_d[:] = 0
    assert L.dfb_wire_unpack_depth(C.byref(msg(padded, 2 * w + 12, b"16UC1", big=1)), w, h, out_d.ctypes.data) == 0
    assert np.array_equal(out_d, depth)
    padded_c = np.full((h, 3 * w + 5), 0xCD, np.uint8)
    padded_c[:, : 3 * w] = color.reshape(h, -1)
    out_c[:] = 0
    assert L.dfb_wire_unpack_color(C.byref(msg(padded_c, 3 * w + 5, b"rgb8")), w, h, out_c.ctypes.data) == 0
    assert np.array_equal(out_c, color)
    # refused: wrong encoding, wrong size, short step, short payload
    flat = color.reshape(h, -1).copy()
    assert L.dfb_wire_unpack_color(C.byref(msg(flat, 3 * w, b"bgr8")), w, h, out_c.ctypes.data) == -1
    assert L.dfb_wire_unpack_color(C.byref(msg(flat, 3 * w, b"rgb8")), w + 1, h, out_c.ctypes.data) == -1
    assert L.dfb_wire_unpack_color(C.byref(msg(flat, 3 * w - 1, b"rgb8")), w, h, out_c.ctypes.data) == -1
    short = msg(flat, 3 * w, b"rgb8")
    short.data_size = flat.nbytes - 1
    assert L.dfb_wire_unpack_color(C.byref(short), w, h, out_c.ctypes.data) == -1
    assert L.dfb_wire_unpack_depth(C.byref(msg(flat, 3 * w, b"rgb8")), w, h, out_d.ctypes.data) == -1
    # the fused image header
    fused = Intrinsics()
    fused.width, fused.height = 720, 1280
    frame = np.zeros((1280, 720, 3), np.uint8)
    o = ImageMsg()
    assert L.dfb_wire_fused_image(C.byref(fused), frame.ctypes.data, C.byref(o)) == 0
    assert (o.width, o.height, o.step, o.encoding, o.is_bigendian, o.data_size) == (720, 1280, 2160, b"rgb8", 0, frame.nbytes)
    assert o.data == frame.ctypes.data
