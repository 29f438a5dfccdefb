"""Drop-in boundary test: a C++ program written against include/pointcloud_processing/*.h exactly the way
fusion_node / registration_node use the reference library (tests/cpp/shim_fusion_test.cpp replays
FusionNode::processSyncedFrames, fusion_node.cpp:718-794, and the registration deprojection,
registration_node.cpp:313-333) must produce the oracle's results."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _binary(repo_root, name="shim_fusion_test"):
    exe = os.path.join(repo_root, "pointcloud-depthfusion_b200", "build", name)
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(repo_root, "pointcloud-depthfusion_b200"), "shimtest"], check=True)
    return exe


@pytest.mark.parametrize("post_filter", [0, 1, 2], ids=["splat", "gauss", "median"])
@pytest.mark.parametrize("kw", [dict(width=848, height=480, hole_fraction=0.05), dict(width=1280, height=720, profile="P1")],
                         ids=["C1", "C2-P1"])
def test_cpp_consumer_matches_oracle(gpu, dfb, oracle, repo_root, tmp_path, kw, post_filter):
    cfg = dfb.make_config(**kw)
    ds, cs = oracle.synth_frameset(cfg, 77)
    d = cfg.desc
    intr = d.cam[0].intr
    hdr = struct.pack("<2i4f3f2i2i4f", intr.width, intr.height, intr.fx, intr.fy, intr.ppx, intr.ppy, d.cam[0].depth_scale,
                      d.cam[0].min_depth, d.cam[0].max_depth, d.mirror, post_filter, d.out_intr.width, d.out_intr.height,
                      d.out_intr.fx, d.out_intr.fy, d.out_intr.ppx, d.out_intr.ppy)
    hdr += struct.pack("<4x") if len(hdr) % 8 else b""  # the doubles that follow are 8-byte aligned in the C struct
    hdr += np.asarray(cfg.cam_transforms[1], dtype=np.float64).tobytes()
    hdr += np.asarray(cfg.fused_transform, dtype=np.float64).tobytes()
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(hdr)
        for c in range(2):
            f.write(ds[c].tobytes())
            f.write(cs[c].tobytes())
    r = subprocess.run([_binary(repo_root), str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(outp, "rb").read()
    n_left, n_right = struct.unpack_from("<2I", raw, 0)
    npx = cfg.out_pixels * 3
    img = np.frombuffer(raw, dtype=np.uint8, count=npx, offset=8).reshape(cfg.out_height, cfg.out_width, 3)
    (n_reg,) = struct.unpack_from("<I", raw, 8 + npx)
    reg = np.frombuffer(raw, dtype=np.float32, count=n_reg * 4, offset=12 + npx).reshape(-1, 4)

    exp = oracle.fuse_frameset(d, ds, cs, mode=1)
    assert [n_left, n_right] == exp["counts"]
    want = exp["image"]
    if post_filter:
        want = oracle.filter_color(want, median=(post_filter == 2))
    assert np.array_equal(img, want)
    reg_exp = oracle.registration_cloud(d.cam[1], ds[1])
    assert n_reg == len(reg_exp) and np.array_equal(reg.view(np.uint32), reg_exp.view(np.uint32))


def _write_input(cfg, ds, cs, path, post_filter=0):
    d = cfg.desc
    intr = d.cam[0].intr
    hdr = struct.pack("<2i4f3f2i2i4f", intr.width, intr.height, intr.fx, intr.fy, intr.ppx, intr.ppy, d.cam[0].depth_scale,
                      d.cam[0].min_depth, d.cam[0].max_depth, d.mirror, post_filter, d.out_intr.width, d.out_intr.height,
                      d.out_intr.fx, d.out_intr.fy, d.out_intr.ppx, d.out_intr.ppy)
    hdr += struct.pack("<4x") if len(hdr) % 8 else b""
    hdr += np.asarray(cfg.cam_transforms[1], dtype=np.float64).tobytes()
    hdr += np.asarray(cfg.fused_transform, dtype=np.float64).tobytes()
    with open(path, "wb") as f:
        f.write(hdr)
        for c in range(2):
            f.write(ds[c].tobytes())
            f.write(cs[c].tobytes())


def test_lazy_pointcloud_is_indistinguishable_from_eager(gpu, dfb, oracle, repo_root, tmp_path):
    # Pointcloud records deproject / transform / append and fuses them into one launch in project(); every way of
    # looking at the cloud in between, or of pulling its sources away, must give what the eager operators give
    cfg = dfb.make_config(848, 480, hole_fraction=0.05)
    ds, cs = oracle.synth_frameset(cfg, 31)
    inp, outp = tmp_path / "lazy_in.bin", tmp_path / "lazy_out.bin"
    _write_input(cfg, ds, cs, inp)
    r = subprocess.run([_binary(repo_root, "shim_lazy_test"), str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(outp, "rb").read()
    npx = cfg.out_pixels * 3
    imgs = [np.frombuffer(raw, dtype=np.uint8, count=npx, offset=i * npx).reshape(cfg.out_height, cfg.out_width, 3) for i in range(5)]
    want = oracle.fuse_frameset(cfg.desc, ds, cs, mode=1)["image"]
    assert np.array_equal(imgs[0], want), "fused launch behind project() differs from the oracle"
    assert np.array_equal(imgs[1], want), "materialised (get() before project) differs"
    assert np.array_equal(imgs[2], want), "a source frame overwritten after deproject() leaked into the cloud"
    assert np.array_equal(imgs[3], imgs[4]), "three transforms: lazy and eager sequences disagree"
    assert not np.array_equal(imgs[3], want)
    (n,) = struct.unpack_from("<I", raw, 5 * npx)
    got = np.frombuffer(raw, dtype=np.float32, count=n * 4, offset=5 * npx + 4).reshape(-1, 4)
    cam1 = cfg.desc.cam[1]
    fdepth, _, indices, _ = oracle.filter_depth(ds[1].copy(), cfg.width, cfg.height, cam1.min_depth, cam1.max_depth, cam1.depth_scale)
    pts = oracle.deproject(fdepth, cam1.intr, indices, cam1.depth_scale, color=cs[1])
    T = np.asarray(cfg.cam_transforms[1], dtype=np.float64).astype(np.float32).reshape(16)
    pts = oracle.transform(pts, T)
    assert n == len(pts) and np.array_equal(got.view(np.uint32), pts.view(np.uint32)), "deproject + transform + copyToHost"


def test_cpp_consumer_align_frames_path(gpu, dfb, oracle, repo_root, tmp_path):
    # fusion_node with align_frames: setDepthFrameUnaligned + alignDepthToColor through the drop-in classes
    from helpers import align_case

    depth, di, ci, ex, scale = align_case(np.random.default_rng(23), depth_model=2, color_model=1, coeff_scale=0.05)
    inp, outp = tmp_path / "align_in.bin", tmp_path / "align_out.bin"
    with open(inp, "wb") as f:
        f.write(bytes(di) + bytes(ci) + bytes(ex) + struct.pack("<f", float(scale)) + depth.tobytes())
    r = subprocess.run([_binary(repo_root, "shim_align_test"), str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(outp, dtype=np.uint16).reshape(ci.height, ci.width)
    assert np.array_equal(got, oracle.align_depth_to_color(depth, di, ci, ex, float(scale)))
