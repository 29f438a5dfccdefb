"""GPU parity tests of the eager operators (one per reference Kernels::* wrapper), the registration feed,
the device synthetic generator and the division self test — all through the C ABI, against the oracle."""
import ctypes as C

import numpy as np
import pytest

from dfb_b200 import capi
from gpu_util import stream_handle, synth_pool_gpu, to_dev, u16
from helpers import random_rigid

pytestmark = pytest.mark.gpu


def test_shared_reciprocal_division_is_ieee(gpu, dfb):
    # dfb_math.cuh divide_pair == __fdiv_rn bit for bit on 2^28 operand triples of the projection's domain
    for seed in (1, 2):
        assert dfb.lib().dfb_selftest_division(1 << 28, seed) == 0


def test_device_generator_matches_cpu_generator(gpu, dfb, oracle):
    for kw in [dict(width=848, height=480, hole_fraction=0.1), dict(width=1280, height=720, n_cameras=4, density="sparse", profile="P1")]:
        cfg = dfb.make_config(**kw)
        d_dev, c_dev = synth_pool_gpu(gpu, dfb, cfg, 17, 2)
        d_cpu, c_cpu = oracle.synth_pool(cfg, 17, 2)
        assert np.array_equal(u16(d_dev), d_cpu)
        assert np.array_equal(c_dev.cpu().numpy(), c_cpu)


@pytest.mark.parametrize("roi", [(-1, -1, -1, -1), (0, 0, 848, 480), (100, 50, 500, 300), (840, 470, 100, 100)])
def test_filter_depth(gpu, dfb, oracle, roi):
    torch = gpu
    cfg = dfb.make_config(848, 480, hole_fraction=0.3)
    ds, _ = oracle.synth_frameset(cfg, 3)
    w, h = 848, 480
    stale = np.zeros((h, w), dtype=np.uint8)
    stale[0, :7] = 255  # mask state left over from an earlier frame (App. A.2)
    d = to_dev(torch, ds[1].view(np.int16))
    m = to_dev(torch, stale)
    idx = torch.zeros(w * h, dtype=torch.int32, device="cuda")
    ws = dfb.Workspace()
    cnt = C.c_uint32()
    r = (C.c_int32 * 4)(*roi)
    capi.check(dfb.lib().dfb_op_filter_depth(ws.handle, d.data_ptr(), m.data_ptr(), idx.data_ptr(), w, h, 0.5, 3.0, 0.001, r,
                                              C.byref(cnt), stream_handle(torch)))
    fd, mask, indices, count = oracle.filter_depth(ds[1], w, h, 0.5, 3.0, 0.001, roi, mask=stale)
    assert cnt.value == count
    assert np.array_equal(u16(d), fd) and np.array_equal(m.cpu().numpy(), mask)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32)[:count], indices)


@pytest.mark.parametrize("w,h,roi,offset", [(37, 23, (5, 3, 20, 11), 0), (37, 23, (-1, -1, -1, -1), 0), (124, 9, (3, 0, 118, 9), 0),
                                            (848, 480, (100, 50, 500, 300), 1), (64, 16, (8, 2, 48, 12), 3)])
def test_filter_depth_vector_and_scalar_forms(gpu, dfb, oracle, w, h, roi, offset):
    # the filter moves 8 pixels per thread as one 16-byte piece when it can: groups that straddle image rows (width not a
    # multiple of 8), ROI edges inside a group, a stale mask outside the ROI, and buffers that are not 16-byte aligned
    # (offset > 0: the scalar form)
    torch = gpu
    rng = np.random.default_rng(w * 1000 + h + offset)
    depth = rng.integers(0, 4000, size=(h, w), dtype=np.uint16)
    depth[rng.random((h, w)) < 0.2] = 0
    stale = np.where(rng.random((h, w)) < 0.3, 255, 0).astype(np.uint8)
    dbuf = torch.zeros(w * h + 16, dtype=torch.int16, device="cuda")
    mbuf = torch.zeros(w * h + 16, dtype=torch.uint8, device="cuda")
    d = dbuf[offset:offset + w * h]
    m = mbuf[offset:offset + w * h]
    d.copy_(torch.from_numpy(depth.view(np.int16).reshape(-1)))
    m.copy_(torch.from_numpy(stale.reshape(-1)))
    idx = torch.zeros(w * h, dtype=torch.int32, device="cuda")
    ws = dfb.Workspace()
    cnt = C.c_uint32()
    r = (C.c_int32 * 4)(*roi)
    capi.check(dfb.lib().dfb_op_filter_depth(ws.handle, d.data_ptr(), m.data_ptr(), idx.data_ptr(), w, h, 0.5, 3.0, 0.001, r,
                                              C.byref(cnt), stream_handle(torch)))
    fd, mask, indices, count = oracle.filter_depth(depth, w, h, 0.5, 3.0, 0.001, roi, mask=stale)
    assert cnt.value == count
    assert np.array_equal(d.cpu().numpy().view(np.uint16).reshape(h, w), fd)
    assert np.array_equal(m.cpu().numpy().reshape(h, w), mask)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32)[:count], indices)
    assert not dbuf[offset + w * h:].any() and not mbuf[offset + w * h:].any() and not dbuf[:offset].any()   # nothing outside


def test_eager_pipeline_equals_fused_and_oracle(gpu, dfb, oracle):
    """filter -> deproject -> transform -> append -> transform -> project with the eager operators, i.e. the
    reference's own operator sequence (fusion_node.cpp:739-784), gives the oracle's cloud bit for bit and the
    same image as the fused kernel."""
    torch = gpu
    L = dfb.lib()
    s = stream_handle(torch)
    cfg = dfb.make_config(848, 480, hole_fraction=0.05)
    ds, cs = oracle.synth_frameset(cfg, 8)
    w, h = cfg.width, cfg.height
    ws = dfb.Workspace()
    cloud = torch.zeros((2 * w * h, 4), dtype=torch.float32, device="cuda")
    counts = []
    for c in range(2):
        cd = cfg.desc.cam[c]
        d = to_dev(torch, ds[c].view(np.int16))
        col = to_dev(torch, cs[c])
        m = torch.zeros(w * h, dtype=torch.uint8, device="cuda")
        idx = torch.zeros(w * h, dtype=torch.int32, device="cuda")
        cnt = C.c_uint32()
        capi.check(L.dfb_op_filter_depth(ws.handle, d.data_ptr(), m.data_ptr(), idx.data_ptr(), w, h, cd.min_depth, cd.max_depth,
                                         cd.depth_scale, cd.roi, C.byref(cnt), s))
        seg = cloud[sum(counts):]
        capi.check(L.dfb_op_deproject(seg.data_ptr(), d.data_ptr(), col.data_ptr(), C.byref(cd.intr), idx.data_ptr(), cnt.value,
                                      cd.depth_scale, s))
        if cd.has_transform:
            capi.check(L.dfb_op_transform(seg.data_ptr(), cnt.value, cd.to_ref, s))
        counts.append(cnt.value)
    total = sum(counts)
    capi.check(L.dfb_op_transform(cloud.data_ptr(), total, cfg.desc.fused_transform, s))
    keys = torch.empty(cfg.out_pixels, dtype=torch.int64, device="cuda")
    capi.check(L.dfb_op_keys_clear(keys.data_ptr(), cfg.out_pixels, s))
    img = torch.empty((cfg.out_height, cfg.out_width, 3), dtype=torch.uint8, device="cuda")
    seg_counts = (C.c_uint32 * 2)(*counts)
    capi.check(L.dfb_op_project(img.data_ptr(), keys.data_ptr(), cloud.data_ptr(), seg_counts, 2, C.byref(cfg.desc.out_intr),
                                cfg.desc.mirror, s))
    torch.cuda.synchronize()
    ref = oracle.fuse_frameset(cfg.desc, ds, cs, mode=1, want_cloud=True)
    assert counts == ref["counts"]
    got = cloud[:total].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), ref["cloud"].view(np.uint32)), "cloud differs from the oracle (bit level)"
    assert np.array_equal(img.cpu().numpy(), ref["image"])
    assert bool((keys == -1).all()), "keys must be all-ones again after dfb_op_project"
    # fused kernel, same input
    fuser = dfb.Fuser(cfg.desc)
    out = torch.empty_like(img)
    dd = to_dev(torch, np.stack(ds).view(np.int16))
    cc = to_dev(torch, np.stack(cs))
    fuser.fuse_device(dd.data_ptr(), cc.data_ptr(), out.data_ptr(), 1, s)
    torch.cuda.synchronize()
    assert torch.equal(out, img)


def test_deproject_without_colour_and_transform_kat(gpu, dfb, oracle):
    torch = gpu
    L = dfb.lib()
    s = stream_handle(torch)
    rng = np.random.default_rng(4)
    pts = rng.uniform(-3, 3, size=(100003, 4)).astype(np.float32)
    T = random_rigid(rng).astype(np.float32).reshape(16)
    d = to_dev(torch, pts)
    capi.check(L.dfb_op_transform(d.data_ptr(), len(pts), (C.c_float * 16)(*T), s))
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32), oracle.transform(pts, T).view(np.uint32))


@pytest.mark.parametrize("density", ["dense", "sparse"])
def test_registration_cloud(gpu, dfb, oracle, density):
    # BASELINE configs[2]: deproject of an 848x480 depth pair feeding GICP (registration_node.cpp:313-333)
    torch = gpu
    cfg = dfb.make_config(848, 480, density=density, hole_fraction=0.1)
    ds, _ = oracle.synth_frameset(cfg, 1)
    ws = dfb.Workspace()
    for c in range(2):
        cd = cfg.desc.cam[c]
        d = to_dev(torch, ds[c].view(np.int16))
        pts = torch.zeros((848 * 480, 4), dtype=torch.float32, device="cuda")
        cnt = C.c_uint32()
        capi.check(dfb.lib().dfb_registration_cloud(ws.handle, C.byref(cd), d.data_ptr(), pts.data_ptr(), C.byref(cnt),
                                                    stream_handle(torch)))
        exp = oracle.registration_cloud(cd, ds[c])
        assert cnt.value == len(exp)
        assert np.array_equal(pts[:cnt.value].cpu().numpy().view(np.uint32), exp.view(np.uint32))
        assert np.array_equal(u16(d), ds[c])  # input depth is not modified


ALIGN_CASES = [
    ("D455-like: no distortion applies", dict(depth_model=4, color_model=2, coeff_scale=0.05)),
    ("inverse Brown-Conrady depth", dict(depth_model=2, color_model=0, coeff_scale=0.05)),
    ("modified Brown-Conrady colour", dict(depth_model=0, color_model=1, coeff_scale=0.05)),
    ("both, larger coefficients, same size", dict(depth_model=2, color_model=1, coeff_scale=0.2, dw=424, dh=240, cw=424, ch=240)),
]


@pytest.mark.parametrize("name,kw", ALIGN_CASES, ids=[c[0] for c in ALIGN_CASES])
def test_align_depth_to_color_matches_oracle(gpu, dfb, oracle, name, kw):
    # Frameset::alignDepthToColor (fusion_node.cpp:443-457 with align_frames): bit-exact, incl. the pixel-pair quirk
    from helpers import align_case

    torch = gpu
    depth, di, ci, ex, scale = align_case(np.random.default_rng(11), **kw)
    d = to_dev(torch, depth.view(np.int16))
    out = torch.zeros((ci.height, ci.width), dtype=torch.int16, device="cuda")
    capi.check(dfb.lib().dfb_op_align_depth_to_color(out.data_ptr(), d.data_ptr(), C.byref(di), C.byref(ci), C.byref(ex), float(scale),
                                                     stream_handle(torch)))
    torch.cuda.synchronize()
    got = u16(out)
    exp = oracle.align_depth_to_color(depth, di, ci, ex, float(scale))
    assert np.array_equal(got, exp)
    assert (got > 0).mean() > 0.5 and np.array_equal(u16(d), depth)  # a real image came out; the input is untouched


def test_align_depth_to_color_rejects_unsupported_models(gpu, dfb):
    from helpers import align_case

    torch = gpu
    depth, di, ci, ex, scale = align_case(np.random.default_rng(1), dw=64, dh=32, cw=64, ch=32)
    d = to_dev(torch, depth.view(np.int16))
    out = torch.zeros((32, 64), dtype=torch.int16, device="cuda")
    di.model = 3  # F-theta cannot be DEprojected (kernels.cu:48-49 asserts it away)
    assert dfb.lib().dfb_op_align_depth_to_color(out.data_ptr(), d.data_ptr(), C.byref(di), C.byref(ci), C.byref(ex), float(scale),
                                                 stream_handle(torch)) == -1
    ci.model, di.model = 0, 1  # the reference asserts modified Brown-Conrady away for deprojection
    assert dfb.lib().dfb_op_align_depth_to_color(out.data_ptr(), d.data_ptr(), C.byref(di), C.byref(ci), C.byref(ex), float(scale),
                                                 stream_handle(torch)) == -1


def _voxel(dfb, torch, ws, pts_np, leaf):
    n = len(pts_np)
    d_in = to_dev(torch, np.ascontiguousarray(pts_np, dtype=np.float32)) if n else torch.zeros((1, 4), dtype=torch.float32, device="cuda")
    d_out = torch.zeros((max(n, 1), 4), dtype=torch.float32, device="cuda")
    cnt = C.c_uint32()
    rc = dfb.lib().dfb_voxel_downsample(ws.handle, d_in.data_ptr(), n, leaf, d_out.data_ptr(), C.byref(cnt), stream_handle(torch))
    return rc, d_out[:cnt.value].cpu().numpy()


def test_voxel_downsample_registration_path(gpu, dfb, oracle):
    # BASELINE configs[2]: deproject + voxel downsample of an 848x480 depth pair feeding GICP
    # (registration_node.cpp:313-333, registration.cpp:144-150, leaf 0.01 m)
    torch = gpu
    cfg = dfb.make_config(848, 480, hole_fraction=0.1)
    ds, _ = oracle.synth_frameset(cfg, 3)
    ws = dfb.Workspace()
    for c in range(2):
        cloud = oracle.registration_cloud(cfg.desc.cam[c], ds[c])
        rc, got = _voxel(dfb, torch, ws, cloud, 0.01)
        assert rc == 0, dfb.last_error()
        exp, ids = oracle.voxel_exact(cloud, 0.01)
        assert len(got) == len(exp)                       # one centroid per occupied voxel
        assert np.array_equal(got[:, 3], np.ones(len(got), np.float32))
        # same voxel order (iz, iy, ix) and the same centroid bits: double-precision sums in ascending source order
        # on both sides (the radix sort is stable, the oracle's comparator breaks ties by source index)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
        # a centroid lies in its own voxel (checked away from voxel faces, where re-quantising the mean is ill-posed)
        inv = np.float32(1.0) / np.float32(0.01)
        q = got[:, :3] * inv
        interior = np.abs(q - np.round(q)).min(axis=1) > 1e-3
        assert np.array_equal(np.floor(q).astype(np.int32)[interior], ids[interior])
        assert np.all(np.diff(ids[:, 2].astype(np.int64) * (1 << 42) + ids[:, 1].astype(np.int64) * (1 << 21) + ids[:, 0]) > 0)
        # PCL's approximate grid (restated in the oracle) may emit several centroids per voxel, never fewer
        ap = oracle.voxel_approximate(cloud, 0.01)
        assert len(got) <= len(ap) <= len(cloud)
        # deterministic: a second call returns identical bits
        rc, again = _voxel(dfb, torch, ws, cloud, 0.01)
        assert rc == 0 and np.array_equal(got.view(np.uint32), again.view(np.uint32))


def test_voxel_downsample_edge_cases(gpu, dfb, oracle):
    torch = gpu
    ws = dfb.Workspace()
    rc, out = _voxel(dfb, torch, ws, np.zeros((0, 4), np.float32), 0.01)
    assert rc == 0 and len(out) == 0                      # empty cloud
    one = np.array([[0.105, -0.204, 1.5, 1.0]] * 5, np.float32)
    rc, out = _voxel(dfb, torch, ws, one, 0.01)
    assert rc == 0 and len(out) == 1 and np.allclose(out[0, :3], one[0, :3], rtol=1e-6)
    rng = np.random.default_rng(9)
    for n in (1, 2, 2047, 2048, 2049, 5000, 70001):       # around the sort-tile and scan-tile sizes, not powers of two
        pts = np.ones((n, 4), np.float32)
        pts[:, :3] = rng.uniform(-0.2, 0.2, size=(n, 3)).astype(np.float32)
        rc, got = _voxel(dfb, torch, ws, pts, 0.02)
        exp, _ = oracle.voxel_exact(pts, 0.02)
        assert rc == 0 and len(got) == len(exp)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    # a cloud spread over +-150 m at 1 cm: 3 x 15 bits of occupied box = six radix passes, many points per voxel
    # (duplicates) and negative indices on every axis
    pts = np.ones((40000, 4), np.float32)
    pts[:, :3] = rng.uniform(-150.0, 150.0, size=(40000, 3)).astype(np.float32)
    pts[20000:] = pts[rng.integers(0, 20000, size=20000)]
    pts[20000:, :3] += rng.uniform(0.0, 1e-4, size=(20000, 3)).astype(np.float32)
    rc, got = _voxel(dfb, torch, ws, pts, 0.01)
    exp, _ = oracle.voxel_exact(pts, 0.01)
    assert rc == 0 and len(got) == len(exp) and len(got) < len(pts)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    bad = pts[:8].copy()
    bad[3, 1] = np.nan                                     # a non-finite coordinate is refused as well
    rc, _ = _voxel(dfb, torch, ws, bad, 0.01)
    assert rc != 0 and "voxel index out of range" in dfb.last_error()
    far = np.array([[2.0e4, 0.0, 0.0, 1.0]], np.float32)   # |p / leaf| >= 2^20: refused, not wrapped
    rc, _ = _voxel(dfb, torch, ws, far, 0.01)
    assert rc != 0 and "voxel index out of range" in dfb.last_error()
    rc, _ = _voxel(dfb, torch, ws, one, 0.0)
    assert rc != 0


@pytest.mark.parametrize("median", [False, True])
def test_filter_color_matches_oracle(gpu, dfb, oracle, median):
    torch = gpu
    rng = np.random.default_rng(6)
    for (w, h) in [(720, 1280), (37, 5), (3, 3), (2, 9)]:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        img[rng.random((h, w)) < 0.3] = 0  # holes, as a splat leaves them
        d = to_dev(torch, img)
        scratch = torch.empty_like(d)
        capi.check(dfb.lib().dfb_op_filter_color(d.data_ptr(), scratch.data_ptr(), w, h, int(median), stream_handle(torch)))
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), oracle.filter_color(img, median=median))


def _voxel_approx(dfb, torch, ws, pts_np, leaf):
    n = len(pts_np)
    d_in = to_dev(torch, np.ascontiguousarray(pts_np, dtype=np.float32)) if n else torch.zeros((1, 4), dtype=torch.float32, device="cuda")
    d_out = torch.zeros((max(n, 1), 4), dtype=torch.float32, device="cuda")
    cnt = C.c_uint32()
    rc = dfb.lib().dfb_voxel_downsample_approximate(ws.handle, d_in.data_ptr(), n, leaf, d_out.data_ptr(), C.byref(cnt),
                                                    stream_handle(torch))
    return rc, d_out[:cnt.value].cpu().numpy()


def test_voxel_approximate_has_pcl_semantics(gpu, dfb, oracle):
    # pcl::ApproximateVoxelGrid restated in the oracle (512-entry history, float sums in visiting order, eviction order):
    # the GPU version reproduces it point for point, bit for bit, in the same order — duplicates included
    torch = gpu
    ws = dfb.Workspace()
    cfg = dfb.make_config(848, 480, hole_fraction=0.1)
    ds, _ = oracle.synth_frameset(cfg, 3)
    for c in range(2):
        cloud = oracle.registration_cloud(cfg.desc.cam[c], ds[c])
        rc, got = _voxel_approx(dfb, torch, ws, cloud, 0.01)
        assert rc == 0, dfb.last_error()
        exp = oracle.voxel_approximate(cloud, 0.01)
        assert len(got) == len(exp)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    rng = np.random.default_rng(17)
    for n in (0, 1, 2, 31, 32, 33, 1023, 1024, 1025, 5000, 70001):
        pts = np.ones((n, 4), np.float32)
        pts[:, :3] = rng.uniform(-0.3, 0.3, size=(n, 3)).astype(np.float32)
        if n > 100:
            pts[50:90, :3] = pts[50, :3]  # a long run inside one voxel
        for leaf in (0.02, 0.5, 10.0):   # many voxels .. everything in one
            rc, got = _voxel_approx(dfb, torch, ws, pts, leaf)
            exp = oracle.voxel_approximate(pts, leaf)
            assert rc == 0 and len(got) == len(exp), (n, leaf)
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (n, leaf)


def test_exclusive_scan_building_block(gpu, dfb):
    # csrc/scan.cu under both voxel grids: tile scan + last-block tile offsets + add, against torch.cumsum (int64, mod 2^32)
    torch = gpu
    L = dfb.lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    for n in (0, 1, 3, 4095, 4096, 4097, 70001, 4096 * 1024 + 5, 6_000_001):   # > 1024 tiles: two rounds in the last block
        for off in (0, 1):                                                     # 16-byte aligned and not
            buf = torch.randint(0, 1 << 20, (n + 8,), dtype=torch.int64, device="cuda", generator=g)
            if n > 100:
                buf[:50] = 0xFFFFFFF0                                          # wraps mod 2^32 like the kernel's u32 sums
            data = buf.to(torch.int32)                                         # keeps the low 32 bits
            view = data[off:off + n]
            want = (torch.cumsum(buf[off:off + n], 0) - buf[off:off + n]) & 0xFFFFFFFF
            total = C.c_uint32(123)
            capi.check(L.dfb_selftest_exclusive_scan(view.data_ptr() if n else None, n, C.byref(total), stream_handle(torch)))
            got = view.to(torch.int64) & 0xFFFFFFFF
            assert torch.equal(got, want), (n, off)
            assert total.value == (int(buf[off:off + n].sum().item()) & 0xFFFFFFFF), (n, off)
            assert torch.equal(data[off + n:], buf[off + n:].to(torch.int32))   # nothing written past the end


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("DFB_RANDOM_CLOUDS", "8"))))
def test_voxel_grids_on_random_clouds(gpu, dfb, oracle, seed):
    # both voxel grids bit for bit against the oracle on clouds of random size, extent, leaf and clustering
    # (duplicates, runs inside one voxel, negative indices, 1 .. 6 radix passes)
    torch = gpu
    rng = np.random.default_rng(500 + seed)
    n = int(rng.choice([1, 7, 1023, 1025, 4097, 20000, 150001]))
    extent = float(rng.choice([0.05, 1.0, 8.0, 300.0]))
    leaf = float(rng.choice([0.005, 0.01, 0.07, 0.5]))
    pts = np.ones((n, 4), np.float32)
    pts[:, :3] = rng.uniform(-extent, extent, size=(n, 3)).astype(np.float32)
    if n > 100:
        k = n // 3
        pts[:k] = pts[rng.integers(k, n, size=k)]                       # exact duplicates, scattered
        pts[k:2 * k, :3] = pts[k, :3] + rng.uniform(0, leaf / 4, size=(k, 3)).astype(np.float32)   # one long cluster
        pts = pts[rng.permutation(n)] if seed % 2 else pts
    ws = dfb.Workspace()
    rc, got = _voxel(dfb, torch, ws, pts, leaf)
    assert rc == 0, dfb.last_error()
    exp, _ = oracle.voxel_exact(pts, leaf)
    assert len(got) == len(exp) and np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (n, extent, leaf)
    rc, got = _voxel_approx(dfb, torch, ws, pts, leaf)
    assert rc == 0, dfb.last_error()
    exp = oracle.voxel_approximate(pts, leaf)
    assert len(got) == len(exp) and np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (n, extent, leaf)
