"""GPU parity tests of the fused hot path, through the C ABI (dfb_fuser_*), against the CPU oracle.

Bar (BASELINE.json north_star): the fused image is bit-exact against the oracle's canonical splat
(minimum 64-bit key); against the oracle's SEQUENTIAL restatement of the reference it may differ only on
pixels the oracle flags ambiguous (same camera, identical Z bits, different colour), < 1e-4 of the image."""
import numpy as np
import pytest

from gpu_util import fuse_pool, stream_handle, synth_pool_gpu, to_dev, u16
from helpers import random_rigid, simple_desc

pytestmark = pytest.mark.gpu

CASES = [
    # name, make_config kwargs
    ("C1-P0-dense", dict(width=848, height=480)),
    ("C1-P1-dense", dict(width=848, height=480, profile="P1")),
    ("C1-P0-sparse-holes", dict(width=848, height=480, density="sparse", hole_fraction=0.2)),
    ("C2-P0-dense", dict(width=1280, height=720)),
    ("C2-P1-sparse", dict(width=1280, height=720, profile="P1", density="sparse")),
    ("C2-horizontal-nomirror", dict(width=1280, height=720, vertical=False, mirror=False, hole_fraction=0.05)),
    ("C4-4cam", dict(width=1280, height=720, n_cameras=4)),
    ("3cam-small-roi", dict(width=424, height=240, n_cameras=3, roi=[(10, 20, 300, 200), (-1, -1, -1, -1), (0, 5, 9999, 100)])),
]


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
def test_fused_matches_oracle(gpu, dfb, oracle, name, kw):
    cfg = dfb.make_config(**kw)
    nfs = 2
    depth, color = oracle.synth_pool(cfg, 100, nfs)
    fuser = dfb.Fuser(cfg.desc)
    outs = []
    for variant in (0, 1, 2, 3, 4):  # automatic, row, column, TMA-tiled, single-launch stream: identical keys
        fuser.set_variant(variant)
        outs.append(fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width)))
    out = outs[0]
    for i in (1, 2, 3, 4):
        assert np.array_equal(outs[i], out), f"kernel variant {i} disagrees with the automatic choice"
    for f in range(nfs):
        can = oracle.fuse_frameset(cfg.desc, list(depth[f]), list(color[f]), mode=1)
        assert np.array_equal(out[f], can["image"]), f"{name} frameset {f}: fused image differs from canonical oracle"
        seq = oracle.fuse_frameset(cfg.desc, list(depth[f]), list(color[f]), mode=0, want_ambiguous=True, fast=True)
        diff = (out[f] != seq["image"]).any(axis=2)
        assert not (diff & (seq["ambiguous"] == 0)).any(), "differs from the sequential reference on an unambiguous pixel"
        assert diff.mean() <= 1e-4
    fuser.close()


@pytest.mark.parametrize("size", [(848, 480, 418.0), (1280, 720, 631.0), (3840, 2160, 1893.0)])
@pytest.mark.parametrize("mirror", [False, True])
def test_identity_roundtrip_full_size(gpu, dfb, size, mirror):
    # KAT B.3-1/2 at BASELINE.json's full sizes (no oracle needed): identity transforms, same intrinsics
    w, h, f = size
    rng = np.random.default_rng(w + mirror)
    desc = simple_desc(w, h, fx=f, mirror=mirror)
    depth = rng.integers(400, 3100, size=(1, 1, h, w), dtype=np.uint16)
    color = rng.integers(0, 256, size=(1, 1, h, w, 3), dtype=np.uint8)
    fuser = dfb.Fuser(desc)
    out = fuse_pool(gpu, fuser, depth, color, (h, w))[0]
    valid = (depth[0, 0] >= 499) & (depth[0, 0] <= 2999)
    exp = np.where(valid[..., None], color[0, 0], 0).astype(np.uint8)
    if mirror:
        exp = exp[:, ::-1]
    assert np.array_equal(out, exp)


def test_group_size_invariance_and_key_rearm(gpu, dfb, oracle):
    cfg = dfb.make_config(848, 480)
    depth, color = oracle.synth_pool(cfg, 0, 5)
    fuser = dfb.Fuser(cfg.desc)
    ref = None
    for g in (0, 1, 2, 5, 7):
        fuser.set_group(g)
        out = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
        again = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
        assert np.array_equal(out, again), "second run differs: keys were not re-armed"
        if ref is None:
            ref = out
        assert np.array_equal(out, ref), f"group={g} changes the result"
    # stale keys must not leak into a sparser frameset
    sparse = dfb.make_config(848, 480, density="sparse")
    d2, c2 = oracle.synth_pool(sparse, 0, 1)
    fuser.update(sparse.desc)
    out2 = fuse_pool(gpu, fuser, d2, c2, (cfg.out_height, cfg.out_width))
    can = oracle.fuse_frameset(sparse.desc, list(d2[0]), list(c2[0]), mode=1)
    assert np.array_equal(out2[0], can["image"])


@pytest.mark.parametrize("slots,lag", [(1, 0), (2, 1), (3, 1), (5, 4), (16, 8)])
def test_stream_kernel_ring_dependencies(gpu, dfb, oracle, monkeypatch, slots, lag):
    # the single-launch kernel with tiny / odd key rings: every frameset waits on a resolve (slot reuse) and every
    # resolve on its splat; 40 framesets over 3 cameras, each compared with the canonical oracle
    monkeypatch.setenv("DFB_STREAM_SLOTS", str(slots))
    monkeypatch.setenv("DFB_STREAM_LAG", str(lag))
    cfg = dfb.make_config(424, 240, n_cameras=3, hole_fraction=0.1)
    nfs = 40
    depth, color = oracle.synth_pool(cfg, 7, nfs)
    fuser = dfb.Fuser(cfg.desc)   # reads the ring geometry at creation
    fuser.set_variant(4)
    out = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
    again = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
    assert np.array_equal(out, again), "second batch differs: the ring was not left armed"
    for f in range(nfs):
        can = oracle.fuse_frameset(cfg.desc, list(depth[f]), list(color[f]), mode=1, fast=True)
        assert np.array_equal(out[f], can["image"]), f"slots={slots} lag={lag}: frameset {f} differs from the oracle"
    fuser.close()


@pytest.mark.parametrize("slots", [1, 2, 3, 8])
def test_ring_epochs_ignore_what_earlier_uses_left_behind(gpu, dfb, oracle, monkeypatch, slots):
    # The stream kernel re-arms a ring slot only every third use; in between, the keys of a later use are biased below
    # everything an earlier use left behind (dfb_math.cuh key_epoch_bias). The worst case for that: a NEAR, dense
    # frameset followed by FAR, sparse ones through the same slot — a stale near key must neither win a pixel nor fill a
    # hole. Sequences of every length carry the ring phase from launch to launch (single framesets, odd batches), and
    # the two-launch variants in between need the ring cleared for them.
    monkeypatch.setenv("DFB_STREAM_SLOTS", str(slots))
    monkeypatch.setenv("DFB_STREAM_LAG", str(slots // 2))
    cfg = dfb.make_config(424, 240, hole_fraction=0.0)
    rng = np.random.default_rng(21)
    shape = (2, cfg.height, cfg.width)
    near_d = np.full(shape, 700, np.uint16)
    far_d = np.where(rng.random(shape) < 0.5, 2900, 0).astype(np.uint16)
    far2_d = np.where(rng.random(shape) < 0.3, 1800, 0).astype(np.uint16)
    sets = []
    for d in (near_d, far_d, far2_d):
        c = rng.integers(1, 256, size=shape + (3,), dtype=np.uint8)
        sets.append((d, c, oracle.fuse_frameset(cfg.desc, list(d), list(c), mode=1, fast=True)["image"]))
    assert (sets[1][2].any(axis=2) < sets[0][2].any(axis=2)).any()   # the far frameset really has holes where near has points
    fuser = dfb.Fuser(cfg.desc)
    fuser.set_variant(4)
    hw = (cfg.out_height, cfg.out_width)
    order = rng.integers(0, 3, size=64)
    order[:6] = [0, 1, 2, 0, 1, 1]
    pos = 0
    for n in (1, 1, 1, 2, 1, 3, 5, 1, 7, 4, 1, 1, 9, 11, 16):   # 64 framesets in launches of every kind
        pick = order[pos:pos + n]
        pos += n
        out = fuse_pool(gpu, fuser, np.stack([sets[i][0] for i in pick]), np.stack([sets[i][1] for i in pick]), hw)
        for k, i in enumerate(pick):
            assert np.array_equal(out[k], sets[i][2]), f"slots={slots}: launch of {n}, frameset {k} (kind {i})"
        if pos in (9, 27):   # a two-launch variant in between: it must find (and leave) KEY_EMPTY everywhere
            fuser.set_variant(2)
            out = fuse_pool(gpu, fuser, np.stack([sets[1][0], sets[0][0]]), np.stack([sets[1][1], sets[0][1]]), hw)
            assert np.array_equal(out[0], sets[1][2]) and np.array_equal(out[1], sets[0][2])
            fuser.set_variant(4)
    assert pos == 64
    fuser.close()


def test_stream_scratch_does_not_overlap_the_ring_of_the_grouped_variants(gpu, dfb, oracle):
    # the stream kernel parks rejected atomics in a scratch area BEHIND the key ring; the ring is allocated for
    # max(group, stream slots) framesets and the grouped variants (1-3) index it by frameset-in-group: with the automatic
    # group (15 at 848x480) larger than the stream kernel's slots (14), junk keys in slot 14 would end up in frameset 14
    cfg = dfb.make_config(848, 480, hole_fraction=0.1, profile="P1")  # P1: 43 % of the points leave the image -> scratch
    nfs = 20
    depth, color = oracle.synth_pool(cfg, 40, nfs)
    fuser = dfb.Fuser(cfg.desc)
    fuser.set_variant(4)
    ref = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
    for variant in (1, 2, 3, 4):
        fuser.set_variant(variant)
        out = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
        assert np.array_equal(out, ref), f"variant {variant} after the stream kernel differs"
    for f in (0, 13, 14, 15, 19):
        can = oracle.fuse_frameset(cfg.desc, list(depth[f]), list(color[f]), mode=1, fast=True)
        assert np.array_equal(ref[f], can["image"]), f"frameset {f}"
    # the same with a group that shrinks below the ALLOCATED ring and grows back: the scratch area must stay behind the
    # allocation, not behind whatever the current group needs
    fuser.set_group(20)
    fuser.set_variant(2)
    assert np.array_equal(fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width)), ref)
    fuser.set_group(3)
    fuser.set_variant(4)
    assert np.array_equal(fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width)), ref)
    fuser.set_group(20)
    fuser.set_variant(2)
    assert np.array_equal(fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width)), ref)
    fuser.close()


def test_measured_configuration_against_the_oracle(gpu, dfb, oracle):
    # the configuration bench.py times — C2, 512 framesets in one launch, default ring (it wraps 85 times) — with
    # framesets spread over the batch, both sides of the first slot reuse included, against the canonical oracle
    torch = gpu
    cfg = dfb.make_config(1280, 720)
    F = 512
    depth, color = synth_pool_gpu(torch, dfb, cfg, 0, F)
    sh = stream_handle(torch)
    out = torch.zeros((F, cfg.out_height, cfg.out_width, 3), dtype=torch.uint8, device="cuda")
    fuser = dfb.Fuser(cfg.desc)
    for _ in range(2):  # the second pass starts from the ring the first one left behind
        fuser.fuse_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
    torch.cuda.synchronize()
    for f in (0, 5, 6, 7, 11, 12, 13, 255, 256, 500, 505, 506, 511):
        d, c = oracle.synth_pool(cfg, f, 1)
        # the device generator and the oracle's generator are the same function of (config, frameset)
        assert np.array_equal(depth[f].cpu().numpy().view(np.uint16), d[0])
        can = oracle.fuse_frameset(cfg.desc, list(d[0]), list(c[0]), mode=1, fast=True)
        assert np.array_equal(out[f].cpu().numpy(), can["image"]), f"frameset {f} of the 512-frameset launch differs from the oracle"
    fuser.close()


@pytest.mark.parametrize("mode", [1, 2], ids=["gauss", "median"])
def test_fuser_post_filter_matches_eager_sequence(gpu, dfb, oracle, mode):
    # fusion_node applies Frameset::filterColor to every fused frame (fusion_node.cpp:789): the batched form must equal
    # splat -> filter per frame; 70 framesets cross two chunk boundaries of the fuser's scratch
    cfg = dfb.make_config(424, 240, hole_fraction=0.05)
    nfs = 70
    depth, color = oracle.synth_pool(cfg, 3, nfs)
    fuser = dfb.Fuser(cfg.desc)
    plain = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
    fuser.set_post_filter(mode)
    out = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
    for f in (0, 1, 31, 32, 63, 64, 69):
        assert np.array_equal(out[f], oracle.filter_color(plain[f], median=(mode == 2))), f"frameset {f}"
    # and through the host entry point
    h_out = np.zeros_like(out)
    fuser.fuse_host(depth.ctypes.data, color.ctypes.data, h_out.ctypes.data, nfs)
    assert np.array_equal(h_out, out)
    fuser.set_post_filter(0)
    assert np.array_equal(fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width)), plain)
    fuser.close()


def test_fuse_host_equals_fuse_device(gpu, dfb, oracle):
    torch = gpu
    cfg = dfb.make_config(848, 480)
    n = 7
    depth, color = oracle.synth_pool(cfg, 40, n)
    fuser = dfb.Fuser(cfg.desc)
    fuser.set_group(2)
    dev = fuse_pool(torch, fuser, depth, color, (cfg.out_height, cfg.out_width))
    hd = torch.from_numpy(depth.view(np.int16)).pin_memory()
    hc = torch.from_numpy(color).pin_memory()
    ho = torch.empty((n, cfg.out_height, cfg.out_width, 3), dtype=torch.uint8).pin_memory()
    fuser.fuse_host(hd.data_ptr(), hc.data_ptr(), ho.data_ptr(), n)
    assert np.array_equal(ho.numpy(), dev)
    # pageable host memory works too (slower)
    out2 = np.zeros_like(dev)
    fuser.fuse_host(depth.ctypes.data, color.ctypes.data, out2.ctypes.data, n)
    assert np.array_equal(out2, dev)


def test_odd_sizes_scalar_path(gpu, dfb, oracle):
    rng = np.random.default_rng(9)
    for (w, h) in [(37, 23), (130, 3), (1, 1), (5, 64)]:
        T = random_rigid(rng, 5, 0.05).astype(np.float32)
        F = random_rigid(rng, 5, 0.05).astype(np.float32)
        desc = simple_desc(w, h, n_cameras=2, fx=30.0, mirror=True, cam_T=[None, T], fused=F)
        depth = rng.integers(0, 3500, size=(3, 2, h, w), dtype=np.uint16)
        color = rng.integers(0, 256, size=(3, 2, h, w, 3), dtype=np.uint8)
        fuser = dfb.Fuser(desc)
        out = fuse_pool(gpu, fuser, depth, color, (h, w))
        for f in range(3):
            can = oracle.fuse_frameset(desc, list(depth[f]), list(color[f]), mode=1)
            assert np.array_equal(out[f], can["image"]), (w, h, f)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("DFB_RANDOM_PLANS", "10"))))  # more seeds: a fuzzing run
def test_random_plans_match_oracle(gpu, dfb, oracle, seed):
    # differential test over the plan space: rotations up to 180 degrees (points behind the virtual camera, upside-down
    # views, both traversals of the stream kernel), non-integer principal points, unequal focal lengths, ROIs, depth
    # ranges, scales, a fused image of another size than the cameras, 1-3 cameras — every variant against the oracle
    from helpers import make_intrinsics
    rng = np.random.default_rng(1000 + seed)
    w, h = [(256, 192), (320, 128), (192, 256), (448, 64)][seed % 4]
    nc = 1 + seed % 3
    ow, oh = [(w, h), (h, w), (w // 2 + 32, h + 16)][seed % 3]
    out = make_intrinsics(ow, oh, float(rng.uniform(0.3, 1.2) * ow), float(rng.uniform(0.3, 1.2) * ow),
                          float(ow / 2 + rng.uniform(-20, 20)), float(oh / 2 + rng.uniform(-20, 20)))
    big = seed % 2 == 1
    cam_T = [None if (c == 0 and seed % 4 < 2) else random_rigid(rng, 180.0 if big else 25.0, 0.6).astype(np.float32) for c in range(nc)]
    fused = random_rigid(rng, 180.0 if big else 40.0, 0.5).astype(np.float32)
    roi = [(-1, -1, -1, -1) if rng.random() < 0.5 else (int(rng.integers(0, w // 2)), int(rng.integers(0, h // 2)),
                                                      int(rng.integers(1, w)), int(rng.integers(1, h))) for _ in range(nc)]
    lo = float(rng.uniform(0.1, 1.0))
    desc = simple_desc(w, h, n_cameras=nc, mirror=bool(seed & 1), fx=float(rng.uniform(0.4, 1.5) * w), out=out, min_depth=lo,
                       max_depth=lo + float(rng.uniform(0.5, 4.0)), scale=float(rng.choice([0.001, 0.00025, 0.0123])),
                       fused=fused, cam_T=cam_T, roi=roi, ppx=float(w / 2 + rng.uniform(-30, 30)), ppy=float(h / 2 + rng.uniform(-30, 30)))
    nfs = 5
    depth = rng.integers(0, 5000, size=(nfs, nc, h, w), dtype=np.uint16)
    depth[rng.random(depth.shape) < 0.15] = 0
    smooth = (1500 + 800 * np.sin(np.arange(w) / 23.0)[None, :] + 300 * np.cos(np.arange(h) / 11.0)[:, None]).astype(np.uint16)
    depth[::2] = smooth  # every other frameset: a smooth surface (many neighbouring points per output pixel)
    color = rng.integers(0, 256, size=(nfs, nc, h, w, 3), dtype=np.uint8)
    fuser = dfb.Fuser(desc)
    outs = []
    for variant in (4, 1):
        fuser.set_variant(variant)
        outs.append(fuse_pool(gpu, fuser, depth, color, (oh, ow)))
    assert np.array_equal(outs[0], outs[1]), "stream kernel and direct kernel disagree"
    for f in range(nfs):
        can = oracle.fuse_frameset(desc, list(depth[f]), list(color[f]), mode=1)
        assert np.array_equal(outs[0][f], can["image"]), f"seed {seed} frameset {f}"
    fuser.close()


@pytest.mark.parametrize("kw", [dict(width=848, height=480), dict(width=424, height=240, n_cameras=3), dict(width=1280, height=720, vertical=False)])
def test_output_guard_bands_stay_untouched(gpu, dfb, oracle, kw):
    # no memcheck tool on this pool: the fused frames are written into the middle of a larger buffer whose guard bands
    # (before, after, and the 16-byte rounding slack of the last frame) must keep their pattern, for every variant
    torch = gpu
    cfg = dfb.make_config(**kw)
    nfs = 7
    depth, color = synth_pool_gpu(torch, dfb, cfg, 5, nfs)
    frame = cfg.out_height * cfg.out_width * 3
    guard = 1 << 16
    buf = torch.full((guard + nfs * frame + guard,), 0xA5, dtype=torch.uint8, device="cuda")
    out = buf[guard:guard + nfs * frame]
    fuser = dfb.Fuser(cfg.desc)
    ref = None
    for variant in (4, 3, 2, 1):
        buf.fill_(0xA5)
        fuser.set_variant(variant)
        fuser.fuse_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), nfs, stream_handle(torch))
        torch.cuda.synchronize()
        assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + nfs * frame:] == 0xA5).all()), f"variant {variant} wrote outside"
        img = out.clone()
        if ref is None:
            ref = img
        assert torch.equal(img, ref)
    fuser.close()


def test_empty_and_degenerate_inputs(gpu, dfb, oracle):
    cfg = dfb.make_config(848, 480)
    fuser = dfb.Fuser(cfg.desc)
    depth = np.zeros((1, 2, 480, 848), dtype=np.uint16)
    color = np.full((1, 2, 480, 848, 3), 200, dtype=np.uint8)
    out = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
    assert not out.any()  # empty mask -> black image (kernels.cu:582)
    # all pixels beyond max depth
    depth[:] = 3000
    assert not fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width)).any()
    # zero framesets is a no-op
    fuser.fuse_device(1, 1, 4, 0, 0)
    # heavy collisions: every point lands on a handful of pixels (tiny focal length)
    from helpers import make_intrinsics

    desc = simple_desc(256, 128, n_cameras=2, fx=200.0, cam_T=[None, np.eye(4, dtype=np.float32)],
                       out=make_intrinsics(256, 128, 0.5, 0.5, 128.0, 64.0))
    rng = np.random.default_rng(2)
    d = rng.integers(499, 3000, size=(1, 2, 128, 256), dtype=np.uint16)
    c = rng.integers(0, 256, size=(1, 2, 128, 256, 3), dtype=np.uint8)
    f2 = dfb.Fuser(desc)
    o = fuse_pool(gpu, f2, d, c, (128, 256))
    can = oracle.fuse_frameset(desc, list(d[0]), list(c[0]), mode=1)
    assert np.array_equal(o[0], can["image"]) and 0 < (o[0].any(axis=2)).sum() < 16


@pytest.mark.parametrize("kw", [dict(width=848, height=480, hole_fraction=0.1), dict(width=848, height=480, vertical=False, profile="P1"),
                                dict(width=424, height=240, n_cameras=3, roi=[(10, 20, 300, 200), (-1, -1, -1, -1), (0, 5, 9999, 100)])],
                         ids=["column", "row-major", "3cam-roi"])
@pytest.mark.parametrize("shrink", [6.0, 24.0])
def test_heavy_collisions_stay_exact(gpu, dfb, oracle, kw, shrink):
    # a fused image much coarser than the cameras: tens to hundreds of points per key, many of them from the lanes of ONE
    # warp instruction — the keys must still be the exact minima (collision study: profiles/r02_stream_ncu.md section 7)
    cfg = dfb.make_config(**kw)
    cfg.desc.out_intr.fx /= shrink
    cfg.desc.out_intr.fy /= shrink
    depth, color = oracle.synth_pool(cfg, 11, 3)
    fuser = dfb.Fuser(cfg.desc)
    out = fuse_pool(gpu, fuser, depth, color, (cfg.out_height, cfg.out_width))
    fuser.close()
    for f in range(3):
        can = oracle.fuse_frameset(cfg.desc, list(depth[f]), list(color[f]), mode=1, fast=True)
        assert np.array_equal(out[f], can["image"])


def test_transform_update(gpu, dfb, oracle):
    # registration publishes a new right->left transform at 0.5 Hz (fusion_node.cpp:610-613)
    a = dfb.make_config(848, 480, profile="P0")
    b = dfb.make_config(848, 480, profile="P1")
    depth, color = oracle.synth_pool(a, 3, 1)
    fuser = dfb.Fuser(a.desc)
    out_a = fuse_pool(gpu, fuser, depth, color, (a.out_height, a.out_width))
    fuser.update(b.desc)
    out_b = fuse_pool(gpu, fuser, depth, color, (a.out_height, a.out_width))
    assert np.array_equal(out_b[0], oracle.fuse_frameset(b.desc, list(depth[0]), list(color[0]), mode=1)["image"])
    assert not np.array_equal(out_a, out_b)


def test_full_batch_properties_c2(gpu, dfb, oracle):
    """BASELINE configs[1] at full batch size: device-generated pool, size-independent properties
    (batch == per-frameset runs; permutation of the batch permutes the output) + oracle on a sample."""
    torch = gpu
    cfg = dfb.make_config(1280, 720)
    n = 24
    depth, color = synth_pool_gpu(torch, dfb, cfg, 0, n)
    out = torch.empty((n, cfg.out_height, cfg.out_width, 3), dtype=torch.uint8, device="cuda")
    fuser = dfb.Fuser(cfg.desc)
    fuser.fuse_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), n, stream_handle(torch))
    torch.cuda.synchronize()
    # per-frameset runs give the same images
    single = torch.empty_like(out[0])
    for f in (0, 7, n - 1):
        fuser.fuse_device(depth[f].data_ptr(), color[f].data_ptr(), single.data_ptr(), 1, stream_handle(torch))
        torch.cuda.synchronize()
        assert torch.equal(single, out[f])
    # reversing the batch reverses the output
    rd, rc = depth.flip(0).contiguous(), color.flip(0).contiguous()
    out_r = torch.empty_like(out)
    fuser.fuse_device(rd.data_ptr(), rc.data_ptr(), out_r.data_ptr(), n, stream_handle(torch))
    torch.cuda.synchronize()
    assert torch.equal(out_r.flip(0), out)
    # oracle on two framesets of the device-generated pool
    for f in (3, 19):
        can = oracle.fuse_frameset(cfg.desc, list(u16(depth[f])), list(color[f].cpu().numpy()), mode=1, fast=True)
        assert np.array_equal(out[f].cpu().numpy(), can["image"])
    cov = (out.any(dim=3)).float().mean().item()
    assert 0.85 < cov < 0.95  # ~0.91 coverage for P0 dense


def test_c5_4k_against_oracle(gpu, dfb, oracle):
    torch = gpu
    cfg = dfb.make_config(3840, 2160)
    depth, color = synth_pool_gpu(torch, dfb, cfg, 0, 1)
    out = torch.empty((1, cfg.out_height, cfg.out_width, 3), dtype=torch.uint8, device="cuda")
    fuser = dfb.Fuser(cfg.desc)
    fuser.fuse_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), 1, stream_handle(torch))
    torch.cuda.synchronize()
    can = oracle.fuse_frameset(cfg.desc, list(u16(depth[0])), list(color[0].cpu().numpy()), mode=1, fast=True)
    assert np.array_equal(out[0].cpu().numpy(), can["image"])
