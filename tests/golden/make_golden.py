"""Generates tests/golden/ref_b200.npz: outputs of the reference's OWN kernels (unmodified
pointcloud_processing/src/kernels.cu compiled in place for sm_100a, oracle/_ref/libref_kernels.so) on small
seeded inputs. The reference ships no golden vectors; these fixtures are the committed
"outputs of the reference itself" pin that tests/test_golden_ref.py replays on CPU against the oracle.

Run on a B200 (needs the GPU; /root/reference is NOT needed at run time, only the prebuilt oracle/_ref):
    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_b200.npz'
then copy gpurun_out/ref_b200.npz to tests/golden/. Inputs are stored next to the outputs so the CPU test
does not depend on the generator staying bit-stable.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REPO, "pointcloud-depthfusion_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")]


def main(out_path):
    import dfb_b200
    import pyoracle
    import ref_kernels
    from helpers import align_case, make_intrinsics, random_rigid

    dfb_b200.load()
    dfb_b200.init(0)
    fix = {}
    rng = np.random.default_rng(20261018)

    # 1. kernel_transform (kernels.cu:161-181): the FMA nest
    pts = rng.uniform(-4, 4, size=(4096, 4)).astype(np.float32)
    pts[:256] *= 1e-3
    T = random_rigid(rng, 180, 1.0).astype(np.float32).reshape(16)
    fix["transform_in"], fix["transform_T"] = pts, T
    fix["transform_out"] = ref_kernels.transform(pts, T)

    # 2. the whole operator sequence of FusionNode::processSyncedFrames (fusion_node.cpp:718-794) on small framesets
    cases = {
        "seq_a": dict(width=192, height=96, hole_fraction=0.05),                      # 18 x 1024 px: no unwritten tail
        "seq_b": dict(width=160, height=96, profile="P1", hole_fraction=0.1),         # 15 x 1024 px, upside-down camera
        "seq_c": dict(width=192, height=96, density="sparse", vertical=False, mirror=False, roi=[(8, 4, 150, 80), (-1, -1, -1, -1)]),
        "seq_d": dict(width=100, height=60),                                          # 6000 px: reference leaves a tail unwritten (App. A.6 i)
    }
    for name, kw in cases.items():
        cfg = dfb_b200.make_config(**kw)
        ds, cs = pyoracle.synth_frameset(cfg, 5)
        ref = ref_kernels.fuse_frameset(cfg.desc, ds, cs)
        fix[name + "_desc"] = np.frombuffer(bytes(cfg.desc), dtype=np.uint8).copy()
        for c in range(cfg.n_cameras):
            fix[f"{name}_depth{c}"], fix[f"{name}_color{c}"] = ds[c], cs[c]
            fix[f"{name}_filtered{c}"], fix[f"{name}_mask{c}"] = ref["depth_filtered"][c], ref["mask"][c]
            fix[f"{name}_indices{c}"] = ref["indices"][c]
        fix[name + "_counts"] = np.asarray(ref["counts"], dtype=np.uint32)
        fix[name + "_cloud"], fix[name + "_zbuf"], fix[name + "_image"] = ref["cloud"], ref["zbuf"], ref["image"]

    # 3. kernel_project_pointcloud (kernels.cu:239-274) on a collision-free cloud: race-free, so the image is pinned too
    w, h = 96, 64
    intr = make_intrinsics(w, h, 50.0, 50.0, 48.0, 32.0)
    v, u = np.mgrid[0:h, 0:w]
    z = rng.uniform(0.6, 2.5, size=u.shape).astype(np.float32)
    cloud = np.zeros((u.size, 4), np.float32)
    cloud[:, 0] = (z * ((u.astype(np.float32) - 48) / np.float32(50))).reshape(-1)
    cloud[:, 1] = (z * ((v.astype(np.float32) - 32) / np.float32(50))).reshape(-1)
    cloud[:, 2] = z.reshape(-1)
    cloud[:, 3] = rng.integers(0, 1 << 24, size=u.size, dtype=np.uint32).view(np.float32)
    fix["project_cloud"] = cloud
    fix["project_intr"] = np.frombuffer(bytes(intr), dtype=np.uint8).copy()
    for mirror in (0, 1):
        img, zb = ref_kernels.project(cloud, intr, bool(mirror))
        fix[f"project_image_m{mirror}"], fix[f"project_zbuf_m{mirror}"] = img, zb

    # 4. NPP 12.x post filters as the reference calls them (kernels.cu:594-653)
    img = rng.integers(0, 256, size=(48, 64, 3), dtype=np.uint8)
    img[rng.random((48, 64)) < 0.3] = 0
    fix["filter_in"] = img
    fix["filter_gauss"] = ref_kernels.filter_color(img, median=False)
    fix["filter_median"] = ref_kernels.filter_color(img, median=True)

    # 5. Kernels::align_depth_to_color (kernels.cu:714-747): unaligned depth -> colour frame, with and without distortion
    for name, kw in {"align_a": dict(depth_model=4, color_model=2, coeff_scale=0.05),
                     "align_b": dict(depth_model=2, color_model=1, coeff_scale=0.1)}.items():
        depth, di, ci, ex, scale = align_case(np.random.default_rng(77), dw=192, dh=96, cw=256, ch=128, **kw)
        fix[name + "_depth"] = depth
        fix[name + "_depth_intr"] = np.frombuffer(bytes(di), dtype=np.uint8).copy()
        fix[name + "_color_intr"] = np.frombuffer(bytes(ci), dtype=np.uint8).copy()
        fix[name + "_extr"] = np.frombuffer(bytes(ex), dtype=np.uint8).copy()
        fix[name + "_scale"] = np.float32(scale)
        fix[name + "_out"] = ref_kernels.align_depth(depth, di, ci, ex, float(scale))

    np.savez_compressed(out_path, **fix)
    print(f"wrote {out_path}: {len(fix)} arrays, {os.path.getsize(out_path)} bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_b200.npz"))
