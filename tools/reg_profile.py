#!/usr/bin/env python
"""Registration feed (BASELINE configs[2]) in isolation: dfb_registration_cloud + both voxel grids on one 848x480 pair.
    python tools/reg_profile.py [--reps 50]        host wall clock per stage
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv python tools/reg_profile.py --reps 1
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointcloud-depthfusion_b200"))
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=50)
args = ap.parse_args()
import torch  # noqa: E402
import dfb_b200  # noqa: E402
from dfb_b200 import capi  # noqa: E402

dfb_b200.init(0)
L = dfb_b200.lib()
sh = torch.cuda.current_stream().cuda_stream
cfg = dfb_b200.make_config(848, 480, hole_fraction=0.1)
rd = torch.empty((1, 2, 480, 848), dtype=torch.int16, device="cuda")
rc = torch.empty((1, 2, 480, 848, 3), dtype=torch.uint8, device="cuda")
capi.check(L.dfb_synth_generate(cfg.scenes(0, 1), 1, cfg.cams_array(), 2, rd.data_ptr(), rc.data_ptr(), sh))
ws = dfb_b200.Workspace()
pts = torch.empty((848 * 480, 4), dtype=torch.float32, device="cuda")
vox = torch.empty((848 * 480, 4), dtype=torch.float32, device="cuda")
cnt, nv = C.c_uint32(), C.c_uint32()


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def cloud():
    capi.check(L.dfb_registration_cloud(ws.handle, C.byref(cfg.desc.cam[0]), rd[0, 0].data_ptr(), pts.data_ptr(), C.byref(cnt), sh))


res = {"cloud_us": timed(cloud, args.reps), "points": cnt.value}
for name, fn in (("voxel_exact", L.dfb_voxel_downsample), ("voxel_pcl_approximate", L.dfb_voxel_downsample_approximate)):
    l0 = L.dfb_launch_count()
    res[name + "_us"] = timed(lambda: capi.check(fn(ws.handle, pts.data_ptr(), cnt.value, 0.01, vox.data_ptr(), C.byref(nv), sh)), args.reps)
    res[name + "_launches"] = (L.dfb_launch_count() - l0) // (args.reps + 2)
    res[name + "_out"] = nv.value
print(json.dumps(res))
