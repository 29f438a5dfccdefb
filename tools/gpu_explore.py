#!/usr/bin/env python
"""One-off GPU measurements that steer the kernel design (run under gpurun; writes gpurun_out/explore.json):
  - L2 64-bit atomicMin throughput by address pattern, L2-resident vs not, RED vs ATOM
  - per-kernel times of the fused path for each traversal variant
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointcloud-depthfusion_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import dfb_b200  # noqa: E402

os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
dfb_b200.init(0, verbose=True)
L = dfb_b200.lib()
res = {"atomics": [], "variants": []}
n_ops = 1 << 27
PATTERNS = [(0, 0, "contiguous"), (3, 0, "warp-runs-scattered"), (4, 0, "pair-duplicated"), (1, 720, "lane-stride-720"),
            (1, 4, "lane-stride-4"), (2, 0, "random"), (0x100, 0, "contiguous-ATOM"), (0x102, 0, "random-ATOM")]
for slots_mb in (8, 64, 512):
    n_slots = slots_mb * (1 << 20) // 8
    for pattern, stride, name in PATTERNS:
        ms = L.dfb_probe_atomics(n_slots, n_ops, pattern, stride, 5)
        rate = n_ops * 5 / (ms * 1e-3)
        res["atomics"].append(dict(slots_mb=slots_mb, pattern=name, ms=ms, gops=rate / 1e9))
        print(f"atomicMin.u64 {name:20s} keys={slots_mb:4d} MiB: {rate / 1e9:8.2f} Gop/s", flush=True)

import torch  # noqa: E402
from dfb_b200 import capi  # noqa: E402

for wl, F in (("C2", 64), ("C1", 128), ("C5", 8)):
    kw = dict(C1=dict(width=848, height=480), C2=dict(width=1280, height=720), C5=dict(width=3840, height=2160))[wl]
    cfg = dfb_b200.make_config(**kw)
    depth = torch.empty((F, 2, cfg.height, cfg.width), dtype=torch.int16, device="cuda")
    color = torch.empty((F, 2, cfg.height, cfg.width, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((F, cfg.out_height, cfg.out_width, 3), dtype=torch.uint8, device="cuda")
    sh = torch.cuda.current_stream().cuda_stream
    capi.check(L.dfb_synth_generate(cfg.scenes(0, F), F, cfg.cams_array(), 2, depth.data_ptr(), color.data_ptr(), sh))
    for variant in (3, 4):
        for group in ((0, 8) if variant == 3 else (0,)):
            fuser = dfb_b200.Fuser(cfg.desc)
            fuser.set_variant(variant)
            fuser.set_group(group)
            for _ in range(3):
                p = fuser.profile_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
            fs_per_s = F / ((p["splat_ms"] + p["resolve_ms"]) * 1e-3)
            row = dict(workload=wl, variant=variant, group=fuser.group, splat_ms=p["splat_ms"], resolve_ms=p["resolve_ms"],
                       framesets_per_s=fs_per_s, hbm_frac=fs_per_s * cfg.algorithmic_bytes / 6550.7e9)
            res["variants"].append(row)
            print(row, flush=True)
            fuser.close()

with open(os.path.join(REPO, "gpurun_out", "explore.json"), "w") as f:
    json.dump(res, f, indent=1)
