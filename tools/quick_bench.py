#!/usr/bin/env python
"""Device-resident framesets/s of one library build / environment (kernel experiments; not the driver's bench).
    DFB_B200_LIB=variants/x/libdfb_b200.so python tools/quick_bench.py [--workload C2] [--framesets 512] [--steps 8] [--check]
Prints one JSON line. --check compares a few framesets against the CPU oracle."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointcloud-depthfusion_b200"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--framesets", type=int, default=0)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--profile", default="P0")
ap.add_argument("--density", default="dense")
ap.add_argument("--horizontal", action="store_true")
ap.add_argument("--check", type=int, default=0, help="framesets to compare with the oracle (spread over the batch)")
ap.add_argument("--out-shrink", type=float, default=1.0,
                help="divide the fused image's focal lengths: K*K times as many points per covered pixel (collision study)")
ap.add_argument("--tag", default="")
args = ap.parse_args()

import torch  # noqa: E402
import dfb_b200  # noqa: E402
from dfb_b200 import capi  # noqa: E402

WL = {"C1": (848, 480, 2, 1024), "C2": (1280, 720, 2, 512), "C4": (1280, 720, 4, 256), "C5": (3840, 2160, 2, 32)}
w, h, ncam, fdef = WL[args.workload]
F = args.framesets or fdef
dfb_b200.init(0)
L = dfb_b200.lib()
cfg = dfb_b200.make_config(w, h, n_cameras=ncam, profile=args.profile, density=args.density, vertical=not args.horizontal)
if args.out_shrink != 1.0:
    cfg.desc.out_intr.fx /= args.out_shrink
    cfg.desc.out_intr.fy /= args.out_shrink
fuser = dfb_b200.Fuser(cfg.desc)
stream = torch.cuda.current_stream()
sh = stream.cuda_stream
depth = torch.empty((F, ncam, h, w), dtype=torch.int16, device="cuda")
color = torch.empty((F, ncam, h, w, 3), dtype=torch.uint8, device="cuda")
capi.check(L.dfb_synth_generate(cfg.scenes(0, F), F, cfg.cams_array(), ncam, depth.data_ptr(), color.data_ptr(), sh))
out = torch.zeros((F, cfg.out_height, cfg.out_width, 3), dtype=torch.uint8, device="cuda")
for _ in range(args.warmup):
    fuser.fuse_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(args.steps):
    fuser.fuse_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
res = {"tag": args.tag, "lib": os.environ.get("DFB_B200_LIB", "default"), "workload": args.workload, "framesets": F,
       "ms_per_step": ms, "framesets_per_s": F / (ms * 1e-3), "frac": F / (ms * 1e-3) * cfg.algorithmic_bytes / 6550.7e9,
       "out_shrink": args.out_shrink,
       "covered_fraction": float(out[0].any(dim=2).float().mean().item()),
       "env": {k: v for k, v in os.environ.items() if k.startswith("DFB_")}}
if args.check:
    import numpy as np
    import pyoracle

    idx = sorted(set(int(round(i * (F - 1) / max(args.check - 1, 1))) for i in range(args.check)))
    bad = 0
    for f in idx:
        d, c = pyoracle.synth_pool(cfg, f, 1)
        exp = pyoracle.fuse_frameset(cfg.desc, list(d[0]), list(c[0]), mode=1)["image"]
        got = out[f].cpu().numpy()
        bad += int((got != exp).any(axis=2).sum())
    res["check"] = {"framesets": idx, "mismatched_pixels": bad}
print(json.dumps(res), flush=True)
