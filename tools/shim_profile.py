#!/usr/bin/env python
"""Writes the C2 input file of the drop-in C++ consumer (tests/cpp/shim_fusion_test.cpp) to gpurun_out/shim_in.bin and runs
the consumer's bench mode once; tools/shim_profile.sh then lists the kernels of a few frames under ncu."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "pointcloud-depthfusion_b200"))
import torch  # noqa: E402
import dfb_b200  # noqa: E402
from dfb_b200 import capi  # noqa: E402
import bench  # noqa: E402

dfb_b200.init(0)
L = dfb_b200.lib()
cfg = dfb_b200.make_config(1280, 720)
sh = torch.cuda.current_stream().cuda_stream
depth = torch.empty((1, 2, 720, 1280), dtype=torch.int16, device="cuda")
color = torch.empty((1, 2, 720, 1280, 3), dtype=torch.uint8, device="cuda")
capi.check(L.dfb_synth_generate(cfg.scenes(0, 1), 1, cfg.cams_array(), 2, depth.data_ptr(), color.data_ptr(), sh))
torch.cuda.synchronize()
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
print(json.dumps(bench.shim_leg(cfg, depth, color, torch, keep_input=os.path.join(REPO, "gpurun_out", "shim_in.bin"), frames=2000)))
