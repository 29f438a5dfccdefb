#!/usr/bin/env python
"""Dump NPP gauss/median (through the unmodified reference wrappers) next to our kernel and the oracle on the
same random image -> gpurun_out/npp_dump.npz (used once to pin the post-filter semantics offline)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointcloud-depthfusion_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(REPO, p))
import torch  # noqa: E402
import dfb_b200  # noqa: E402
import pyoracle  # noqa: E402
import ref_kernels  # noqa: E402
from dfb_b200 import capi  # noqa: E402

dfb_b200.init(0)
rng = np.random.default_rng(5)
img = rng.integers(0, 256, size=(96, 128, 3), dtype=np.uint8)
img[rng.random((96, 128)) < 0.3] = 0
out = {"img": img}
for median in (0, 1):
    out[f"npp{median}"] = ref_kernels.filter_color(img, median=bool(median))
    out[f"oracle{median}"] = pyoracle.filter_color(img, median=bool(median))
    d = torch.from_numpy(img.copy()).cuda()
    s = torch.empty_like(d)
    capi.check(dfb_b200.lib().dfb_op_filter_color(d.data_ptr(), s.data_ptr(), 128, 96, median, 0))
    torch.cuda.synchronize()
    out[f"ours{median}"] = d.cpu().numpy()
    for a, b in (("npp", "oracle"), ("ours", "oracle")):
        diff = (out[f"{a}{median}"] != out[f"{b}{median}"])
        print(f"median={median} {a} vs {b}: {diff.sum()} differing bytes")
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(REPO, "gpurun_out", "npp_dump.npz"), **out)
