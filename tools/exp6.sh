#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp6.jsonl; : > $O
timeout 600 python -m pytest tests/test_shim_gpu.py -x -q -m gpu > gpurun_out/exp6_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/exp6_pytest.log
# reproduce with the debug binary + backtrace
python - <<'PY' > gpurun_out/exp6_shimdbg.log 2>&1
import os, struct, subprocess, sys
sys.path.insert(0, "pointcloud-depthfusion_b200"); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import numpy as np, dfb_b200, pyoracle
cfg = dfb_b200.make_config(848, 480, hole_fraction=0.05)
ds, cs = pyoracle.synth_frameset(cfg, 77)
d = cfg.desc; intr = d.cam[0].intr
hdr = struct.pack("<2i4f3f2i2i4f", intr.width, intr.height, intr.fx, intr.fy, intr.ppx, intr.ppy, d.cam[0].depth_scale, d.cam[0].min_depth, d.cam[0].max_depth, d.mirror, 0, d.out_intr.width, d.out_intr.height, d.out_intr.fx, d.out_intr.fy, d.out_intr.ppx, d.out_intr.ppy)
hdr += struct.pack("<4x") if len(hdr) % 8 else b""
hdr += np.asarray(cfg.cam_transforms[1], dtype=np.float64).tobytes() + np.asarray(cfg.fused_transform, dtype=np.float64).tobytes()
with open("/tmp/in.bin", "wb") as f:
    f.write(hdr)
    for c in range(2):
        f.write(ds[c].tobytes()); f.write(cs[c].tobytes())
r = subprocess.run(["pointcloud-depthfusion_b200/build/shim_fusion_test_dbg", "/tmp/in.bin", "/tmp/out.bin", "200"], capture_output=True, text=True)
print("rc", r.returncode); print(r.stdout[-3000:]); print(r.stderr[-6000:])
PY
tail -30 gpurun_out/exp6_shimdbg.log
for v in c1i c2i; do
DFB_B200_LIB=variants/$v/libdfb_b200.so DFB_STREAM_PROFILE=1 timeout 300 python tools/quick_bench.py --steps 1 --warmup 1 --framesets 128 --tag $v > gpurun_out/exp6_$v.log 2>&1
grep -E "cta +(37|74) compute" gpurun_out/exp6_$v.log | sort | head -4
done
