#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp12.jsonl; : > $O
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp12.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), r['env'])"; }
for v in c2s3 c2s4 c2s3dyn c2s3ra0 c2s3late c2s2 c2s3asm; do run $v DFB_B200_LIB=variants/$v/libdfb_b200.so; done
for sl in "6 3" "7 3" "7 4" "8 4" "9 4" "10 5"; do set -- $sl; run c2s3_s$1_l$2 DFB_B200_LIB=variants/c2s3/libdfb_b200.so DFB_STREAM_SLOTS=$1 DFB_STREAM_LAG=$2; done
Q="timeout 300 python tools/quick_bench.py --steps 5 --workload C1"; run C1 DFB_B200_LIB=variants/c2s3/libdfb_b200.so
Q="timeout 300 python tools/quick_bench.py --steps 5 --workload C4"; run C4 DFB_B200_LIB=variants/c2s3/libdfb_b200.so
Q="timeout 300 python tools/quick_bench.py --steps 5 --workload C5"; run C5 DFB_B200_LIB=variants/c2s3/libdfb_b200.so
Q="timeout 300 python tools/quick_bench.py --steps 5 --horizontal"; run C2h DFB_B200_LIB=variants/c2s3/libdfb_b200.so
Q="timeout 300 python tools/quick_bench.py --steps 5 --profile P1"; run C2P1 DFB_B200_LIB=variants/c2s3/libdfb_b200.so
# shim leg with the deferred filter count
python - <<'PY'
import os, struct, subprocess, sys
sys.path.insert(0, "pointcloud-depthfusion_b200"); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import numpy as np, dfb_b200, pyoracle
cfg = dfb_b200.make_config(1280, 720)
ds, cs = pyoracle.synth_frameset(cfg, 77)
d = cfg.desc; intr = d.cam[0].intr
hdr = struct.pack("<2i4f3f2i2i4f", intr.width, intr.height, intr.fx, intr.fy, intr.ppx, intr.ppy, d.cam[0].depth_scale, d.cam[0].min_depth, d.cam[0].max_depth, d.mirror, 0, d.out_intr.width, d.out_intr.height, d.out_intr.fx, d.out_intr.fy, d.out_intr.ppx, d.out_intr.ppy)
hdr += struct.pack("<4x") if len(hdr) % 8 else b""
hdr += np.asarray(cfg.cam_transforms[1], dtype=np.float64).tobytes() + np.asarray(cfg.fused_transform, dtype=np.float64).tobytes()
with open("/tmp/in.bin", "wb") as f:
    f.write(hdr)
    for c in range(2):
        f.write(ds[c].tobytes()); f.write(cs[c].tobytes())
r = subprocess.run(["pointcloud-depthfusion_b200/build/shim_fusion_test", "/tmp/in.bin", "/tmp/out.bin", "2000"], capture_output=True, text=True)
print("shim rc", r.returncode, r.stdout[-300:], r.stderr[-300:])
PY
timeout 600 python -m pytest tests/test_shim_gpu.py tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -3
