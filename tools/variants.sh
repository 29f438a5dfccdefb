#!/bin/bash
# Builds experiment variants of libdfb_b200.so into variants/<tag>/libdfb_b200.so (git-ignored, travels with gpurun):
#   tools/variants.sh tag "-DFOO=1 -DBAR=2" [tag2 "flags2" ...]
# Select one at run time with DFB_B200_LIB=variants/<tag>/libdfb_b200.so.
set -e
REPO=$(cd "$(dirname "$0")/.." && pwd)
PKG=$REPO/pointcloud-depthfusion_b200
unset CC CXX
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  out=$REPO/variants/$tag
  mkdir -p $out/build
  make -s -C $PKG -j8 BUILD=$out/build LIB=$out/libdfb_b200.so DFB_EXTRA_NVFLAGS="$flags" >/dev/null
  echo "built $out/libdfb_b200.so [$flags]"
done
