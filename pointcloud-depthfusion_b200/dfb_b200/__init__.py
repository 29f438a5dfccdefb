"""dfb_b200 — ctypes binding of libdfb_b200.so (the C ABI in include/dfb/dfb.h).

This is harness glue for tests/ and bench.py, not the product: the product is the CUDA library
and the C++ drop-in headers in include/pointcloud_processing/. Nothing here computes anything;
there is no CPU fallback — if the library (or a B200) is missing, calls raise.
"""
from .capi import (  # noqa: F401
    DfbError,
    Fuser,
    Workspace,
    init,
    last_error,
    launch_count,
    lib,
    lib_path,
    load,
)
from .configs import (  # noqa: F401
    CONFIGS,
    FusionConfig,
    make_config,
)
from .ctypes_defs import (  # noqa: F401
    CameraDesc,
    DfsCamera,
    DfsScene,
    Extrinsics,
    FuseDesc,
    Intrinsics,
    MAX_CAMERAS,
)
