#!/bin/bash
timeout 900 python -m pytest tests/test_ref_kernels_gpu.py tests/test_ops_gpu.py tests/test_shim_gpu.py -x -q -m gpu 2>&1 | tail -15
