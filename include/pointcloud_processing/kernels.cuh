#pragma once
/*
 * kernels.cuh — what callers of the reference use from `namespace Kernels`
 * (pointcloud_processing/include/pointcloud_processing/kernels.cuh:12-92): only init_cuda
 * (fusion_node/src/fusion_node.cpp:204, kernels.cu:325-351). The per-operator wrappers of the reference are
 * replaced by the C ABI (include/dfb/dfb.h: dfb_op_*), which Frameset / Pointcloud call.
 */
#include <cuda_runtime.h>

#include "dfb/dfb.h"

namespace Kernels {
extern DFB_API bool cuda_initialized;
/* device 0, as the reference hard-codes it */
DFB_API void init_cuda(bool verbose = false);
/* extension for multi-GPU hosts: one process (or thread) per device */
DFB_API void init_cuda(int device, bool verbose);
}  // namespace Kernels
