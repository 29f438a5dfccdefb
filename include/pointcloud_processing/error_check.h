#pragma once
/*
 * error_check.h — the reference's error convention (pointcloud_processing/src/error_check.cpp:10-51):
 * no return codes, no exceptions; any failure prints `Cuda error: <msg> failed: <reason>` to stderr and
 * exits the process. The C ABI underneath returns status codes; these helpers restore the drop-in behaviour.
 */
#include <cuda_runtime.h>

#include "dfb/dfb.h"

DFB_API void checkError(const char* msg, bool sync = false);
DFB_API void checkError(cudaError_t error, const char* msg, bool sync = false);
/* NPP is not a dependency of this implementation; the status is any integer with 0 = success */
DFB_API void checkNppError(int status, const char* msg, bool sync = false);
/* status code of a dfb_* call (reason taken from dfb_last_error()) */
DFB_API void checkDfbError(int status, const char* msg);
