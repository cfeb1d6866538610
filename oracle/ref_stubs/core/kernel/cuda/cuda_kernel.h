// oracle/_ref build stub (TEST INFRASTRUCTURE): the reference's cuda_kernel.h declares every CUDA launcher of the engine
// and pulls cuda_common.h (cublas, protobuf enums ...).  The cache kernels only need the CUDA runtime types from it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
