// oracle/_ref build stub (TEST INFRASTRUCTURE): stands in for the reference's csrc/utility/check_cuda.h, which drags in
// glog / cublas / nccl / AsException that are not in this image.  Only the one macro the cache kernels use.
#pragma once
#include <cuda_runtime_api.h>

#include <stdexcept>
#include <string>

#define AS_CHECK_CUDA_LAST_ERROR()                                                       \
  do {                                                                                   \
    cudaError_t err_ = cudaGetLastError();                                               \
    if (err_ != cudaSuccess) throw std::runtime_error(std::string("[Cuda error]") + cudaGetErrorString(err_)); \
  } while (0)
#define AS_CHECK_CUDA(cmd)                                                               \
  do {                                                                                   \
    cudaError_t err_ = (cmd);                                                            \
    if (err_ != cudaSuccess) throw std::runtime_error(std::string("[Cuda error]") + cudaGetErrorString(err_)); \
  } while (0)
