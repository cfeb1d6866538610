// AsTensor / Shape / TensorMap — the subset of csrc/core/tensor/{tensor,shape}.h the hot-path operators touch.
#pragma once
#include <cuda_runtime.h>

#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"

namespace allspark {

class DeviceContext;

class Shape {
 public:
  Shape() = default;
  Shape(std::initializer_list<int64_t> d) : dims_(d) {}
  explicit Shape(const std::vector<int64_t>& d) : dims_(d) {}
  int Size() const { return (int)dims_.size(); }
  int64_t operator[](int i) const { return dims_[i < 0 ? i + (int)dims_.size() : i]; }
  void Append(int64_t d) { dims_.push_back(d); }
  int64_t Count() const { return Count(0, Size()); }
  int64_t Count(int start, int end) const {
    int64_t c = 1;
    for (int i = start; i < end; ++i) c *= dims_[i];
    return c;
  }
  bool operator==(const Shape& o) const { return dims_ == o.dims_; }
  const std::vector<int64_t>& dims() const { return dims_; }

 private:
  std::vector<int64_t> dims_;
};

// Dense tensor that owns its storage (device memory through the CUDA runtime, host memory on the heap).
// SetShape grows the allocation in place like DenseData::Resize (csrc/core/tensor/tensor.cpp:721-746): the shared
// "workspace" tensor relies on that.
class AsTensor {
 public:
  explicit AsTensor(const std::string& name = "", DeviceType backend = DeviceType::CPU, DataType dtype = DATATYPE_UNDEFINED,
                    DataMode mode = DataMode::DENSE, const Shape& shape = {})
      : name_(name), backend_(backend), dtype_(dtype), mode_(mode) {
    SetShape(Shape(shape));
  }
  ~AsTensor() { Release(); }
  AsTensor(const AsTensor&) = delete;
  AsTensor& operator=(const AsTensor&) = delete;

  const std::string& GetName() const { return name_; }
  const Shape& GetShape() const { return shape_; }
  DataType GetDataType() const { return dtype_; }
  DeviceType GetDeviceType() const { return backend_; }
  DataMode GetDataMode() const { return mode_; }
  void* GetDataPtr() const { return data_; }
  size_t GetSizeInByte() const { return (size_t)shape_.Count() * SizeofType(dtype_); }

  AsStatus SetDataType(DataType t) {
    dtype_ = t;
    return Reserve(GetSizeInByte());
  }
  AsStatus SetShape(Shape&& s) {
    shape_ = std::move(s);
    return Reserve(GetSizeInByte());
  }
  void CopyDataFrom(const void* src, size_t bytes, DeviceType src_dev, const DeviceContext* = nullptr) {
    if (!bytes) return;
    Reserve(bytes);
    if (backend_ == DeviceType::CUDA)
      cudaMemcpy(data_, src, bytes, src_dev == DeviceType::CUDA ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
    else if (src_dev == DeviceType::CUDA)
      cudaMemcpy(data_, src, bytes, cudaMemcpyDeviceToHost);
    else
      std::memcpy(data_, src, bytes);
  }
  void CopyDataTo(void* dst, size_t bytes, DeviceType dst_dev, const DeviceContext* = nullptr) const {
    if (!bytes) return;
    if (backend_ == DeviceType::CUDA)
      cudaMemcpy(dst, data_, bytes, dst_dev == DeviceType::CUDA ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost);
    else if (dst_dev == DeviceType::CUDA)
      cudaMemcpy(dst, data_, bytes, cudaMemcpyHostToDevice);
    else
      std::memcpy(dst, data_, bytes);
  }

 private:
  AsStatus Reserve(size_t bytes) {
    if (bytes <= capacity_) return AsStatus::ALLSPARK_SUCCESS;
    Release();
    if (backend_ == DeviceType::CUDA) {
      if (cudaMalloc(&data_, bytes) != cudaSuccess) return AsStatus::ALLSPARK_MEMORY_ERROR;
    } else {
      data_ = std::malloc(bytes);
      if (!data_) return AsStatus::ALLSPARK_MEMORY_ERROR;
    }
    capacity_ = bytes;
    return AsStatus::ALLSPARK_SUCCESS;
  }
  void Release() {
    if (data_) {
      if (backend_ == DeviceType::CUDA) cudaFree(data_); else std::free(data_);
    }
    data_ = nullptr;
    capacity_ = 0;
  }
  std::string name_;
  DeviceType backend_;
  DataType dtype_;
  DataMode mode_;
  Shape shape_;
  void* data_ = nullptr;
  size_t capacity_ = 0;
};

using TensorMap = std::unordered_map<std::string, std::shared_ptr<AsTensor>>;

}  // namespace allspark
