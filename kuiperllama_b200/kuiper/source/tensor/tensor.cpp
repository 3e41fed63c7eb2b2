#include "tensor/tensor.h"

#include <cuda_runtime_api.h>

#include <functional>
#include <numeric>

namespace tensor {
namespace {
size_t element_count(const std::vector<int32_t>& dims) {
  if (dims.empty()) return 0;
  size_t n = 1;
  for (int32_t d : dims) n *= static_cast<size_t>(d);
  return n;
}
}  // namespace

Tensor::Tensor(base::DataType data_type, int32_t dim0, bool need_alloc,
               std::shared_ptr<base::DeviceAllocator> alloc, void* ptr)
    : size_(dim0), dims_{dim0}, data_type_(data_type) {
  if (need_alloc && alloc) {
    allocate(alloc);
  } else if (ptr != nullptr) {
    CHECK(!need_alloc) << "The need_alloc is is true when ptr parameter is not a null pointer.";
    init_buffer(alloc, data_type_, need_alloc, ptr);
  }
  // else: stays empty (pinned by the reference's test_tensor.cpp:115-123)
}

Tensor::Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, bool need_alloc,
               std::shared_ptr<base::DeviceAllocator> alloc, void* ptr)
    : Tensor(data_type, std::vector<int32_t>{dim0, dim1}, need_alloc, std::move(alloc), ptr) {}

Tensor::Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, int32_t dim2, bool need_alloc,
               std::shared_ptr<base::DeviceAllocator> alloc, void* ptr)
    : Tensor(data_type, std::vector<int32_t>{dim0, dim1, dim2}, need_alloc, std::move(alloc), ptr) {}

Tensor::Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, int32_t dim2, int32_t dim3,
               bool need_alloc, std::shared_ptr<base::DeviceAllocator> alloc, void* ptr)
    : Tensor(data_type, std::vector<int32_t>{dim0, dim1, dim2, dim3}, need_alloc, std::move(alloc), ptr) {}

Tensor::Tensor(base::DataType data_type, std::vector<int32_t> dims, bool need_alloc,
               std::shared_ptr<base::DeviceAllocator> alloc, void* ptr)
    : dims_(std::move(dims)), data_type_(data_type) {
  size_ = element_count(dims_);
  if (need_alloc && alloc) {
    allocate(alloc);
  } else {
    init_buffer(alloc, data_type_, need_alloc, ptr);
  }
}

void Tensor::init_buffer(std::shared_ptr<base::DeviceAllocator> alloc, base::DataType data_type,
                         bool need_alloc, void* ptr) {
  if (!alloc && !need_alloc) {
    // view on memory owned elsewhere (mmap'd weights, a KV-cache row, ...)
    buffer_ = std::make_shared<base::Buffer>(base::DataTypeSize(data_type) * size_, nullptr, ptr, true);
  } else {
    allocate(alloc, true);
  }
}

void Tensor::to_cuda(cudaStream_t stream) {
  CHECK_NE(buffer_, nullptr);
  const base::DeviceType dev = device_type();
  if (dev == base::DeviceType::kDeviceUnknown) {
    LOG(ERROR) << "The device type of the tensor is unknown.";
  } else if (dev == base::DeviceType::kDeviceCPU) {
    auto cu_alloc = base::CUDADeviceAllocatorFactory::get_instance();
    auto cu_buffer = std::make_shared<base::Buffer>(byte_size(), cu_alloc);
    cu_alloc->memcpy(buffer_->ptr(), cu_buffer->ptr(), byte_size(), base::MemcpyKind::kMemcpyCPU2CUDA, stream);
    buffer_ = cu_buffer;
  } else {
    LOG(INFO) << "The device type of the tensor is already cuda.";
  }
}

void Tensor::to_cpu() {
  CHECK_NE(buffer_, nullptr);
  const base::DeviceType dev = device_type();
  if (dev == base::DeviceType::kDeviceUnknown) {
    LOG(ERROR) << "The device type of the tensor is unknown.";
  } else if (dev == base::DeviceType::kDeviceCUDA) {
    auto cpu_alloc = base::CPUDeviceAllocatorFactory::get_instance();
    auto cpu_buffer = std::make_shared<base::Buffer>(byte_size(), cpu_alloc);
    cpu_alloc->memcpy(buffer_->ptr(), cpu_buffer->ptr(), byte_size(), base::MemcpyKind::kMemcpyCUDA2CPU);
    buffer_ = cpu_buffer;
  } else {
    LOG(INFO) << "The device type of the tensor is already cpu.";
  }
}

bool Tensor::is_empty() const { return size_ == 0 || buffer_ == nullptr || buffer_->ptr() == nullptr; }
size_t Tensor::size() const { return size_; }
size_t Tensor::byte_size() const { return size_ * base::DataTypeSize(data_type_); }
int32_t Tensor::dims_size() const { return static_cast<int32_t>(dims_.size()); }
base::DataType Tensor::data_type() const { return data_type_; }
const std::vector<int32_t>& Tensor::dims() const { return dims_; }
std::shared_ptr<base::Buffer> Tensor::get_buffer() const { return buffer_; }

int32_t Tensor::get_dim(int32_t idx) const {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(dims_.size()));
  return dims_[idx];
}

base::DeviceType Tensor::device_type() const {
  return buffer_ ? buffer_->device_type() : base::DeviceType::kDeviceUnknown;
}

void Tensor::set_device_type(base::DeviceType device_type) const {
  if (buffer_) buffer_->set_device_type(device_type);
}

bool Tensor::assign(std::shared_ptr<base::Buffer> buffer) {
  if (!buffer) {
    LOG(ERROR) << "The buffer parameter in the assign function is null pointer!";
    return false;
  }
  if (buffer_ && buffer_->device_type() != buffer->device_type()) {
    LOG(ERROR) << "The device type of the new buffer is different from the original one.";
  }
  if (byte_size() > buffer->byte_size()) {
    LOG(ERROR) << "The size of buffer is too small for the tensor!";
    return false;
  }
  buffer_ = std::move(buffer);
  return true;
}

bool Tensor::allocate(std::shared_ptr<base::DeviceAllocator> allocator, bool need_realloc) {
  if (!allocator) {
    LOG(ERROR) << "The allocator parameter in the allocate function is null pointer!";
    return false;
  }
  const size_t bytes = byte_size();
  if (bytes == 0) {
    LOG(ERROR) << "The byte_size parameter in the allocate function is equal to zero!";
    return false;
  }
  if (buffer_ && bytes <= buffer_->byte_size() && !need_realloc) return true;
  buffer_ = std::make_shared<base::Buffer>(bytes, allocator, nullptr);
  if (buffer_->ptr() == nullptr) {
    LOG(ERROR) << "The memory allocated is a null pointer!";
    return false;
  }
  return true;
}

void Tensor::reset(base::DataType data_type, const std::vector<int32_t>& dims) {
  data_type_ = data_type;
  dims_ = dims;
  size_ = element_count(dims);
  buffer_ = nullptr;
}

void Tensor::reshape(const std::vector<int32_t>& dims) {
  const size_t n = element_count(dims);
  if (buffer_ && n > size_) {
    // growing: new buffer from the same allocator, old contents preserved
    auto grown = std::make_shared<base::Buffer>(n * base::DataTypeSize(data_type_), buffer_->allocator());
    CHECK(grown->ptr() != nullptr || grown->allocate());
    grown->copy_from(buffer_.get());
    buffer_ = grown;
  }
  dims_ = dims;
  size_ = n;
}

Tensor Tensor::clone() const {
  Tensor copy = *this;
  copy.buffer_ = std::make_shared<base::Buffer>(byte_size(), buffer_->allocator());
  copy.buffer_->copy_from(buffer_.get());
  return copy;
}

std::vector<size_t> Tensor::strides() const {
  std::vector<size_t> s(dims_.size(), 1);
  for (int i = static_cast<int>(dims_.size()) - 2; i >= 0; --i) s[i] = s[i + 1] * static_cast<size_t>(dims_[i + 1]);
  return s;
}
}  // namespace tensor
