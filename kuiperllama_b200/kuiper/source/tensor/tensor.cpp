// tensor::Tensor (see tensor/tensor.h for the semantics this file implements).
#include "tensor/tensor.h"

#include <cuda_runtime_api.h>

namespace tensor {
namespace {
size_t count_of(const std::vector<int32_t>& dims) {
  size_t n = dims.empty() ? 0 : 1;
  for (int32_t d : dims) n *= static_cast<size_t>(d);
  return n;
}
const char* name_of(base::DeviceType d) {
  return d == base::DeviceType::kDeviceCUDA ? "CUDA" : d == base::DeviceType::kDeviceCPU ? "CPU" : "unknown";
}
}  // namespace

// The 1-D form is the odd one out: without need_alloc and without a pointer it stays empty even if
// an allocator is passed (the reference's test_tensor.cpp `init2` depends on it).
Tensor::Tensor(base::DataType data_type, int32_t dim0, bool need_alloc, AllocPtr alloc, void* ptr)
    : size_(static_cast<size_t>(dim0)), dims_{dim0}, data_type_(data_type) {
  if (need_alloc && alloc) {
    allocate(std::move(alloc));
  } else if (ptr != nullptr) {
    CHECK(!need_alloc) << "a tensor that wraps `ptr` cannot also be asked to allocate";
    init_buffer(std::move(alloc), data_type_, false, ptr);
  }
}
Tensor::Tensor(base::DataType data_type, std::vector<int32_t> dims, bool need_alloc, AllocPtr alloc, void* ptr)
    : size_(count_of(dims)), dims_(std::move(dims)), data_type_(data_type) {
  if (need_alloc && alloc) allocate(std::move(alloc));
  else init_buffer(std::move(alloc), data_type_, need_alloc, ptr);
}
Tensor::Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, bool need_alloc, AllocPtr alloc, void* ptr)
    : Tensor(data_type, std::vector<int32_t>{dim0, dim1}, need_alloc, std::move(alloc), ptr) {}
Tensor::Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, int32_t dim2, bool need_alloc, AllocPtr alloc,
               void* ptr)
    : Tensor(data_type, std::vector<int32_t>{dim0, dim1, dim2}, need_alloc, std::move(alloc), ptr) {}
Tensor::Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, int32_t dim2, int32_t dim3, bool need_alloc,
               AllocPtr alloc, void* ptr)
    : Tensor(data_type, std::vector<int32_t>{dim0, dim1, dim2, dim3}, need_alloc, std::move(alloc), ptr) {}

void Tensor::init_buffer(AllocPtr alloc, base::DataType data_type, bool need_alloc, void* ptr) {
  if (alloc || need_alloc) {
    allocate(std::move(alloc), true);
    return;
  }
  // a view on memory owned elsewhere: mmap'd weights, a KV-cache row, an embedding row
  buffer_ = std::make_shared<base::Buffer>(base::DataTypeSize(data_type) * size_, nullptr, ptr, true);
}

int32_t Tensor::get_dim(int32_t idx) const {
  CHECK(idx >= 0 && idx < dims_size()) << "dimension " << idx << " of a " << dims_size() << "-d tensor";
  return dims_[idx];
}

std::vector<size_t> Tensor::strides() const {
  std::vector<size_t> s(dims_.size(), 1);
  for (size_t i = dims_.size(); i-- > 1;) s[i - 1] = s[i] * static_cast<size_t>(dims_[i]);
  return s;
}

void Tensor::reset(base::DataType data_type, const std::vector<int32_t>& dims) {
  data_type_ = data_type;
  dims_ = dims;
  size_ = count_of(dims);
  buffer_.reset();
}

void Tensor::reshape(const std::vector<int32_t>& dims) {
  const size_t n = count_of(dims);
  if (buffer_ && n > size_) {  // grow: fresh storage from the same allocator, contents carried over
    auto grown = std::make_shared<base::Buffer>(n * base::DataTypeSize(data_type_), buffer_->allocator());
    CHECK(grown->ptr() != nullptr || grown->allocate());
    grown->copy_from(buffer_.get());
    buffer_ = std::move(grown);
  }
  dims_ = dims;
  size_ = n;
}

base::DeviceType Tensor::device_type() const {
  return buffer_ ? buffer_->device_type() : base::DeviceType::kDeviceUnknown;
}
void Tensor::set_device_type(base::DeviceType device_type) const {
  if (buffer_) buffer_->set_device_type(device_type);
}

bool Tensor::assign(std::shared_ptr<base::Buffer> buffer) {
  if (!buffer) {
    LOG(ERROR) << "Tensor::assign: no buffer given";
    return false;
  }
  if (buffer_ && buffer_->device_type() != buffer->device_type())
    LOG(ERROR) << "Tensor::assign: the new buffer lives on " << name_of(buffer->device_type()) << ", the old one on "
               << name_of(buffer_->device_type());
  if (byte_size() > buffer->byte_size()) {
    LOG(ERROR) << "Tensor::assign: " << buffer->byte_size() << " bytes cannot hold this tensor (" << byte_size() << ")";
    return false;
  }
  buffer_ = std::move(buffer);
  return true;
}

bool Tensor::allocate(AllocPtr allocator, bool need_realloc) {
  const size_t bytes = byte_size();
  if (!allocator || bytes == 0) {
    LOG(ERROR) << "Tensor::allocate: " << (allocator ? "the tensor has no elements" : "no allocator given");
    return false;
  }
  if (buffer_ && bytes <= buffer_->byte_size() && !need_realloc) return true;  // what is there is enough
  buffer_ = std::make_shared<base::Buffer>(bytes, std::move(allocator), nullptr);
  if (buffer_->ptr() == nullptr) {
    LOG(ERROR) << "Tensor::allocate: the allocator returned no memory for " << bytes << " bytes";
    return false;
  }
  return true;
}

Tensor Tensor::clone() const {
  Tensor copy = *this;
  copy.buffer_ = std::make_shared<base::Buffer>(byte_size(), buffer_->allocator());
  copy.buffer_->copy_from(buffer_.get());
  return copy;
}

void Tensor::to_cuda(cudaStream_t stream) {
  CHECK(buffer_ != nullptr) << "Tensor::to_cuda on a tensor without storage";
  switch (device_type()) {
    case base::DeviceType::kDeviceCPU: {
      auto device = base::CUDADeviceAllocatorFactory::get_instance();
      auto there = std::make_shared<base::Buffer>(byte_size(), device);
      device->memcpy(buffer_->ptr(), there->ptr(), byte_size(), base::MemcpyKind::kMemcpyCPU2CUDA, stream);
      buffer_ = std::move(there);
      break;
    }
    case base::DeviceType::kDeviceCUDA: LOG(INFO) << "Tensor::to_cuda: already on the device"; break;
    default: LOG(ERROR) << "Tensor::to_cuda: the tensor's device is unknown (wrapped pointer?)";
  }
}

void Tensor::to_cpu() {
  CHECK(buffer_ != nullptr) << "Tensor::to_cpu on a tensor without storage";
  switch (device_type()) {
    case base::DeviceType::kDeviceCUDA: {
      auto host = base::CPUDeviceAllocatorFactory::get_instance();
      auto here = std::make_shared<base::Buffer>(byte_size(), host);
      host->memcpy(buffer_->ptr(), here->ptr(), byte_size(), base::MemcpyKind::kMemcpyCUDA2CPU);
      buffer_ = std::move(here);
      break;
    }
    case base::DeviceType::kDeviceCPU: LOG(INFO) << "Tensor::to_cpu: already on the host"; break;
    default: LOG(ERROR) << "Tensor::to_cpu: the tensor's device is unknown (wrapped pointer?)";
  }
}
}  // namespace tensor
