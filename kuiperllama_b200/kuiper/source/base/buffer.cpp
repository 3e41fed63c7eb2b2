#include "base/memory.h"

#include <glog/logging.h>

#include <algorithm>

namespace base {
Buffer::Buffer(size_t byte_size, std::shared_ptr<DeviceAllocator> allocator, void* ptr,
               bool use_external)
    : byte_size_(byte_size), ptr_(ptr), use_external_(use_external), allocator_(std::move(allocator)) {
  if (ptr_ == nullptr && allocator_ != nullptr) {
    device_type_ = allocator_->device_type();
    use_external_ = false;
    ptr_ = allocator_->allocate(byte_size_);
  }
}

Buffer::~Buffer() {
  if (!use_external_ && ptr_ != nullptr && allocator_ != nullptr) {
    allocator_->release(ptr_);
    ptr_ = nullptr;
  }
}

bool Buffer::allocate() {
  if (allocator_ == nullptr || byte_size_ == 0) return false;
  use_external_ = false;
  ptr_ = allocator_->allocate(byte_size_);
  return ptr_ != nullptr;
}

namespace {
MemcpyKind kind_for(DeviceType from, DeviceType to) {
  if (from == DeviceType::kDeviceCPU && to == DeviceType::kDeviceCPU) return MemcpyKind::kMemcpyCPU2CPU;
  if (from == DeviceType::kDeviceCUDA && to == DeviceType::kDeviceCPU) return MemcpyKind::kMemcpyCUDA2CPU;
  if (from == DeviceType::kDeviceCPU && to == DeviceType::kDeviceCUDA) return MemcpyKind::kMemcpyCPU2CUDA;
  return MemcpyKind::kMemcpyCUDA2CUDA;
}
}  // namespace

void Buffer::copy_from(const Buffer& buffer) const { copy_from(&buffer); }

void Buffer::copy_from(const Buffer* buffer) const {
  CHECK(allocator_ != nullptr);
  CHECK(buffer != nullptr && buffer->ptr_ != nullptr);
  CHECK(buffer->device_type_ != DeviceType::kDeviceUnknown && device_type_ != DeviceType::kDeviceUnknown);
  const size_t n = std::min(byte_size_, buffer->byte_size_);
  allocator_->memcpy(buffer->ptr_, ptr_, n, kind_for(buffer->device_type_, device_type_));
}

}  // namespace base
