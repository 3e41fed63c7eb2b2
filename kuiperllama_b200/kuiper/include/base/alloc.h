// Device allocators (reference kuiper/include/base/alloc.h:1-93).
//
// Same interface; different CUDA allocator: the reference scans two vectors linearly on every
// allocate/release and its argmax leaks one cudaMalloc per token (argmax_kernel.cu:74-76).  Here
// the pooled allocator keeps size-bucketed free lists (O(log n)) and nothing on the per-token
// path allocates at all.
#ifndef KLLM_KUIPER_BASE_ALLOC_H_
#define KLLM_KUIPER_BASE_ALLOC_H_
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "base.h"
namespace base {
enum class MemcpyKind {
  kMemcpyCPU2CPU = 0,
  kMemcpyCPU2CUDA = 1,
  kMemcpyCUDA2CPU = 2,
  kMemcpyCUDA2CUDA = 3,
};

class DeviceAllocator {
 public:
  explicit DeviceAllocator(DeviceType device_type) : device_type_(device_type) {}
  virtual ~DeviceAllocator() = default;
  virtual DeviceType device_type() const { return device_type_; }
  virtual void release(void* ptr) const = 0;
  virtual void* allocate(size_t byte_size) const = 0;
  virtual void memcpy(const void* src_ptr, void* dest_ptr, size_t byte_size,
                      MemcpyKind memcpy_kind = MemcpyKind::kMemcpyCPU2CPU, void* stream = nullptr,
                      bool need_sync = false) const;
  virtual void memset_zero(void* ptr, size_t byte_size, void* stream, bool need_sync = false);

 private:
  DeviceType device_type_ = DeviceType::kDeviceUnknown;
};

class CPUDeviceAllocator : public DeviceAllocator {
 public:
  explicit CPUDeviceAllocator();
  void* allocate(size_t byte_size) const override;
  void release(void* ptr) const override;
};

class CUDADeviceAllocator : public DeviceAllocator {
 public:
  explicit CUDADeviceAllocator();
  ~CUDADeviceAllocator() override;
  void* allocate(size_t byte_size) const override;
  void release(void* ptr) const override;
  // bytes currently parked in the free lists (all devices)
  size_t cached_bytes() const;

 private:
  struct Block {
    int device;
    size_t bytes;
  };
  mutable std::mutex mu_;
  mutable std::unordered_map<void*, Block> live_;                          // handed out
  mutable std::map<std::pair<int, size_t>, std::vector<void*>> free_;      // (device, size) -> blocks
  mutable size_t cached_ = 0;
};

class CPUDeviceAllocatorFactory {
 public:
  static std::shared_ptr<CPUDeviceAllocator> get_instance() {
    if (instance == nullptr) instance = std::make_shared<CPUDeviceAllocator>();
    return instance;
  }

 private:
  static std::shared_ptr<CPUDeviceAllocator> instance;
};

class CUDADeviceAllocatorFactory {
 public:
  static std::shared_ptr<CUDADeviceAllocator> get_instance() {
    if (instance == nullptr) instance = std::make_shared<CUDADeviceAllocator>();
    return instance;
  }

 private:
  static std::shared_ptr<CUDADeviceAllocator> instance;
};
}  // namespace base
#endif  // KLLM_KUIPER_BASE_ALLOC_H_
