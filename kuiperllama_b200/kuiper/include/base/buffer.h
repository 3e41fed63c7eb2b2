// base::Buffer -- a byte range plus who frees it (reference kuiper/include/base/buffer.h:6-43).
#ifndef KLLM_KUIPER_BASE_BUFFER_H_
#define KLLM_KUIPER_BASE_BUFFER_H_
#include <memory>

#include "base/alloc.h"
namespace base {
class Buffer : public NoCopyable, std::enable_shared_from_this<Buffer> {
 public:
  explicit Buffer() = default;
  // ptr == nullptr and an allocator: allocate now, the buffer owns the memory.
  // ptr != nullptr: wrap it; `use_external` says the buffer must not free it.
  explicit Buffer(size_t byte_size, std::shared_ptr<DeviceAllocator> allocator = nullptr,
                  void* ptr = nullptr, bool use_external = false);
  virtual ~Buffer();

  bool allocate();
  void copy_from(const Buffer& buffer) const;
  void copy_from(const Buffer* buffer) const;
  void* ptr();
  const void* ptr() const;
  size_t byte_size() const;
  std::shared_ptr<DeviceAllocator> allocator() const;
  DeviceType device_type() const;
  void set_device_type(DeviceType device_type);
  std::shared_ptr<Buffer> get_shared_from_this();
  bool is_external() const;

 private:
  size_t byte_size_ = 0;
  void* ptr_ = nullptr;
  bool use_external_ = false;
  DeviceType device_type_ = DeviceType::kDeviceUnknown;
  std::shared_ptr<DeviceAllocator> allocator_;
};
}  // namespace base
#endif  // KLLM_KUIPER_BASE_BUFFER_H_
