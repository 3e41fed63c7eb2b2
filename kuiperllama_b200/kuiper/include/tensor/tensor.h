// tensor::Tensor -- dims + dtype + a shared Buffer (reference kuiper/include/tensor/tensor.h:12-95).
// Value type: copies share the buffer; to_cuda()/to_cpu() swap the buffer of THIS copy only.
#ifndef KLLM_KUIPER_TENSOR_TENSOR_H_
#define KLLM_KUIPER_TENSOR_TENSOR_H_
#include <driver_types.h>
#include <glog/logging.h>

#include <memory>
#include <vector>

#include "base/base.h"
#include "base/buffer.h"
namespace tensor {

class Tensor {
 public:
  explicit Tensor() = default;
  // Constructor quirks the reference's tests pin: the 1-D form with
  // need_alloc=false and no ptr stays EMPTY even when given an allocator; the N-D forms allocate
  // whenever an allocator is supplied.  With `ptr` the tensor wraps external memory whose device
  // type is unknown until set_device_type().
  explicit Tensor(base::DataType data_type, int32_t dim0, bool need_alloc = false,
                  std::shared_ptr<base::DeviceAllocator> alloc = nullptr, void* ptr = nullptr);
  explicit Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, bool need_alloc = false,
                  std::shared_ptr<base::DeviceAllocator> alloc = nullptr, void* ptr = nullptr);
  explicit Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, int32_t dim2,
                  bool need_alloc = false, std::shared_ptr<base::DeviceAllocator> alloc = nullptr,
                  void* ptr = nullptr);
  explicit Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, int32_t dim2, int32_t dim3,
                  bool need_alloc = false, std::shared_ptr<base::DeviceAllocator> alloc = nullptr,
                  void* ptr = nullptr);
  explicit Tensor(base::DataType data_type, std::vector<int32_t> dims, bool need_alloc = false,
                  std::shared_ptr<base::DeviceAllocator> alloc = nullptr, void* ptr = nullptr);

  void to_cpu();
  void to_cuda(cudaStream_t stream = nullptr);
  bool is_empty() const;
  void init_buffer(std::shared_ptr<base::DeviceAllocator> alloc, base::DataType data_type,
                   bool need_alloc, void* ptr);
  void reshape(const std::vector<int32_t>& dims);
  std::shared_ptr<base::Buffer> get_buffer() const;
  size_t size() const;
  size_t byte_size() const;
  int32_t dims_size() const;
  base::DataType data_type() const;
  int32_t get_dim(int32_t idx) const;
  const std::vector<int32_t>& dims() const;
  std::vector<size_t> strides() const;
  bool assign(std::shared_ptr<base::Buffer> buffer);
  void reset(base::DataType data_type, const std::vector<int32_t>& dims);
  void set_device_type(base::DeviceType device_type) const;
  base::DeviceType device_type() const;
  bool allocate(std::shared_ptr<base::DeviceAllocator> allocator, bool need_realloc = false);
  tensor::Tensor clone() const;

  template <typename T>
  T* ptr() {
    return buffer_ ? reinterpret_cast<T*>(buffer_->ptr()) : nullptr;
  }
  template <typename T>
  const T* ptr() const {
    return buffer_ ? reinterpret_cast<const T*>(buffer_->ptr()) : nullptr;
  }
  template <typename T>
  T* ptr(int64_t index) {
    CHECK(buffer_ != nullptr && buffer_->ptr() != nullptr)
        << "The data area buffer of this tensor is empty or it points to a null pointer.";
    return reinterpret_cast<T*>(buffer_->ptr()) + index;
  }
  template <typename T>
  const T* ptr(int64_t index) const {
    CHECK(buffer_ != nullptr && buffer_->ptr() != nullptr)
        << "The data area buffer of this tensor is empty or it points to a null pointer.";
    return reinterpret_cast<const T*>(buffer_->ptr()) + index;
  }
  // Host-side element access (CPU tensors; the demo writes the position through this).
  template <typename T>
  T& index(int64_t offset) {
    CHECK_GE(offset, 0);
    CHECK_LT(offset, static_cast<int64_t>(this->size()));
    return *(reinterpret_cast<T*>(buffer_->ptr()) + offset);
  }
  template <typename T>
  const T& index(int64_t offset) const {
    CHECK_GE(offset, 0);
    CHECK_LT(offset, static_cast<int64_t>(this->size()));
    return *(reinterpret_cast<const T*>(buffer_->ptr()) + offset);
  }

 private:
  size_t size_ = 0;
  std::vector<int32_t> dims_;
  std::shared_ptr<base::Buffer> buffer_;
  base::DataType data_type_ = base::DataType::kDataTypeUnknown;
};
}  // namespace tensor
#endif  // KLLM_KUIPER_TENSOR_TENSOR_H_
