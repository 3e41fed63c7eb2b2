// tensor::Tensor of the kuiper:: API (method names and constructor forms as in reference
// kuiper/include/tensor/tensor.h): a dtype, up to N extents, and a shared base::Buffer.
//
// It is a VALUE type with shared storage: copying a Tensor shares the buffer (layers keep copies of
// the caller's tensors in their slots and write through them); clone() is the deep copy;
// to_cuda() / to_cpu() move the data and re-point THIS copy only.
//
// Construction rules the reference's tests pin down and this class keeps:
//   * (dtype, dim0, need_alloc=false, alloc) -- the 1-D form without need_alloc and without a
//     pointer stays EMPTY even when an allocator is supplied;
//   * the 2-D..4-D and vector forms allocate whenever an allocator is supplied;
//   * with `ptr` the tensor wraps memory the caller keeps alive; its device type is unknown until
//     set_device_type().
#ifndef KLLM_KUIPER_TENSOR_TENSOR_H_
#define KLLM_KUIPER_TENSOR_TENSOR_H_
#include <driver_types.h>
#include <glog/logging.h>

#include <memory>
#include <utility>
#include <vector>

#include "base/base.h"
#include "base/memory.h"

namespace tensor {
class Tensor {
  using AllocPtr = std::shared_ptr<base::DeviceAllocator>;

 public:
  Tensor() = default;
  explicit Tensor(base::DataType data_type, int32_t dim0, bool need_alloc = false, AllocPtr alloc = nullptr,
                  void* ptr = nullptr);
  explicit Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, bool need_alloc = false,
                  AllocPtr alloc = nullptr, void* ptr = nullptr);
  explicit Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, int32_t dim2, bool need_alloc = false,
                  AllocPtr alloc = nullptr, void* ptr = nullptr);
  explicit Tensor(base::DataType data_type, int32_t dim0, int32_t dim1, int32_t dim2, int32_t dim3,
                  bool need_alloc = false, AllocPtr alloc = nullptr, void* ptr = nullptr);
  explicit Tensor(base::DataType data_type, std::vector<int32_t> dims, bool need_alloc = false,
                  AllocPtr alloc = nullptr, void* ptr = nullptr);

  // ---- shape ----------------------------------------------------------------------------------------
  bool is_empty() const { return size_ == 0 || buffer_ == nullptr || buffer_->ptr() == nullptr; }
  size_t size() const { return size_; }  // elements
  size_t byte_size() const { return size_ * base::DataTypeSize(data_type_); }
  base::DataType data_type() const { return data_type_; }
  int32_t dims_size() const { return static_cast<int32_t>(dims_.size()); }
  const std::vector<int32_t>& dims() const { return dims_; }
  int32_t get_dim(int32_t idx) const;
  std::vector<size_t> strides() const;  // row-major, in elements
  // new extents; a larger element count reallocates and carries the old contents over
  void reshape(const std::vector<int32_t>& dims);
  // new dtype + extents, storage dropped
  void reset(base::DataType data_type, const std::vector<int32_t>& dims);

  // ---- storage --------------------------------------------------------------------------------------
  std::shared_ptr<base::Buffer> get_buffer() const { return buffer_; }
  // adopt `buffer` (must be large enough and, if this tensor has storage, on the same device)
  bool assign(std::shared_ptr<base::Buffer> buffer);
  bool allocate(AllocPtr allocator, bool need_realloc = false);
  void init_buffer(AllocPtr alloc, base::DataType data_type, bool need_alloc, void* ptr);
  Tensor clone() const;
  base::DeviceType device_type() const;
  void set_device_type(base::DeviceType device_type) const;
  void to_cuda(cudaStream_t stream = nullptr);
  void to_cpu();

  // ---- element access: ptr<T>() raw, ptr<T>(i) checked for storage, index<T>(i) bounds-checked ---------
  template <typename T>
  T* ptr() {
    return buffer_ ? static_cast<T*>(buffer_->ptr()) : nullptr;
  }
  template <typename T>
  const T* ptr() const {
    return buffer_ ? static_cast<const T*>(buffer_->ptr()) : nullptr;
  }
  template <typename T>
  T* ptr(int64_t index) {
    return const_cast<T*>(std::as_const(*this).template ptr<T>(index));
  }
  template <typename T>
  const T* ptr(int64_t index) const {
    CHECK(buffer_ != nullptr && buffer_->ptr() != nullptr)
        << "The data area buffer of this tensor is empty or it points to a null pointer.";
    return static_cast<const T*>(buffer_->ptr()) + index;
  }
  template <typename T>
  T& index(int64_t offset) {
    return const_cast<T&>(std::as_const(*this).template index<T>(offset));
  }
  template <typename T>
  const T& index(int64_t offset) const {  // host tensors (the demo writes the position through this)
    CHECK_GE(offset, 0);
    CHECK_LT(offset, static_cast<int64_t>(size_));
    return static_cast<const T*>(buffer_->ptr())[offset];
  }

 private:
  size_t size_ = 0;
  std::vector<int32_t> dims_;
  std::shared_ptr<base::Buffer> buffer_;
  base::DataType data_type_ = base::DataType::kDataTypeUnknown;
};
}  // namespace tensor
#endif  // KLLM_KUIPER_TENSOR_TENSOR_H_
