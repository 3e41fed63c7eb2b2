#ifndef KLLM_KUIPER_MODEL_RAW_MODEL_DATA_H_
#define KLLM_KUIPER_MODEL_RAW_MODEL_DATA_H_
#include <cstddef>
#include <cstdint>
namespace model {
// The mmap'd checkpoint.  weight(offset) addresses the payload after the header in ELEMENTS of
// the file's weight type: floats for fp32 files, bytes for int8 files (reference
// raw_model_data.h:6-23).
struct RawModelData {
  virtual ~RawModelData();
  int32_t fd = -1;
  size_t file_size = 0;
  void* data = nullptr;
  void* weight_data = nullptr;
  virtual const void* weight(size_t offset) const = 0;
};
struct RawModelDataFp32 : RawModelData {
  const void* weight(size_t offset) const override;
};
struct RawModelDataInt8 : RawModelData {
  const void* weight(size_t offset) const override;
};
}  // namespace model
#endif
