// kuiper::base -- status codes, enums and logging glue of the KuiperLLama API surface.
// Interface mirrors zjhellofss/KuiperLLama kuiper/include/base/base.h:1-155 so code written against
// the reference (demo/main.cpp, the gtest files) compiles unchanged; the implementation is ours.
#ifndef KLLM_KUIPER_BASE_BASE_H_
#define KLLM_KUIPER_BASE_BASE_H_
#include <glog/logging.h>

#include <cstdint>
#include <cstdio>
#include <ostream>
#include <string>

#define UNUSED(expr) \
  do {               \
    (void)(expr);    \
  } while (0)

namespace model {
// Named activation / cache buffers of a model (reference base.h:12-32; the numbering is part of
// the API: demo code asks for kInputPos).
enum class ModelBufferType {
  kInputTokens = 0,
  kInputEmbeddings = 1,
  kOutputRMSNorm = 2,
  kKeyCache = 3,
  kValueCache = 4,
  kQuery = 5,
  kInputPos = 6,
  kScoreStorage = 7,
  kOutputMHA = 8,
  kAttnOutput = 9,
  kW1Output = 10,
  kW2Output = 11,
  kW3Output = 12,
  kFFNRMSNorm = 13,
  kForwardOutput = 15,
  kForwardOutputCPU = 16,
  kSinCache = 17,
  kCosCache = 18,
};
}  // namespace model

namespace base {

enum class DeviceType : uint8_t { kDeviceUnknown = 0, kDeviceCPU = 1, kDeviceCUDA = 2 };

enum class DataType : uint8_t {
  kDataTypeUnknown = 0,
  kDataTypeFp32 = 1,
  kDataTypeInt8 = 2,
  kDataTypeInt32 = 3,
};

enum class ModelType : uint8_t { kModelTypeUnknown = 0, kModelTypeLLama2 = 1 };

enum class TokenizerType { kEncodeUnknown = -1, kEncodeSpe = 0, kEncodeBpe = 1 };

enum StatusCode : uint8_t {
  kSuccess = 0,
  kFunctionUnImplement = 1,
  kPathNotValid = 2,
  kModelParseError = 3,
  kInternalError = 5,
  kKeyValueHasExist = 6,
  kInvalidArgument = 7,
};

inline size_t DataTypeSize(DataType t) {
  switch (t) {
    case DataType::kDataTypeFp32: return sizeof(float);
    case DataType::kDataTypeInt8: return sizeof(int8_t);
    case DataType::kDataTypeInt32: return sizeof(int32_t);
    default: return 0;
  }
}

class NoCopyable {
 protected:
  NoCopyable() = default;
  ~NoCopyable() = default;
  NoCopyable(const NoCopyable&) = delete;
  NoCopyable& operator=(const NoCopyable&) = delete;
};

// Value-type result: converts to bool (true = success) and to int (the code).
class Status {
 public:
  Status(int code = StatusCode::kSuccess, std::string err_message = "");
  Status(const Status& other) = default;
  Status& operator=(const Status& other) = default;
  Status& operator=(int code);
  bool operator==(int code) const;
  bool operator!=(int code) const;
  operator int() const;
  operator bool() const;
  int32_t get_err_code() const;
  const std::string& get_err_msg() const;
  void set_err_msg(const std::string& err_msg);

 private:
  int code_ = StatusCode::kSuccess;
  std::string message_;
};

namespace error {
// A failed Status is fatal to the caller, exactly as in the reference (base.h:123-134).
#define STATUS_CHECK(call)                                                                    \
  do {                                                                                        \
    const base::Status& kllm_status_ = (call);                                                \
    if (!kllm_status_) {                                                                      \
      char kllm_buf_[512];                                                                    \
      snprintf(kllm_buf_, sizeof(kllm_buf_) - 1,                                              \
               "Infer error\n File:%s Line:%d\n Error code:%d\n Error msg:%s\n", __FILE__,    \
               __LINE__, int(kllm_status_), kllm_status_.get_err_msg().c_str());              \
      LOG(FATAL) << kllm_buf_;                                                                \
    }                                                                                         \
  } while (0)

Status Success(const std::string& err_msg = "");
Status FunctionNotImplement(const std::string& err_msg = "");
Status PathNotValid(const std::string& err_msg = "");
Status ModelParseError(const std::string& err_msg = "");
Status InternalError(const std::string& err_msg = "");
Status KeyHasExits(const std::string& err_msg = "");
Status InvalidArgument(const std::string& err_msg = "");
}  // namespace error

std::ostream& operator<<(std::ostream& os, const Status& x);

}  // namespace base
#endif  // KLLM_KUIPER_BASE_BASE_H_
