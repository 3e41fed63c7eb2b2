// Common vocabulary of the kuiper:: API surface: device / dtype / tokenizer enums, the named
// model buffers, and base::Status with its error factories.  Enumerator names and values are part
// of the interface the reference's demos and tests compile against (kuiper/include/base/base.h);
// everything here is header-only.
#ifndef KLLM_KUIPER_BASE_BASE_H_
#define KLLM_KUIPER_BASE_BASE_H_
#include <glog/logging.h>

#include <cstdint>
#include <cstdio>
#include <ostream>
#include <string>
#include <utility>

#define UNUSED(expr) static_cast<void>(expr)

namespace base {
enum class DeviceType : uint8_t { kDeviceUnknown = 0, kDeviceCPU = 1, kDeviceCUDA = 2 };
enum class DataType : uint8_t { kDataTypeUnknown = 0, kDataTypeFp32 = 1, kDataTypeInt8 = 2, kDataTypeInt32 = 3 };
enum class ModelType : uint8_t { kModelTypeUnknown = 0, kModelTypeLLama2 = 1 };
enum class TokenizerType { kEncodeUnknown = -1, kEncodeSpe = 0, kEncodeBpe = 1 };

// bytes of one element
inline size_t DataTypeSize(DataType t) {
  return t == DataType::kDataTypeFp32 || t == DataType::kDataTypeInt32 ? 4 : t == DataType::kDataTypeInt8 ? 1 : 0;
}

// ---- Status -------------------------------------------------------------------------------------------
// A small value type: an error code plus a message.  Converts to bool (true = success) and to int
// (the code), compares against codes.  Layer and model methods return it; STATUS_CHECK turns a
// failure into LOG(FATAL), which is how the reference treats every failed call.
enum StatusCode : uint8_t {
  kSuccess = 0, kFunctionUnImplement = 1, kPathNotValid = 2, kModelParseError = 3, kInternalError = 5,
  kKeyValueHasExist = 6, kInvalidArgument = 7,
};

class Status {
 public:
  Status(int code = kSuccess, std::string err_message = "") : code_(code), message_(std::move(err_message)) {}
  Status& operator=(int code) {
    code_ = code;
    return *this;
  }
  operator bool() const { return code_ == kSuccess; }
  operator int() const { return code_; }
  bool operator==(int code) const { return code_ == code; }
  bool operator!=(int code) const { return code_ != code; }
  int32_t get_err_code() const { return code_; }
  const std::string& get_err_msg() const { return message_; }
  void set_err_msg(const std::string& err_msg) { message_ = err_msg; }

 private:
  int code_;
  std::string message_;
};
inline std::ostream& operator<<(std::ostream& os, const Status& s) { return os << s.get_err_msg(); }

namespace error {
#define KLLM_STATUS_FACTORY(name, code) \
  inline Status name(const std::string& err_msg = "") { return Status(code, err_msg); }
KLLM_STATUS_FACTORY(Success, kSuccess)
KLLM_STATUS_FACTORY(FunctionNotImplement, kFunctionUnImplement)
KLLM_STATUS_FACTORY(PathNotValid, kPathNotValid)
KLLM_STATUS_FACTORY(ModelParseError, kModelParseError)
KLLM_STATUS_FACTORY(InternalError, kInternalError)
KLLM_STATUS_FACTORY(KeyHasExits, kKeyValueHasExist)
KLLM_STATUS_FACTORY(InvalidArgument, kInvalidArgument)
#undef KLLM_STATUS_FACTORY
}  // namespace error

// Mixin for types that own a resource.
class NoCopyable {
 protected:
  NoCopyable() = default;
  ~NoCopyable() = default;
  NoCopyable(const NoCopyable&) = delete;
  NoCopyable& operator=(const NoCopyable&) = delete;
};
}  // namespace base

// A failed Status is fatal where it is checked (file, line, code and message go to the log).
#define STATUS_CHECK(call)                                                                                      \
  do {                                                                                                          \
    const base::Status kllm_status_ = (call);                                                                   \
    if (!kllm_status_)                                                                                          \
      LOG(FATAL) << "Infer error\n File:" << __FILE__ << " Line:" << __LINE__ << "\n Error code:"               \
                 << int(kllm_status_) << "\n Error msg:" << kllm_status_.get_err_msg() << "\n";                 \
  } while (0)

namespace model {
// The named activation / cache tensors a model keeps (Model::get_buffer).  The numbering is API:
// demo code asks for kInputPos, tests for kForwardOutput.
enum class ModelBufferType {
  kInputTokens = 0, kInputEmbeddings = 1, kOutputRMSNorm = 2, kKeyCache = 3, kValueCache = 4, kQuery = 5,
  kInputPos = 6, kScoreStorage = 7, kOutputMHA = 8, kAttnOutput = 9, kW1Output = 10, kW2Output = 11,
  kW3Output = 12, kFFNRMSNorm = 13, kForwardOutput = 15, kForwardOutputCPU = 16, kSinCache = 17, kCosCache = 18,
};
}  // namespace model
#endif  // KLLM_KUIPER_BASE_BASE_H_
