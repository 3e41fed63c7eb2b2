#include "base/base.h"

#include <utility>

namespace base {
Status::Status(int code, std::string err_message) : code_(code), message_(std::move(err_message)) {}

Status& Status::operator=(int code) {
  code_ = code;
  return *this;
}
bool Status::operator==(int code) const { return code_ == code; }
bool Status::operator!=(int code) const { return code_ != code; }
Status::operator int() const { return code_; }
Status::operator bool() const { return code_ == kSuccess; }
int32_t Status::get_err_code() const { return code_; }
const std::string& Status::get_err_msg() const { return message_; }
void Status::set_err_msg(const std::string& err_msg) { message_ = err_msg; }

namespace error {
Status Success(const std::string& m) { return Status(kSuccess, m); }
Status FunctionNotImplement(const std::string& m) { return Status(kFunctionUnImplement, m); }
Status PathNotValid(const std::string& m) { return Status(kPathNotValid, m); }
Status ModelParseError(const std::string& m) { return Status(kModelParseError, m); }
Status InternalError(const std::string& m) { return Status(kInternalError, m); }
Status KeyHasExits(const std::string& m) { return Status(kKeyValueHasExist, m); }
Status InvalidArgument(const std::string& m) { return Status(kInvalidArgument, m); }
}  // namespace error

std::ostream& operator<<(std::ostream& os, const Status& x) { return os << x.get_err_msg(); }
}  // namespace base
