// common.h -- shared by the translation units that implement the C ABI.
#pragma once
#include <string>

namespace emme {
// records `msg` as the calling thread's emme_last_error() text and returns `code`
int capi_fail(int code, const std::string& msg);
}  // namespace emme
