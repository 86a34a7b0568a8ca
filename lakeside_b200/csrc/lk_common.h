// Shared host-side helpers for liblakeside_b200 (error type, status codes).
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/lakeside_b200.h"

namespace lk {

// Every failure inside the library is an lk::Error; the C-ABI boundary converts it into a status code plus a
// thread-local message (lk_last_error).  There is deliberately no CPU fallback anywhere: a query the GPU path
// cannot run is LK_ERR_UNSUPPORTED and the adapter maps it to the reference's "stream nothing" behaviour
// (core/src/main/scala/com/cardinal/utils/Commons.scala:249-253).
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const std::string& msg) { throw Error(code, msg); }

#define LK_CHECK(cond, code, msg)                    \
  do {                                               \
    if (!(cond)) ::lk::fail((code), std::string(msg)); \
  } while (0)

inline std::string strf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
}  // namespace lk

#include <cstdarg>
inline std::string lk::strf(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return std::string(buf);
}
