// altro/utils/assert.hpp (B200 host mirror) — ALTRO_ASSERT(condition, message): the reference's contract
// (altro/utils/assert.hpp:6-10 there) is "report and abort() in debug builds, nothing under NDEBUG".
// Contract violations that reach the device library are reported a second time, in every build type, as
// altro::DeviceError carrying altro_b200_last_error() (altro/device_solver.hpp).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <string>

namespace altro {
namespace utils {

#if defined(NDEBUG)
constexpr bool kAssertionsCompiledIn = false;
#else
constexpr bool kAssertionsCompiledIn = true;
#endif
constexpr bool AssertionsActive() { return kAssertionsCompiledIn; }

// prints what was expected, where, and the caller's message; then aborts
inline void AssertMsg(bool holds, const std::string& message, const char* condition_text, int line, const char* file) {
  if (holds) return;
  std::fprintf(stderr, "%s:%d: Assertion (%s) failed: %s\n", file, line, condition_text, message.c_str());
  std::abort();
}

}  // namespace utils
}  // namespace altro

#if defined(NDEBUG)
#define ALTRO_ASSERT(condition, message) ;
#else
#define ALTRO_ASSERT(condition, message) \
  ::altro::utils::AssertMsg(static_cast<bool>(condition), (message), #condition, __LINE__, __FILE__)
#endif
