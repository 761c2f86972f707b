// altro/utils/assert.hpp (B200 host mirror) — ALTRO_ASSERT with the reference's contract
// (altro/utils/assert.hpp:6-10 there): message + abort() in debug builds, compiled out under NDEBUG.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <string>

#ifndef NDEBUG
#define ALTRO_ASSERT(Expr, Msg) altro::utils::AssertMsg((Expr), Msg, #Expr, __LINE__, __FILE__)
#else
#define ALTRO_ASSERT(Expr, Msg) ;
#endif

namespace altro {
namespace utils {

inline void AssertMsg(bool expr, const std::string& msg, const char* expr_str, int line, const char* file) {
  if (!expr) {
    std::fprintf(stderr, "Assert failed:\t%s\nExpected:\t%s\nSource:\t\t%s, line %d\n", msg.c_str(), expr_str, file, line);
    std::abort();
  }
}

constexpr bool AssertionsActive() {
#ifndef NDEBUG
  return true;
#else
  return false;
#endif
}

}  // namespace utils
}  // namespace altro
