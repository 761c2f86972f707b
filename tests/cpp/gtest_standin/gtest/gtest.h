// gtest/gtest.h — a small stand-in for the slice of GoogleTest the reference's unit tests use (TEST, TEST_F,
// EXPECT_* / ASSERT_* with streamed messages, EXPECT_DEATH, EXPECT_THROW / EXPECT_NO_THROW, testing::Test),
// so that those test sources compile UNMODIFIED against this repo's host mirror where GoogleTest is not
// installed (tests/test_reference_unit_tests.py).  Test infrastructure; not part of the product.
#pragma once

#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <regex>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace testing {

class Message {
 public:
  template <class T>
  Message& operator<<(const T& v) {
    os_ << v;
    return *this;
  }
  Message& operator<<(std::ostream& (*manip)(std::ostream&)) {
    os_ << manip;
    return *this;
  }
  std::string str() const { return os_.str(); }

 private:
  std::ostringstream os_;
};

class Test {
 public:
  virtual ~Test() = default;
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
  static void SetUpTestSuite() {}
  static void TearDownTestSuite() {}
  static void SetUpTestCase() {}
  static void TearDownTestCase() {}
};

namespace internal {

struct State {
  int failures_in_current = 0;
  bool fatal = false;
  static State& Get() {
    static State s;
    return s;
  }
};

struct Entry {
  std::string suite, name;
  std::function<std::unique_ptr<Test>()> make;
};
inline std::vector<Entry>& Registry() {
  static std::vector<Entry> r;
  return r;
}
struct Registrar {
  Registrar(const char* suite, const char* name, std::function<std::unique_ptr<Test>()> make) {
    Registry().push_back({suite, name, std::move(make)});
  }
};

// outcome of one check: ok + what to say when it is not
struct Result {
  bool ok;
  std::string detail;
  explicit operator bool() const { return ok; }
};

template <class T, class = void>
struct Streamable : std::false_type {};
template <class T>
struct Streamable<T, decltype(void(std::declval<std::ostream&>() << std::declval<const T&>()))> : std::true_type {};

template <class T>
typename std::enable_if<Streamable<T>::value, std::string>::type Print(const T& v) {
  std::ostringstream os;
  os.precision(17);
  os << v;
  return os.str();
}
template <class T>
typename std::enable_if<!Streamable<T>::value, std::string>::type Print(const T&) {
  return "<object of " + std::to_string(sizeof(T)) + " bytes>";
}
inline std::string Print(std::nullptr_t) { return "nullptr"; }
inline std::string Print(bool v) { return v ? "true" : "false"; }

template <class A, class B, class Op>
Result Compare(const A& a, const B& b, Op op, const char* opname, const char* ta, const char* tb) {
  if (op(a, b)) return {true, ""};
  return {false, std::string("Expected: (") + ta + ") " + opname + " (" + tb + "), actual: " + Print(a) + " vs " + Print(b)};
}

template <class Raw, class Bits>
bool AlmostEqualUlps(Raw a, Raw b) {  // GoogleTest's rule: within 4 units in the last place
  if (std::isnan(a) || std::isnan(b)) return false;
  Bits ia, ib;
  std::memcpy(&ia, &a, sizeof(Raw));
  std::memcpy(&ib, &b, sizeof(Raw));
  const Bits sign = Bits(1) << (sizeof(Bits) * 8 - 1);
  auto biased = [sign](Bits x) { return (x & sign) ? Bits(~x + 1) : Bits(sign | x); };
  const Bits ba = biased(ia), bb = biased(ib);
  return (ba >= bb ? ba - bb : bb - ba) <= 4;
}
inline Result DoubleEq(double a, double b, const char* ta, const char* tb) {
  if (AlmostEqualUlps<double, std::uint64_t>(a, b)) return {true, ""};
  return {false, std::string("Expected equality of ") + ta + " and " + tb + ": " + Print(a) + " vs " + Print(b)};
}
inline Result FloatEq(float a, float b, const char* ta, const char* tb) {
  if (AlmostEqualUlps<float, std::uint32_t>(a, b)) return {true, ""};
  return {false, std::string("Expected equality of ") + ta + " and " + tb + ": " + Print(a) + " vs " + Print(b)};
}
inline Result Near(double a, double b, double tol, const char* ta, const char* tb) {
  if (std::fabs(a - b) <= tol) return {true, ""};
  return {false, std::string("The difference between ") + ta + " and " + tb + " is " + Print(std::fabs(a - b)) +
                     ", which exceeds " + Print(tol)};
}
inline Result Truth(bool v, bool want, const char* text) {
  if (v == want) return {true, ""};
  return {false, std::string("Value of: ") + text + "\n  Actual: " + (v ? "true" : "false") + "\nExpected: " + (want ? "true" : "false")};
}

// runs `body` in a forked child; passes when the child dies (signal or non-zero exit) and its stderr matches
inline Result Death(const std::function<void()>& body, const char* pattern, const char* text) {
  int fds[2];
  if (pipe(fds) != 0) return {false, "pipe() failed"};
  std::fflush(nullptr);
  const pid_t pid = fork();
  if (pid < 0) return {false, "fork() failed"};
  if (pid == 0) {
    close(fds[0]);
    dup2(fds[1], 2);
    close(fds[1]);
    body();
    std::fflush(nullptr);
    _exit(0);
  }
  close(fds[1]);
  std::string err;
  char buf[512];
  ssize_t got;
  while ((got = read(fds[0], buf, sizeof(buf))) > 0) err.append(buf, static_cast<size_t>(got));
  close(fds[0]);
  int status = 0;
  waitpid(pid, &status, 0);
  const bool died = WIFSIGNALED(status) || (WIFEXITED(status) && WEXITSTATUS(status) != 0);
  if (!died) return {false, std::string("Death test: ") + text + " failed to die."};
  if (pattern && *pattern && !std::regex_search(err, std::regex(pattern)))
    return {false, std::string("Death test: ") + text + " died but its output does not match \"" + pattern + "\": " + err};
  return {true, ""};
}

class Reporter {
 public:
  Reporter(const char* file, int line, const std::string& detail, bool fatal) : file_(file), line_(line), detail_(detail), fatal_(fatal) {}
  void operator=(const Message& m) const {
    std::fprintf(stderr, "%s:%d: Failure\n%s\n", file_, line_, detail_.c_str());
    const std::string extra = m.str();
    if (!extra.empty()) std::fprintf(stderr, "%s\n", extra.c_str());
    State::Get().failures_in_current++;
    if (fatal_) State::Get().fatal = true;
  }

 private:
  const char* file_;
  int line_;
  std::string detail_;
  bool fatal_;
};

}  // namespace internal

namespace internal {
// --gtest_filter=POSITIVE[-NEGATIVE], each a ':'-separated list of glob patterns ('*' and '?') on "Suite.Name"
inline std::string& Filter() {
  static std::string filter = "*";
  return filter;
}
inline bool GlobMatch(const char* pat, const char* text) {
  if (*pat == '\0') return *text == '\0';
  if (*pat == '*') return GlobMatch(pat + 1, text) || (*text != '\0' && GlobMatch(pat, text + 1));
  return *text != '\0' && (*pat == '?' || *pat == *text) && GlobMatch(pat + 1, text + 1);
}
inline bool AnyPatternMatches(const std::string& patterns, const std::string& name) {
  std::size_t start = 0;
  while (start <= patterns.size()) {
    std::size_t stop = patterns.find(':', start);
    if (stop == std::string::npos) stop = patterns.size();
    if (stop > start && GlobMatch(patterns.substr(start, stop - start).c_str(), name.c_str())) return true;
    start = stop + 1;
  }
  return false;
}
inline bool Selected(const std::string& name) {
  const std::string& f = Filter();
  const std::size_t dash = f.find('-');
  const std::string positive = dash == std::string::npos ? f : f.substr(0, dash);
  const std::string negative = dash == std::string::npos ? std::string() : f.substr(dash + 1);
  return AnyPatternMatches(positive.empty() ? std::string("*") : positive, name) && !AnyPatternMatches(negative, name);
}
}  // namespace internal

inline void InitGoogleTest(int* argc, char** argv) {
  const std::string flag = "--gtest_filter=";
  for (int i = 1; argc && i < *argc; ++i) {
    const std::string arg = argv[i];
    if (arg.compare(0, flag.size(), flag) == 0) internal::Filter() = arg.substr(flag.size());
  }
}
inline void InitGoogleTest() {}

inline int RunAllTests() {
  int failed = 0, ran = 0;
  for (const internal::Entry& e : internal::Registry()) {
    if (!internal::Selected(e.suite + "." + e.name)) continue;
    std::printf("[ RUN      ] %s.%s\n", e.suite.c_str(), e.name.c_str());
    std::fflush(stdout);
    internal::State::Get().failures_in_current = 0;
    internal::State::Get().fatal = false;
    try {
      std::unique_ptr<Test> t = e.make();
      t->SetUp();
      if (!internal::State::Get().fatal) t->TestBody();
      t->TearDown();
    } catch (const std::exception& ex) {
      std::fprintf(stderr, "unexpected exception: %s\n", ex.what());
      internal::State::Get().failures_in_current++;
    } catch (...) {
      std::fprintf(stderr, "unexpected exception of unknown type\n");
      internal::State::Get().failures_in_current++;
    }
    ++ran;
    if (internal::State::Get().failures_in_current) {
      ++failed;
      std::printf("[  FAILED  ] %s.%s\n", e.suite.c_str(), e.name.c_str());
    } else {
      std::printf("[       OK ] %s.%s\n", e.suite.c_str(), e.name.c_str());
    }
  }
  std::printf("[==========] %d tests ran, %d failed.\n", ran, failed);
  return failed ? 1 : 0;
}

}  // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAllTests()

#define GTS_TEST_CLASS_(suite, name) suite##_##name##_Test
#define GTS_DEFINE_TEST_(suite, name, base)                                                              \
  class GTS_TEST_CLASS_(suite, name) : public base {                                                     \
   public:                                                                                               \
    void TestBody() override;                                                                            \
  };                                                                                                     \
  static ::testing::internal::Registrar gts_registrar_##suite##_##name(                                  \
      #suite, #name, [] { return std::unique_ptr<::testing::Test>(new GTS_TEST_CLASS_(suite, name)); }); \
  void GTS_TEST_CLASS_(suite, name)::TestBody()
#define TEST(suite, name) GTS_DEFINE_TEST_(suite, name, ::testing::Test)
#define TEST_F(fixture, name) GTS_DEFINE_TEST_(fixture, name, fixture)

// a check is a statement that may be followed by `<< message`
#define GTS_CHECK_(result_expr, fatal)                                             \
  switch (0)                                                                       \
  case 0:                                                                          \
  default:                                                                         \
    if (const ::testing::internal::Result gts_r_ = (result_expr))                  \
      ;                                                                            \
    else                                                                           \
      GTS_ON_FAILURE_##fatal ::testing::internal::Reporter(__FILE__, __LINE__, gts_r_.detail, fatal) = ::testing::Message()
#define GTS_ON_FAILURE_false
#define GTS_ON_FAILURE_true return

#define GTS_CMP_(a, b, op, opname, fatal) \
  GTS_CHECK_(::testing::internal::Compare((a), (b), [](const auto& x_, const auto& y_) { return x_ op y_; }, opname, #a, #b), fatal)

#define EXPECT_TRUE(c) GTS_CHECK_(::testing::internal::Truth(static_cast<bool>(c), true, #c), false)
#define EXPECT_FALSE(c) GTS_CHECK_(::testing::internal::Truth(static_cast<bool>(c), false, #c), false)
#define ASSERT_TRUE(c) GTS_CHECK_(::testing::internal::Truth(static_cast<bool>(c), true, #c), true)
#define ASSERT_FALSE(c) GTS_CHECK_(::testing::internal::Truth(static_cast<bool>(c), false, #c), true)
#define EXPECT_EQ(a, b) GTS_CMP_(a, b, ==, "==", false)
#define EXPECT_NE(a, b) GTS_CMP_(a, b, !=, "!=", false)
#define EXPECT_LT(a, b) GTS_CMP_(a, b, <, "<", false)
#define EXPECT_LE(a, b) GTS_CMP_(a, b, <=, "<=", false)
#define EXPECT_GT(a, b) GTS_CMP_(a, b, >, ">", false)
#define EXPECT_GE(a, b) GTS_CMP_(a, b, >=, ">=", false)
#define ASSERT_EQ(a, b) GTS_CMP_(a, b, ==, "==", true)
#define ASSERT_NE(a, b) GTS_CMP_(a, b, !=, "!=", true)
#define ASSERT_LT(a, b) GTS_CMP_(a, b, <, "<", true)
#define ASSERT_LE(a, b) GTS_CMP_(a, b, <=, "<=", true)
#define ASSERT_GT(a, b) GTS_CMP_(a, b, >, ">", true)
#define ASSERT_GE(a, b) GTS_CMP_(a, b, >=, ">=", true)
#define EXPECT_DOUBLE_EQ(a, b) GTS_CHECK_(::testing::internal::DoubleEq((a), (b), #a, #b), false)
#define ASSERT_DOUBLE_EQ(a, b) GTS_CHECK_(::testing::internal::DoubleEq((a), (b), #a, #b), true)
#define EXPECT_FLOAT_EQ(a, b) GTS_CHECK_(::testing::internal::FloatEq((a), (b), #a, #b), false)
#define ASSERT_FLOAT_EQ(a, b) GTS_CHECK_(::testing::internal::FloatEq((a), (b), #a, #b), true)
#define EXPECT_NEAR(a, b, tol) GTS_CHECK_(::testing::internal::Near((a), (b), (tol), #a, #b), false)
#define ASSERT_NEAR(a, b, tol) GTS_CHECK_(::testing::internal::Near((a), (b), (tol), #a, #b), true)
#define EXPECT_DEATH(statement, pattern) \
  GTS_CHECK_(::testing::internal::Death([&]() { statement; }, pattern, #statement), false)
#define ASSERT_DEATH(statement, pattern) \
  GTS_CHECK_(::testing::internal::Death([&]() { statement; }, pattern, #statement), true)

#define GTS_THROW_RESULT_(statement, expected_type, want_throw)                                               \
  [&]() -> ::testing::internal::Result {                                                                      \
    try {                                                                                                     \
      statement;                                                                                              \
    } catch (const expected_type&) {                                                                          \
      return {want_throw, std::string(#statement) + " threw " #expected_type};                               \
    } catch (...) {                                                                                           \
      return {false, std::string(#statement) + " threw an exception of another type"};                       \
    }                                                                                                         \
    return {!want_throw, std::string(#statement) + " threw nothing"};                                        \
  }()
#define EXPECT_THROW(statement, expected_type) GTS_CHECK_(GTS_THROW_RESULT_(statement, expected_type, true), false)
#define ASSERT_THROW(statement, expected_type) GTS_CHECK_(GTS_THROW_RESULT_(statement, expected_type, true), true)
#define EXPECT_NO_THROW(statement)                                                    \
  GTS_CHECK_(([&]() -> ::testing::internal::Result {                                  \
               try {                                                                  \
                 statement;                                                           \
               } catch (...) {                                                        \
                 return {false, std::string(#statement) + " threw an exception"};    \
               }                                                                      \
               return {true, ""};                                                     \
             }()),                                                                    \
             false)
