// altro/common/solver_logger.hpp (B200 host mirror) — verbosity levels of the reference's console
// logger (altro/common/log_entry.hpp there).  The tabular logger itself is host-side observability
// outside the hot path (out of scope); what the solver API needs are the levels
// that SolverOptions::verbose takes.
#pragma once

namespace altro {

enum class LogLevel { kSilent = 0, kOuter = 1, kOuterDebug = 2, kInner = 3, kInnerDebug = 4, kDebug = 5 };

class SolverLogger {
 public:
  explicit SolverLogger(LogLevel level = LogLevel::kSilent) : level_(level) {}
  LogLevel GetLevel() const { return level_; }
  void SetLevel(LogLevel level) { level_ = level; }
  void Disable() { level_ = LogLevel::kSilent; }
  void SetFrequency(int freq) { frequency_ = freq; }
  void Clear() {}

 private:
  LogLevel level_;
  int frequency_ = 10;
};

}  // namespace altro
