// altro/common/solver_stats.hpp (B200 host mirror) — SolverStatus with the reference's values and
// SolverStats with its per-iteration history vectors (altro/common/solver_stats.hpp:20-31, 39-129
// there), including the carry-forward semantics of Log / NewIteration that the convergence test
// reads (solver_stats.cpp:54-66 there).  The device keeps the same history per
// instance when asked to (include/altro_b200.h altro_b200_get_history_host); the solvers copy it here.
#pragma once

#include <string>
#include <unordered_map>
#include <vector>

#include "altro/common/solver_logger.hpp"
#include "altro/common/solver_options.hpp"
#include "altro/common/timer.hpp"
#include "altro/utils/utils.hpp"

namespace altro {

enum class SolverStatus {
  kSolved = 0,
  kUnsolved = 1,
  kStateLimit = 2,
  kControlLimit = 3,
  kCostIncrease = 4,
  kMaxIterations = 5,
  kMaxOuterIterations = 6,
  kMaxInnerIterations = 7,
  kMaxPenalty = 8,
  kBackwardPassRegularizationFailed = 9,
};

class SolverStats {
 public:
  SolverStats() : timer_(Timer::MakeShared()) {
    floats_ = {{"cost", &cost},        {"viol", &violations},     {"dJ", &cost_decrease}, {"grad", &gradient},
               {"alpha", &alpha},      {"reg", &regularization},  {"z", &improvement_ratio}, {"pen", &max_penalty}};
  }
  SolverStats(const SolverStats& o) : SolverStats() { CopyFrom(o); }
  SolverStats& operator=(const SolverStats& o) {
    CopyFrom(o);
    return *this;
  }

  double initial_cost = 0.0;
  int iterations_inner = 0;
  int iterations_outer = 0;
  int iterations_total = 0;
  std::vector<double> cost;
  std::vector<double> alpha;
  std::vector<double> improvement_ratio;  // actual / expected cost decrease of the accepted step
  std::vector<double> gradient;
  std::vector<double> cost_decrease;
  std::vector<double> regularization;
  std::vector<double> violations;   // maximum constraint violation
  std::vector<double> max_penalty;  // maximum penalty parameter

  void SetCapacity(int n) {
    for (auto& kv : floats_) kv.second->reserve(static_cast<size_t>(n));
  }
  void Reset() {
    initial_cost = 0.0;
    iterations_inner = iterations_outer = iterations_total = 0;
    len_ = 0;
    for (auto& kv : floats_) kv.second->clear();
    SetCapacity(opts_.max_iterations_total);
    logger_.SetLevel(opts_.verbose);
    // the profiler follows the options of the moment, like everything else Reset() re-reads
    if (opts_.profiler_enable) timer_->Activate();
    else timer_->Deactivate();
    if (opts_.profiler_output_to_file) ProfilerOutputToFile(true);
  }
  void SetTolerances(const double&, const double&, const double&) {}
  void SetVerbosity(LogLevel level) { logger_.SetLevel(level); }
  LogLevel GetVerbosity() const { return logger_.GetLevel(); }
  SolverLogger& GetLogger() { return logger_; }
  const SolverLogger& GetLogger() const { return logger_; }
  TimerPtr& GetTimer() { return timer_; }
  const TimerPtr& GetTimer() const { return timer_; }
  SolverOptions& GetOptions() { return opts_; }
  const SolverOptions& GetOptions() const { return opts_; }
  std::string ProfileOutputFile() { return opts_.log_directory + "/" + opts_.profile_filename; }
  void ProfilerOutputToFile(bool flag) {
    if (flag) timer_->SetOutput(ProfileOutputFile());
    else timer_->SetOutput(stdout);
  }
  void PrintLast() {}

  // records `value` in the current row of the named history (integer entries such as "iters" have no history)
  template <class T>
  void Log(const std::string& title, T value) {
    if (len_ == 0) NewIteration();
    auto it = floats_.find(title);
    if (it != floats_.end()) it->second->back() = static_cast<double>(value);
  }
  // opens a new row that starts as a copy of the previous one
  void NewIteration() {
    ++len_;
    for (auto& kv : floats_) {
      std::vector<double>& v = *kv.second;
      v.resize(static_cast<size_t>(len_));
      v.back() = len_ > 1 ? v[static_cast<size_t>(len_) - 2] : 0.0;
    }
  }
  // whole histories at once (rows as the device recorded them)
  void SetHistory(int rows, const double* cost_, const double* alpha_, const double* z_, const double* grad_,
                  const double* dJ_, const double* reg_, const double* viol_, const double* pen_) {
    len_ = rows;
    cost.assign(cost_, cost_ + rows);
    alpha.assign(alpha_, alpha_ + rows);
    improvement_ratio.assign(z_, z_ + rows);
    gradient.assign(grad_, grad_ + rows);
    cost_decrease.assign(dJ_, dJ_ + rows);
    regularization.assign(reg_, reg_ + rows);
    violations.assign(viol_, viol_ + rows);
    max_penalty.assign(pen_, pen_ + rows);
  }

 private:
  void CopyFrom(const SolverStats& o) {
    initial_cost = o.initial_cost;
    iterations_inner = o.iterations_inner;
    iterations_outer = o.iterations_outer;
    iterations_total = o.iterations_total;
    cost = o.cost; alpha = o.alpha; improvement_ratio = o.improvement_ratio; gradient = o.gradient;
    cost_decrease = o.cost_decrease; regularization = o.regularization; violations = o.violations;
    max_penalty = o.max_penalty;
    len_ = o.len_;
    logger_ = o.logger_;
    opts_ = o.opts_;
  }
  std::unordered_map<std::string, std::vector<double>*> floats_;
  int len_ = 0;
  SolverLogger logger_;
  TimerPtr timer_;
  SolverOptions opts_;
};

}  // namespace altro
