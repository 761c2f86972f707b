// altro/common/solver_stats.hpp (B200 host mirror) — SolverStatus with the reference's values
// (altro/common/solver_stats.hpp:20-31) and the per-solve counters the benchmarks read.
#pragma once

#include "altro/common/solver_options.hpp"

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

// Counters of instance 0 of the batch (the batched solvers expose per-instance arrays as well).
struct SolverStats {
  double initial_cost = 0.0;
  int iterations_inner = 0;
  int iterations_outer = 0;
  int iterations_total = 0;
  SolverOptions& GetOptions() { return opts_; }
  const SolverOptions& GetOptions() const { return opts_; }

 private:
  SolverOptions opts_;
};

}  // namespace altro
