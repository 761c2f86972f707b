// altro/common/solver_options.hpp (B200 host mirror) — same fields, defaults and meaning as the
// reference's SolverOptions (altro/common/solver_options.hpp:19-65 there).  The numeric fields go to
// the device unchanged (include/altro_b200.h altro_b200_options); logging, profiler and thread knobs
// are kept so that existing programs assign them, and have no device effect (the thread pool is
// replaced by the batch axis).
#pragma once

#include <algorithm>
#include <string>
#include <thread>

#include "altro/common/solver_logger.hpp"
#include "altro/utils/utils.hpp"

namespace altro {

constexpr int kPickHardwareThreads = -1;

struct SolverOptions {
  // ---- termination (device: altro_b200_options, same names)
  int max_iterations_total = 300;     // iLQR iterations over a whole AL solve -> kMaxIterations
  int max_iterations_outer = 30;      // AL (dual-update) iterations -> kMaxOuterIterations
  int max_iterations_inner = 100;     // iLQR iterations of one AL iteration -> kMaxInnerIterations
  double cost_tolerance = 1e-4;       // an iLQR solve ends when the cost decrease AND ...
  double gradient_tolerance = 1e-2;   // ... the normalised feed-forward gain are below these
  double constraint_tolerance = 1e-4; // the AL solve ends when the max violation is below this

  // ---- backward pass: regularisation of Quu when its LLT fails (ilqr.hpp:401-442 there)
  bool bp_reg_enable = true;
  double bp_reg_initial = 0.0;
  double bp_reg_increase_factor = 1.6;
  double bp_reg_min = 1e-8;
  double bp_reg_max = 1e8;
  int bp_reg_fail_threshold = 100;    // restarts at bp_reg_max before kBackwardPassRegularizationFailed

  // ---- forward pass: backtracking line search on alpha = 1, 1/f, 1/f^2, ...
  int line_search_max_iterations = 20;
  double line_search_decrease_factor = 2;
  double line_search_lower_bound = 1e-8;  // accepted when lower <= actual / expected decrease <= upper
  double line_search_upper_bound = 10.0;
  bool check_forwardpass_bounds = true;   // a rollout leaving |x| <= state_max, |u| <= control_max is rejected
  double state_max = 1e8;
  double control_max = 1e8;

  // ---- augmented Lagrangian
  double initial_penalty = 1.0;  // every Solve() resets all penalties to this value; 0 keeps them
  double maximum_penalty = 1e8;  // -> kMaxPenalty
  bool reset_duals = true;       // false = warm start from the previous solve's multipliers

  // ---- host-side knobs kept for source compatibility; no effect on the device solve
  LogLevel verbose = LogLevel::kSilent;
  int header_frequency = 10;
  bool profiler_enable = false;
  bool profiler_output_to_file = false;
  std::string log_directory;
  std::string profile_filename = "profiler.out";
  int nthreads = 1;          // the reference's thread pool; here the batch axis takes its place
  int tasks_per_thread = 1;

  int NumThreads() const {
    if (nthreads == kPickHardwareThreads) return static_cast<int>(std::thread::hardware_concurrency());
    return std::max(nthreads, 1);
  }
};

}  // namespace altro
