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
  int max_iterations_total = 300;
  int max_iterations_outer = 30;
  int max_iterations_inner = 100;
  double cost_tolerance = 1e-4;
  double gradient_tolerance = 1e-2;

  double bp_reg_increase_factor = 1.6;
  bool bp_reg_enable = true;
  double bp_reg_initial = 0.0;
  double bp_reg_max = 1e8;
  double bp_reg_min = 1e-8;
  int bp_reg_fail_threshold = 100;
  bool check_forwardpass_bounds = true;
  double state_max = 1e8;
  double control_max = 1e8;

  int line_search_max_iterations = 20;
  double line_search_lower_bound = 1e-8;
  double line_search_upper_bound = 10.0;
  double line_search_decrease_factor = 2;

  double constraint_tolerance = 1e-4;
  double maximum_penalty = 1e8;
  double initial_penalty = 1.0;  // every Solve() resets all penalties to this value; 0 disables
  bool reset_duals = true;

  int header_frequency = 10;
  LogLevel verbose = LogLevel::kSilent;
  bool profiler_enable = false;
  bool profiler_output_to_file = false;
  std::string log_directory;
  std::string profile_filename = "profiler.out";
  int nthreads = 1;
  int tasks_per_thread = 1;

  int NumThreads() const {
    if (nthreads == kPickHardwareThreads) return static_cast<int>(std::thread::hardware_concurrency());
    return std::max(nthreads, 1);
  }
};

}  // namespace altro
