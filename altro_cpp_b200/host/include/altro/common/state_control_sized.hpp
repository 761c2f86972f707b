// altro/common/state_control_sized.hpp (B200 host mirror) — run-time and compile-time state /
// control dimensions of a container (altro/common/state_control_sized.hpp:26,51 there).
//
// The template arguments are the sizes known to the compiler (Eigen::Dynamic when they are not), the
// constructor arguments the sizes known at run time; the two must agree wherever both are known.  On the device
// the same pair exists as the model's constexpr n, m (csrc/device.cuh) and the n, m of altro_b200_problem_create.
#pragma once

#include <eigen3/Eigen/Dense>

#include "altro/utils/assert.hpp"

namespace altro {

// n + m when both are compile-time constants, Eigen::Dynamic otherwise
constexpr int AddSizes(int n, int m) { return (n == Eigen::Dynamic || m == Eigen::Dynamic) ? Eigen::Dynamic : n + m; }

template <int n, int m>
class StateControlSized {
  // a compile-time size is binding, Eigen::Dynamic (negative) accepts anything
  static constexpr bool Agrees(int compile_time, int run_time) { return compile_time <= 0 || compile_time == run_time; }

 public:
  // sizes from the template arguments alone: both must be known
  StateControlSized() : n_(n), m_(m) {
    ALTRO_ASSERT(n > 0, "State dimension must be greater than zero.");
    ALTRO_ASSERT(m > 0, "Control dimension must be greater than zero.");
  }
  StateControlSized(int state_dim, int control_dim) : n_(state_dim), m_(control_dim) {
    ALTRO_ASSERT(Agrees(n, n_), "State sizes must be consistent.");
    ALTRO_ASSERT(Agrees(m, m_), "Control sizes must be consistent.");
  }

  // run-time sizes
  int StateDimension() const { return n_; }
  int ControlDimension() const { return m_; }
  // compile-time sizes (Eigen::Dynamic = not fixed)
  static constexpr int StateMemorySize() { return n; }
  static constexpr int ControlMemorySize() { return m; }

 protected:
  int n_;
  int m_;
};

}  // namespace altro
