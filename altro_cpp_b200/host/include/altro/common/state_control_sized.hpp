// altro/common/state_control_sized.hpp (B200 host mirror) — run-time and compile-time state /
// control dimensions of a container (altro/common/state_control_sized.hpp:26,51 there).
#pragma once

#include <eigen3/Eigen/Dense>

#include "altro/utils/assert.hpp"

namespace altro {

constexpr int AddSizes(int n, int m) { return (n == Eigen::Dynamic || m == Eigen::Dynamic) ? Eigen::Dynamic : n + m; }

template <int n, int m>
class StateControlSized {
 public:
  StateControlSized(int state_dim, int control_dim) : n_(state_dim), m_(control_dim) {
    ALTRO_ASSERT(n <= 0 || n == n_, "State sizes must be consistent.");
    ALTRO_ASSERT(m <= 0 || m == m_, "Control sizes must be consistent.");
  }
  StateControlSized() : n_(n), m_(m) {
    ALTRO_ASSERT(n > 0, "State dimension must be greater than zero.");
    ALTRO_ASSERT(m > 0, "Control dimension must be greater than zero.");
  }
  int StateDimension() const { return n_; }
  int ControlDimension() const { return m_; }
  static constexpr int StateMemorySize() { return n; }
  static constexpr int ControlMemorySize() { return m; }

 protected:
  int n_;
  int m_;
};

}  // namespace altro
