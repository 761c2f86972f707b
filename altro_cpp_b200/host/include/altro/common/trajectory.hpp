// altro/common/trajectory.hpp (B200 host mirror) — the caller's in/out buffer: initial guess in,
// solution out (altro/common/trajectory.hpp:25 there; ilqr.hpp:226-235).  Host-side array of knot points
// like the reference; the device keeps its own tile-major copy (csrc/common.cuh) that Upload / Download of
// altro/device_solver.hpp convert to and from.
#pragma once

#include <cmath>
#include <cstddef>
#include <iostream>
#include <utility>
#include <vector>

#include "altro/common/knotpoint.hpp"
#include "altro/eigentypes.hpp"
#include "altro/utils/assert.hpp"

namespace altro {

template <int n, int m, class T = double>
class Trajectory {
 public:
  using Knot = KnotPoint<n, m, T>;
  using iterator = typename std::vector<Knot>::iterator;
  using const_iterator = typename std::vector<Knot>::const_iterator;

 private:
  using StateVector = VectorN<n, T>;
  using ControlVector = VectorN<m, T>;
  std::vector<Knot> traj_;  // N + 1 knot points; the last one carries u = 0 and h = 0
  Knot& at(int k) { return traj_[static_cast<std::size_t>(k)]; }
  const Knot& at(int k) const { return traj_[static_cast<std::size_t>(k)]; }

 public:
  // ---- construction: N segments = N + 1 states, N controls
  explicit Trajectory(int N) : traj_(static_cast<std::size_t>(N) + 1) {}
  Trajectory(int state_dim, int control_dim, int N) : traj_(static_cast<std::size_t>(N) + 1, Knot(state_dim, control_dim)) {}
  explicit Trajectory(std::vector<Knot> knots) : traj_(std::move(knots)) {}
  // from separate state / control / time sequences (times has one entry per state)
  Trajectory(const std::vector<StateVector>& X, const std::vector<ControlVector>& U, const std::vector<float>& times) {
    ALTRO_ASSERT(X.size() == U.size() + 1, "Length of control vector must be one less than the length of the state trajectory.");
    ALTRO_ASSERT(X.size() == times.size(), "Length of times vector must be equal to the length of the state trajectory.");
    traj_.reserve(X.size());
    for (std::size_t k = 0; k + 1 < X.size(); ++k) traj_.emplace_back(X[k], U[k], times[k], times[k + 1] - times[k]);
    ControlVector no_control = U.back();
    no_control.setZero();
    traj_.emplace_back(X.back(), no_control, times.back(), 0.0F);
  }

  // ---- size and time grid
  int NumSegments() const { return static_cast<int>(traj_.size()) - 1; }
  int StateDimension(int k) const { return at(k).StateDimension(); }
  int ControlDimension(int k) const { return at(k).ControlDimension(); }
  float GetStep(int k) const { return at(k).GetStep(); }
  T GetTime(int k) const { return at(k).GetTime(); }
  void SetStep(int k, float h) { at(k).SetStep(h); }
  void SetTime(int k, float t) { at(k).SetTime(t); }
  // t_k = float(k) * h in float arithmetic; the terminal knot gets step 0 and t_N = float(h) * N — the
  // roundings the device reproduces (trajectory.hpp:122-130 there)
  void SetUniformStep(float h) {
    const int last = NumSegments();
    for (int k = 0; k <= last; ++k) {
      at(k).SetStep(k < last ? h : 0.0F);
      at(k).SetTime(k < last ? static_cast<float>(k) * h : static_cast<float>(h) * last);
    }
  }
  // true when t[k+1] - t[k] agrees with h[k] everywhere
  bool CheckTimeConsistency(double eps = 1e-6, bool verbose = false) {
    for (int k = 0; k < NumSegments(); ++k) {
      const float dt = GetTime(k + 1) - GetTime(k);
      if (std::abs(GetStep(k) - dt) <= eps) continue;
      if (verbose) std::cout << "knot " << k << ": step " << GetStep(k) << " but t[k+1] - t[k] = " << dt << std::endl;
      return false;
    }
    return true;
  }

  // ---- element access
  Knot& GetKnotPoint(int k) { return at(k); }
  const Knot& GetKnotPoint(int k) const { return at(k); }
  Knot& operator[](int k) { return at(k); }
  StateVector& State(int k) { return at(k).State(); }
  const StateVector& State(int k) const { return at(k).State(); }
  ControlVector& Control(int k) { return at(k).Control(); }
  const ControlVector& Control(int k) const { return at(k).Control(); }
  void SetZero() {
    for (Knot& z : traj_) {
      z.State().setZero();
      z.Control().setZero();
    }
  }
  iterator begin() { return traj_.begin(); }
  iterator end() { return traj_.end(); }
  const_iterator begin() const { return traj_.begin(); }
  const_iterator end() const { return traj_.end(); }
};

using TrajectoryXXd = Trajectory<Eigen::Dynamic, Eigen::Dynamic, double>;

}  // namespace altro
