// altro/common/trajectory.hpp (B200 host mirror) — the caller's in/out buffer: initial guess in,
// solution out (altro/common/trajectory.hpp:25 there; ilqr.hpp:226-235).  Host-side AoS like the
// reference; the device keeps its own tile-major copy (csrc/common.cuh).
#pragma once

#include <cmath>
#include <iostream>
#include <vector>

#include "altro/common/knotpoint.hpp"
#include "altro/eigentypes.hpp"

namespace altro {

template <int n, int m, class T = double>
class Trajectory {
  using StateVector = VectorN<n, T>;
  using ControlVector = VectorN<m, T>;
  using Knot = KnotPoint<n, m, T>;

 public:
  explicit Trajectory(int N) : traj_(static_cast<size_t>(N) + 1) {}
  Trajectory(int state_dim, int control_dim, int N) : traj_(static_cast<size_t>(N) + 1, Knot(state_dim, control_dim)) {}
  explicit Trajectory(std::vector<Knot> zs) : traj_(std::move(zs)) {}
  Trajectory(const std::vector<StateVector>& X, const std::vector<ControlVector>& U, const std::vector<float>& times) {
    ALTRO_ASSERT(X.size() == U.size() + 1, "Length of control vector must be one less than the length of the state trajectory.");
    ALTRO_ASSERT(X.size() == times.size(), "Length of times vector must be equal to the length of the state trajectory.");
    const size_t N = U.size();
    for (size_t k = 0; k < N; ++k) traj_.emplace_back(X[k], U[k], times[k], times[k + 1] - times[k]);
    ControlVector uN = U[N - 1];
    uN.setZero();
    traj_.emplace_back(X[N], uN, times[N], 0.0F);
  }

  using iterator = typename std::vector<Knot>::iterator;
  using const_iterator = typename std::vector<Knot>::const_iterator;
  iterator begin() { return traj_.begin(); }
  const_iterator begin() const { return traj_.begin(); }
  iterator end() { return traj_.end(); }
  const_iterator end() const { return traj_.end(); }

  int NumSegments() const { return static_cast<int>(traj_.size()) - 1; }
  StateVector& State(int k) { return traj_[k].State(); }
  ControlVector& Control(int k) { return traj_[k].Control(); }
  const StateVector& State(int k) const { return traj_[k].State(); }
  const ControlVector& Control(int k) const { return traj_[k].Control(); }
  Knot& GetKnotPoint(int k) { return traj_[k]; }
  const Knot& GetKnotPoint(int k) const { return traj_[k]; }
  Knot& operator[](int k) { return traj_[k]; }
  int StateDimension(int k) const { return traj_[k].StateDimension(); }
  int ControlDimension(int k) const { return traj_[k].ControlDimension(); }
  T GetTime(int k) const { return traj_[k].GetTime(); }
  float GetStep(int k) const { return traj_[k].GetStep(); }
  void SetTime(int k, float t) { traj_[k].SetTime(t); }
  void SetStep(int k, float h) { traj_[k].SetStep(h); }
  void SetZero() {
    for (Knot& z : traj_) {
      z.State().setZero();
      z.Control().setZero();
    }
  }
  // t_k = float(k) * h in float arithmetic, the terminal knot gets step 0 (trajectory.hpp:122-130 there)
  void SetUniformStep(float h) {
    const int N = NumSegments();
    for (int k = 0; k < N; ++k) {
      traj_[k].SetStep(h);
      traj_[k].SetTime(static_cast<float>(k) * h);
    }
    traj_[N].SetStep(0.0F);
    traj_[N].SetTime(static_cast<float>(h) * N);
  }
  bool CheckTimeConsistency(double eps = 1e-6, bool verbose = false) {
    for (int k = 0; k < NumSegments(); ++k) {
      const float h_calc = GetTime(k + 1) - GetTime(k);
      if (std::abs(GetStep(k) - h_calc) > eps) {
        if (verbose) std::cout << "k=" << k << "\t h=" << GetStep(k) << "\t dt=" << h_calc << std::endl;
        return false;
      }
    }
    return true;
  }

 private:
  std::vector<Knot> traj_;
};

using TrajectoryXXd = Trajectory<Eigen::Dynamic, Eigen::Dynamic, double>;

}  // namespace altro
