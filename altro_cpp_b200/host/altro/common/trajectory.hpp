// altro/common/trajectory.hpp (B200 host mirror) — KnotPoint / Trajectory as the in/out buffer of
// a solve (reference: altro/common/knotpoint.hpp:32, trajectory.hpp:25).  States and controls
// are double, time and step are float exactly as in the reference (knotpoint.hpp:179-180).
#pragma once

#include <vector>

#include "altro/eigentypes.hpp"

namespace altro {

template <int n, int m>
class KnotPoint {
 public:
  KnotPoint(int nn, int mm) : x_(nn), u_(mm) {}
  VectorXd& State() { return x_; }
  VectorXd& Control() { return u_; }
  const VectorXd& State() const { return x_; }
  const VectorXd& Control() const { return u_; }
  float GetTime() const { return t_; }
  float GetStep() const { return h_; }
  void SetTime(float t) { t_ = t; }
  void SetStep(float h) { h_ = h; }

 private:
  VectorXd x_, u_;
  float t_ = 0.0f, h_ = 0.0f;
};

template <int n, int m>
class Trajectory {
 public:
  explicit Trajectory(int N) : Trajectory(n, m, N) {}
  Trajectory(int nn, int mm, int N) : traj_(N + 1, KnotPoint<n, m>(nn, mm)) {}
  int NumSegments() const { return static_cast<int>(traj_.size()) - 1; }
  VectorXd& State(int k) { return traj_.at(k).State(); }
  VectorXd& Control(int k) { return traj_.at(k).Control(); }
  const VectorXd& State(int k) const { return traj_.at(k).State(); }
  const VectorXd& Control(int k) const { return traj_.at(k).Control(); }
  KnotPoint<n, m>& GetKnotPoint(int k) { return traj_.at(k); }
  float GetTime(int k) const { return traj_.at(k).GetTime(); }
  float GetStep(int k) const { return traj_.at(k).GetStep(); }
  // trajectory.hpp:122-130 (float arithmetic on purpose)
  void SetUniformStep(float h) {
    const int N = NumSegments();
    for (int k = 0; k < N; ++k) {
      traj_[k].SetStep(h);
      traj_[k].SetTime(static_cast<float>(k) * h);
    }
    traj_[N].SetStep(0.0f);
    traj_[N].SetTime(static_cast<float>(h) * N);
  }
  void SetZero() {
    for (auto& z : traj_) {
      z.State().setZero();
      z.Control().setZero();
    }
  }

 private:
  std::vector<KnotPoint<n, m>> traj_;
};

}  // namespace altro
