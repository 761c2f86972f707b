// altro/common/knotpoint.hpp (B200 host mirror) — state, control, float time and float step of one
// knot (altro/common/knotpoint.hpp:32 there; t and h are `float`).
#pragma once

#include <cstdlib>
#include <ostream>
#include <sstream>
#include <string>

#include "altro/common/state_control_sized.hpp"
#include "altro/eigentypes.hpp"

namespace altro {

template <int n, int m, class T = double>
class KnotPoint : public StateControlSized<n, m> {
  using StateVector = VectorN<n, T>;
  using ControlVector = VectorN<m, T>;

 public:
  KnotPoint() : StateControlSized<n, m>(n, m), x_(StateVector::Zero()), u_(ControlVector::Zero()) {}
  KnotPoint(const StateVector& x, const ControlVector& u, float t = 0.0F, float h = 0.0F)
      : StateControlSized<n, m>(static_cast<int>(x.size()), static_cast<int>(u.size())), x_(x), u_(u), t_(t), h_(h) {}
  KnotPoint(int state_dim, int control_dim)
      : StateControlSized<n, m>(state_dim, control_dim), x_(StateVector::Zero(state_dim)),
        u_(ControlVector::Zero(control_dim)) {}
  template <int n2, int m2>
  KnotPoint(const KnotPoint<n2, m2>& z)  // NOLINT: sizes may be given at run time on either side
      : StateControlSized<n, m>(z.StateDimension(), z.ControlDimension()), x_(z.State()), u_(z.Control()),
        t_(z.GetTime()), h_(z.GetStep()) {}

  static KnotPoint Random() {
    ALTRO_ASSERT(n > 0 && m > 0, "Must pass in size if state or control dimension is unknown at compile time.");
    return Random(n, m);
  }
  static KnotPoint Random(int state_dim, int control_dim) {
    const StateVector x = StateVector::Random(state_dim);
    const ControlVector u = ControlVector::Random(control_dim);
    return KnotPoint(x, u, static_cast<float>(std::rand() % 1000) / 100.0F, static_cast<float>(std::rand() % 100) / 100.0F);
  }

  StateVector& State() { return x_; }
  ControlVector& Control() { return u_; }
  const StateVector& State() const { return x_; }
  const ControlVector& Control() const { return u_; }
  VectorN<AddSizes(n, m), T> GetStateControl() const {
    VectorN<AddSizes(n, m), T> z(this->n_ + this->m_);
    for (int i = 0; i < this->n_; ++i) z(i) = x_(i);
    for (int j = 0; j < this->m_; ++j) z(this->n_ + j) = u_(j);
    return z;
  }
  float GetTime() const { return t_; }
  float GetStep() const { return h_; }
  void SetTime(float t) { t_ = t; }
  void SetStep(float h) { h_ = h; }
  bool IsTerminal() const { return h_ == 0; }
  void SetTerminal() {
    h_ = 0;
    u_.setZero();
  }
  std::string ToString(int width = 9) const {
    std::ostringstream os;
    os.precision(3);
    os << "x: [";
    for (int i = 0; i < this->n_; ++i) { os.width(width); os << x_(i) << " "; }
    os << "] u: [";
    for (int j = 0; j < this->m_; ++j) { os.width(width); os << u_(j) << " "; }
    os << "] t=" << t_ << ", h=" << h_;
    return os.str();
  }
  friend std::ostream& operator<<(std::ostream& os, const KnotPoint& z) { return os << z.ToString(); }

 private:
  StateVector x_;
  ControlVector u_;
  float t_ = 0.0F;
  float h_ = 0.0F;
};

}  // namespace altro
