// examples/basic_constraints.hpp (shipped with the B200 host mirror) — goal equality x = xf and
// control bounds lb <= u <= ub (examples/basic_constraints.hpp:15-150 there); on the device:
// csrc/device.cuh for_row_chunks / al_value / al_expansion.
#pragma once

#include <cmath>
#include <limits>
#include <stdexcept>
#include <vector>

#include "altro/constraints/constraint.hpp"
#include "altro/device_registry.hpp"

namespace altro {
namespace examples {

class GoalConstraint : public constraints::Constraint<constraints::Equality>, public device::Describable {
 public:
  explicit GoalConstraint(const VectorXd& xf) : xf_(xf) {}
  static constraints::ConstraintPtr<constraints::Equality> Create(const VectorXd& xf) {
    return std::make_shared<GoalConstraint>(xf);
  }
  std::string GetLabel() const override { return "Goal Constraint"; }
  int StateDimension() const override { return static_cast<int>(xf_.size()); }
  int OutputDimension() const override { return static_cast<int>(xf_.size()); }
  void Evaluate(const VectorXdRef& x, const VectorXdRef&, Eigen::Ref<VectorXd> c) override { c = x - xf_; }
  void Jacobian(const VectorXdRef&, const VectorXdRef&, Eigen::Ref<MatrixXd> jac) override { jac.setIdentity(); }
  bool Describe(device::ConstraintDesc* d) const override {
    d->kind = device::ConstraintDesc::kGoal;
    d->a.assign(xf_.data(), xf_.data() + xf_.size());
    return true;
  }

 private:
  VectorXd xf_;
};

// rows: [lb_j - u_j for finite lb_j] then [u_j - ub_j for finite ub_j]
class ControlBound : public constraints::Constraint<constraints::Inequality>, public device::Describable {
 public:
  explicit ControlBound(int m)
      : lower_bound_(m, -std::numeric_limits<double>::infinity()), upper_bound_(m, +std::numeric_limits<double>::infinity()) {}
  ControlBound(const std::vector<double>& lb, const std::vector<double>& ub) : lower_bound_(lb), upper_bound_(ub) {
    if (lb.size() != ub.size() || lb.empty())
      throw std::invalid_argument("Upper and lower bounds must have the same, non-zero length.");
    Validate();
  }
  void SetUpperBound(const std::vector<double>& ub) {
    upper_bound_ = ub;
    Validate();
  }
  void SetLowerBound(const std::vector<double>& lb) {
    lower_bound_ = lb;
    Validate();
  }
  std::string GetLabel() const override { return "Control Bound"; }
  int ControlDimension() const override { return static_cast<int>(lower_bound_.size()); }
  int OutputDimension() const override { return Count(lower_bound_) + Count(upper_bound_); }
  void Evaluate(const VectorXdRef&, const VectorXdRef& u, Eigen::Ref<VectorXd> c) override {
    int row = 0;
    for (size_t j = 0; j < lower_bound_.size(); ++j)
      if (Finite(lower_bound_[j])) c(row++) = lower_bound_[j] - u(j);
    for (size_t j = 0; j < upper_bound_.size(); ++j)
      if (Finite(upper_bound_[j])) c(row++) = u(j) - upper_bound_[j];
  }
  void Jacobian(const VectorXdRef& x, const VectorXdRef&, Eigen::Ref<MatrixXd> jac) override {
    jac.setZero();
    const int n = static_cast<int>(x.size());
    int row = 0;
    for (size_t j = 0; j < lower_bound_.size(); ++j)
      if (Finite(lower_bound_[j])) jac(row++, n + static_cast<int>(j)) = -1;
    for (size_t j = 0; j < upper_bound_.size(); ++j)
      if (Finite(upper_bound_[j])) jac(row++, n + static_cast<int>(j)) = 1;
  }
  bool Describe(device::ConstraintDesc* d) const override {
    d->kind = device::ConstraintDesc::kControlBound;
    d->a = lower_bound_;
    d->b = upper_bound_;
    return true;
  }

 private:
  static bool Finite(double v) { return std::abs(v) < std::numeric_limits<double>::max(); }
  static int Count(const std::vector<double>& b) {
    int p = 0;
    for (double v : b) p += Finite(v);
    return p;
  }
  void Validate() const {
    for (size_t i = 0; i < lower_bound_.size() && i < upper_bound_.size(); ++i)
      if (lower_bound_[i] > upper_bound_[i]) throw std::invalid_argument("Lower bound isn't less than the upper bound.");
  }
  std::vector<double> lower_bound_, upper_bound_;
};

}  // namespace examples
}  // namespace altro
