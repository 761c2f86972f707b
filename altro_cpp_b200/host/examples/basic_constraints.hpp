// examples/basic_constraints.hpp (B200 host mirror) — goal equality and control-bound inequality
// (reference: examples/basic_constraints.hpp:15-150); device: device.cuh con_row_fast, al_value,
// al_expansion.
#pragma once

#include <limits>
#include <stdexcept>
#include <vector>

#include "altro/constraints/constraint.hpp"

namespace altro {
namespace examples {

class GoalConstraint : public constraints::Constraint<constraints::Equality> {
 public:
  explicit GoalConstraint(const VectorXd& xf) : xf_(xf) {}
  static constraints::ConstraintPtr<constraints::Equality> Create(const VectorXd& xf) {
    return std::make_shared<GoalConstraint>(xf);
  }
  std::string GetLabel() const override { return "Goal Constraint"; }
  int OutputDimension() const override { return xf_.size(); }
  bool Describe(device::ConstraintDesc* d) const override {
    d->kind = device::ConstraintDesc::kGoal;
    d->a.assign(xf_.data(), xf_.data() + xf_.size());
    return true;
  }

 private:
  VectorXd xf_;
};

class ControlBound : public constraints::Constraint<constraints::Inequality> {
 public:
  explicit ControlBound(int m)
      : lower_bound_(m, -std::numeric_limits<double>::infinity()),
        upper_bound_(m, +std::numeric_limits<double>::infinity()) {}
  ControlBound(const std::vector<double>& lb, const std::vector<double>& ub) : lower_bound_(lb), upper_bound_(ub) {
    if (lb.size() != ub.size() || lb.empty())
      throw std::invalid_argument("Upper and lower bounds must have the same, non-zero length.");
    for (size_t i = 0; i < lb.size(); ++i)
      if (lb[i] > ub[i]) throw std::invalid_argument("Lower bound isn't less than the upper bound.");
  }
  void SetUpperBound(const std::vector<double>& ub) { upper_bound_ = ub; }
  void SetLowerBound(const std::vector<double>& lb) { lower_bound_ = lb; }
  std::string GetLabel() const override { return "Control Bound"; }
  int OutputDimension() const override {  // only finite bounds produce rows (basic_constraints.hpp:136-143)
    int p = 0;
    for (double v : lower_bound_) p += std::abs(v) < std::numeric_limits<double>::max();
    for (double v : upper_bound_) p += std::abs(v) < std::numeric_limits<double>::max();
    return p;
  }
  bool Describe(device::ConstraintDesc* d) const override {
    d->kind = device::ConstraintDesc::kControlBound;
    d->a = lower_bound_;
    d->b = upper_bound_;
    return true;
  }

 private:
  std::vector<double> lower_bound_, upper_bound_;
};

}  // namespace examples
}  // namespace altro
