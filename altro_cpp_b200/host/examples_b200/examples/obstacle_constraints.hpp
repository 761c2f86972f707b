// examples/obstacle_constraints.hpp (shipped with the B200 host mirror) — stay outside a set of
// circles: c_i = r_i^2 - |p - c_i|^2 <= 0 (examples/obstacle_constraints.hpp:14-126 there).
#pragma once

#include <cmath>
#include <utility>
#include <vector>

#include "altro/constraints/constraint.hpp"
#include "altro/device_registry.hpp"

namespace altro {
namespace examples {

struct Circle {
  Circle(double px, double py, double radius) : x(px), y(py), r(radius) {}
  double x, y, r;
  double Distance2(double px, double py) const { return (px - x) * (px - x) + (py - y) * (py - y) - r * r; }
  double Distance(double px, double py) const { return std::sqrt((px - x) * (px - x) + (py - y) * (py - y)) - r; }
};

class CircleConstraint : public constraints::Constraint<constraints::Inequality>, public device::Describable {
 public:
  template <class... Args>
  void AddObstacle(Args&&... args) {
    obstacles_.emplace_back(std::forward<Args>(args)...);
  }
  void SetXYIndices(int x_index, int y_index) {
    x_index_ = x_index;
    y_index_ = y_index;
  }
  std::string GetLabel() const override { return "Circle Constraint"; }
  int OutputDimension() const override { return static_cast<int>(obstacles_.size()); }
  void Evaluate(const VectorXdRef& x, const VectorXdRef&, Eigen::Ref<VectorXd> c) override {
    for (size_t i = 0; i < obstacles_.size(); ++i) c(i) = -obstacles_[i].Distance2(x(x_index_), x(y_index_));
  }
  // the derivative with respect to the position goes to columns 0 and 1, as in the reference
  void Jacobian(const VectorXdRef& x, const VectorXdRef&, Eigen::Ref<MatrixXd> jac) override {
    for (size_t i = 0; i < obstacles_.size(); ++i) {
      jac(i, 0) = 2 * (obstacles_[i].x - x(x_index_));
      jac(i, 1) = 2 * (obstacles_[i].y - x(y_index_));
    }
  }
  bool Describe(device::ConstraintDesc* d) const override {
    d->kind = device::ConstraintDesc::kCircle;
    d->a.clear();
    d->b.clear();
    d->c.clear();
    for (const Circle& o : obstacles_) {
      d->a.push_back(o.x);
      d->b.push_back(o.y);
      d->c.push_back(o.r * o.r);
    }
    d->xi = x_index_;
    d->yi = y_index_;
    return true;
  }

 private:
  int x_index_ = 0, y_index_ = 1;
  std::vector<Circle> obstacles_;
};

}  // namespace examples
}  // namespace altro
