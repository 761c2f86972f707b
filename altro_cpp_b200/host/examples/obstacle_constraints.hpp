// examples/obstacle_constraints.hpp (B200 host mirror) — keep-out circles (reference:
// examples/obstacle_constraints.hpp:14-126).
#pragma once

#include <cmath>
#include <utility>
#include <vector>

#include "altro/constraints/constraint.hpp"

namespace altro {
namespace examples {

struct Circle {
  Circle(double px, double py, double radius) : x(px), y(py), r(radius) {}
  double x, y, r;
  double Distance(double px, double py) const { return std::sqrt((px - x) * (px - x) + (py - y) * (py - y)) - r; }
};

class CircleConstraint : public constraints::Constraint<constraints::Inequality> {
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
  bool Describe(device::ConstraintDesc* d) const override {
    d->kind = device::ConstraintDesc::kCircle;
    d->a.clear();
    d->b.clear();
    d->c.clear();
    for (const Circle& o : obstacles_) {
      d->a.push_back(o.x);
      d->b.push_back(o.y);
      d->c.push_back(o.r);
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
