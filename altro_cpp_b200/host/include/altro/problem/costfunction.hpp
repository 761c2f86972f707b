// altro/problem/costfunction.hpp (B200 host mirror) — CostFunction ABC
// (altro/problem/costfunction.hpp:52-73 there): value, split gradient (dx, du), split Hessian
// (dxdx, dxdu, dudu); the joint ScalarFunction interface hands in the matching blocks.
#pragma once

#include <iostream>

#include "altro/common/functionbase.hpp"
#include "altro/eigentypes.hpp"

namespace altro {
namespace problem {

class CostFunction : public altro::ScalarFunction {
 public:
  using altro::ScalarFunction::Evaluate;
  using altro::ScalarFunction::Hessian;

  virtual void Gradient(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> dx, Eigen::Ref<VectorXd> du) = 0;
  virtual void Hessian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> dxdx, Eigen::Ref<MatrixXd> dxdu,
                       Eigen::Ref<MatrixXd> dudu) = 0;

  void Gradient(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> grad) override {
    Gradient(x, u, grad.head(StateDimension()), grad.tail(ControlDimension()));
  }
  void Hessian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> hess) override {
    const int n = StateDimension(), m = ControlDimension();
    Hessian(x, u, hess.topLeftCorner(n, n), hess.topRightCorner(n, m), hess.bottomRightCorner(m, m));
  }
};

}  // namespace problem
}  // namespace altro
