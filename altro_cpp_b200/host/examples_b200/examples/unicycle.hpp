// examples/unicycle.hpp (shipped with the B200 host mirror) — the unicycle of the reference's
// examples (examples/unicycle.cpp:12-33 there): xdot = [v cos(theta), v sin(theta), omega].  A host
// functor with the plug-in interface of altro/problem/dynamics.hpp that also tells the solver which
// device functor it is (csrc/device.cuh `Unicycle`).
#pragma once

#include <cmath>

#include "altro/device_registry.hpp"
#include "altro/problem/dynamics.hpp"

namespace altro {
namespace examples {

class Unicycle : public problem::ContinuousDynamics, public device::Describable {
 public:
  static constexpr int NStates = 3;
  static constexpr int NControls = 2;
  int StateDimension() const override { return NStates; }
  int ControlDimension() const override { return NControls; }
  bool HasHessian() const override { return true; }

  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, const float, Eigen::Ref<VectorXd> xdot) override {
    xdot(0) = u(0) * std::cos(x(2));
    xdot(1) = u(0) * std::sin(x(2));
    xdot(2) = u(1);
  }
  // only the five structurally non-zero entries are written: the caller provides a zeroed matrix
  void Jacobian(const VectorXdRef& x, const VectorXdRef& u, const float, Eigen::Ref<MatrixXd> jac) override {
    const double s = std::sin(x(2)), c = std::cos(x(2));
    jac(0, 2) = -u(0) * s;
    jac(1, 2) = u(0) * c;
    jac(0, 3) = c;
    jac(1, 3) = s;
    jac(2, 4) = 1;
  }
  void Hessian(const VectorXdRef& x, const VectorXdRef& u, const float, const VectorXdRef& b,
               Eigen::Ref<MatrixXd> hess) override {
    const double s = std::sin(x(2)), c = std::cos(x(2));
    hess(2, 2) = -b(0) * u(0) * c - b(1) * u(0) * s;
    hess(2, 3) = -b(0) * s + b(1) * c;
    hess(3, 2) = hess(2, 3);
  }
  bool Describe(device::ModelDesc* d) const override {
    d->model = ALTRO_B200_MODEL_UNICYCLE;
    d->params.clear();
    return true;
  }
};

}  // namespace examples
}  // namespace altro
