// examples/triple_integrator.hpp (shipped with the B200 host mirror) — three integrators per degree
// of freedom driven by the jerk (examples/triple_integrator.cpp:9-33 there); device functor
// `TripleIntegrator<dof>` in csrc/device.cuh.  State [positions, velocities, accelerations].
#pragma once

#include "altro/device_registry.hpp"
#include "altro/problem/dynamics.hpp"

namespace altro {
namespace examples {

class TripleIntegrator : public problem::ContinuousDynamics, public device::Describable {
 public:
  using ContinuousDynamics::Evaluate;
  explicit TripleIntegrator(int dof = 1) : dof_(dof) { ALTRO_ASSERT(dof > 0, "The degrees of freedom must be greater than 0."); }
  int StateDimension() const override { return 3 * dof_; }
  int ControlDimension() const override { return dof_; }
  bool HasHessian() const override { return true; }

  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, float, Eigen::Ref<VectorXd> xdot) override {
    for (int i = 0; i < 2 * dof_; ++i) xdot(i) = x(i + dof_);
    for (int i = 0; i < dof_; ++i) xdot(2 * dof_ + i) = u(i);
  }
  void Jacobian(const VectorXdRef&, const VectorXdRef&, float, Eigen::Ref<MatrixXd> jac) override {
    jac.setZero();
    for (int i = 0; i < 3 * dof_; ++i) jac(i, i + dof_) = 1;  // d xdot_i / d (next block of [x | u])
  }
  void Hessian(const VectorXdRef&, const VectorXdRef&, float, const VectorXdRef&, Eigen::Ref<MatrixXd> hess) override {
    hess.setZero();
  }
  bool Describe(device::ModelDesc* d) const override {
    d->model = ALTRO_B200_MODEL_TRIPLE_INTEGRATOR;
    d->params.clear();
    return true;
  }

 private:
  int dof_;
};

}  // namespace examples
}  // namespace altro
