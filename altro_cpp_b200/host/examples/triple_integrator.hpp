// examples/triple_integrator.hpp (B200 host mirror) — chain of three integrators per degree of
// freedom (reference: examples/triple_integrator.cpp:9-33); device functor `TripleIntegrator<dof>`.
#pragma once

#include "altro/problem/dynamics.hpp"

namespace altro {
namespace examples {

class TripleIntegrator : public problem::ContinuousDynamics {
 public:
  explicit TripleIntegrator(int dof = 1) : dof_(dof) {}
  int StateDimension() const override { return 3 * dof_; }
  int ControlDimension() const override { return dof_; }
  bool HasHessian() const override { return true; }
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
