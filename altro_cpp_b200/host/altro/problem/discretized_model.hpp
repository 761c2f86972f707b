// altro/problem/discretized_model.hpp (B200 host mirror) — DiscretizedModel<Model, RungeKutta4>
// (reference: altro/problem/discretized_model.hpp:25, integration.hpp:113).  The RK4 value and
// its analytic Jacobian are device functions (altro_cpp_b200/csrc/device.cuh rk4_step,
// rk4_jacobian); this class only forwards the continuous model's descriptor.
#pragma once

#include <memory>

#include "altro/problem/dynamics.hpp"

namespace altro {
namespace problem {

struct RungeKutta4Tag {};

template <class Model, class Integrator = RungeKutta4Tag>
class DiscretizedModel : public DiscreteDynamics {
 public:
  explicit DiscretizedModel(const Model& model) : model_(std::make_shared<Model>(model)) {}
  int StateDimension() const override { return model_->StateDimension(); }
  int ControlDimension() const override { return model_->ControlDimension(); }
  bool HasHessian() const override { return model_->HasHessian(); }
  bool Describe(device::ModelDesc* d) const override { return model_->Describe(d); }

 private:
  std::shared_ptr<Model> model_;
};

}  // namespace problem
}  // namespace altro
