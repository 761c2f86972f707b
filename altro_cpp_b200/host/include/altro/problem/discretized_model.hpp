// altro/problem/discretized_model.hpp (B200 host mirror) — a continuous model + an explicit
// integrator = DiscreteDynamics (altro/problem/discretized_model.hpp:24 there).  On the device the
// pair DiscretizedModel<Model, RungeKutta4> is rk4_step / rk4_jacobian<Model> (csrc/device.cuh);
// altro/device_registry.hpp recognises which registered Model it wraps.
#pragma once

#include <memory>
#include <type_traits>

#include "altro/problem/integration.hpp"

namespace altro {
namespace problem {

// type-erased access to the wrapped continuous model (the registry asks for it; the reference's
// own class keeps it private)
class DiscretizedModelBase {
 public:
  virtual ~DiscretizedModelBase() = default;
  virtual std::shared_ptr<ContinuousDynamics> GetContinuousModel() const = 0;
  virtual bool IsRungeKutta4() const = 0;
};

// RungeKutta4 with any compile-time sizes (the reference instantiates it with the model's sizes or with
// the problem's, examples/problems/triple_integrator.hpp:43 there)
template <class T>
struct IsRk4 : std::false_type {};
template <int NStates, int NControls>
struct IsRk4<RungeKutta4<NStates, NControls>> : std::true_type {};

template <class Model, class Integrator = RungeKutta4<Model::NStates, Model::NControls>>
class DiscretizedModel : public DiscreteDynamics, public DiscretizedModelBase {
 public:
  static_assert(std::is_base_of<FunctionBase, Model>::value, "Model must inherit from FunctionBase.");
  using DiscreteDynamics::Evaluate;

  static constexpr int NStates = Model::NStates;
  static constexpr int NControls = Model::NControls;

  explicit DiscretizedModel(const Model& model)
      : model_(std::make_shared<Model>(model)), integrator_(model.StateDimension(), model.ControlDimension()) {}

  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, const float t, const float h,
                Eigen::Ref<VectorXd> xnext) override {
    integrator_.Integrate(model_, x, u, t, h, xnext);
  }
  void Jacobian(const VectorXdRef& x, const VectorXdRef& u, const float t, const float h,
                Eigen::Ref<MatrixXd> jac) override {
    integrator_.Jacobian(model_, x, u, t, h, jac);
  }
  void Hessian(const VectorXdRef& x, const VectorXdRef& u, const float t, const float h, const VectorXdRef& b,
               Eigen::Ref<MatrixXd> hess) override {
    ALTRO_UNUSED(x);
    ALTRO_UNUSED(u);
    ALTRO_UNUSED(t);
    ALTRO_UNUSED(h);
    ALTRO_UNUSED(b);
    ALTRO_UNUSED(hess);
  }
  bool HasHessian() const override { return model_->HasHessian(); }
  int StateDimension() const override { return model_->StateDimension(); }
  int ControlDimension() const override { return model_->ControlDimension(); }
  Integrator& GetIntegrator() { return integrator_; }

  std::shared_ptr<ContinuousDynamics> GetContinuousModel() const override { return model_; }
  bool IsRungeKutta4() const override { return IsRk4<Integrator>::value; }

 private:
  std::shared_ptr<Model> model_;
  Integrator integrator_;
};

}  // namespace problem
}  // namespace altro
