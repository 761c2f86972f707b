// altro/problem/discretized_model.hpp (B200 host mirror) — a continuous model plus an explicit integrator
// is a DiscreteDynamics (altro/problem/discretized_model.hpp:24 there).  On the device the pair
// DiscretizedModel<Model, RungeKutta4> is rk4_step / rk4_jacobian<Model> (csrc/device.cuh);
// altro/device_registry.hpp works out which registered device model `Model` is.
#pragma once

#include <memory>
#include <type_traits>

#include "altro/problem/integration.hpp"

namespace altro {
namespace problem {

// What the registry needs to know about a discretised model without knowing its template arguments.
// (The reference keeps the wrapped model private; there nothing outside ever asks.)
class DiscretizedModelBase {
 public:
  virtual ~DiscretizedModelBase() = default;
  virtual std::shared_ptr<ContinuousDynamics> GetContinuousModel() const = 0;
  virtual bool IsRungeKutta4() const = 0;
};

namespace detail {
// RungeKutta4 of any compile-time size: the reference instantiates it with the model's sizes or with the
// problem's (examples/problems/triple_integrator.hpp:43 there)
template <class Integrator>
struct IsRk4 : std::false_type {};
template <int NStates, int NControls>
struct IsRk4<RungeKutta4<NStates, NControls>> : std::true_type {};
}  // namespace detail
template <class Integrator>
using IsRk4 = detail::IsRk4<Integrator>;

template <class Model, class Integrator = RungeKutta4<Model::NStates, Model::NControls>>
class DiscretizedModel : public DiscreteDynamics, public DiscretizedModelBase {
  static_assert(std::is_base_of<FunctionBase, Model>::value, "Model must inherit from FunctionBase.");
  std::shared_ptr<Model> model_;  // a private copy of the caller's model, shared with the integrator calls
  Integrator integrator_;

 public:
  static constexpr int NStates = Model::NStates;
  static constexpr int NControls = Model::NControls;
  using DiscreteDynamics::Evaluate;

  explicit DiscretizedModel(const Model& model)
      : model_(std::make_shared<Model>(model)), integrator_(model.StateDimension(), model.ControlDimension()) {}

  // ---- sizes and capabilities come from the wrapped model
  int StateDimension() const override { return model_->StateDimension(); }
  int ControlDimension() const override { return model_->ControlDimension(); }
  bool HasHessian() const override { return model_->HasHessian(); }

  // ---- one step x_{k+1} = f(x_k, u_k, t_k, h_k) and its n x (n + m) Jacobian, through the integrator
  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, const float t, const float h,
                Eigen::Ref<VectorXd> xnext) override {
    integrator_.Integrate(model_, x, u, t, h, xnext);
  }
  void Jacobian(const VectorXdRef& x, const VectorXdRef& u, const float t, const float h,
                Eigen::Ref<MatrixXd> jac) override {
    integrator_.Jacobian(model_, x, u, t, h, jac);
  }
  // second-order dynamics terms are not propagated through the integrator (the reference leaves `hess`
  // untouched as well; its DDP terms are stubbed out)
  void Hessian(const VectorXdRef&, const VectorXdRef&, const float, const float, const VectorXdRef&,
               Eigen::Ref<MatrixXd>) override {}

  Integrator& GetIntegrator() { return integrator_; }
  std::shared_ptr<ContinuousDynamics> GetContinuousModel() const override { return model_; }
  bool IsRungeKutta4() const override { return IsRk4<Integrator>::value; }
};

}  // namespace problem
}  // namespace altro
