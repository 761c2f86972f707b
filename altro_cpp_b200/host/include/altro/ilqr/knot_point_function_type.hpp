// altro/ilqr/knot_point_function_type.hpp (B200 host mirror) — what GetKnotPointFunction(k) exposes
// (altro/ilqr/knot_point_function_type.hpp:243-268 there): gains, cost-to-go, cost / dynamics /
// action-value expansions of one knot.  In the reference this object also COMPUTES them
// (CalcActionValueExpansion, CalcGains, CalcCostToGo :149-235); here that arithmetic is
// riccati_step in csrc/device.cuh and this is the host copy of its results, refreshed from the
// device by iLQR::GetKnotPointFunction.
#pragma once

#include <memory>

#include "altro/eigentypes.hpp"
#include "altro/ilqr/cost_expansion.hpp"
#include "altro/ilqr/dynamics_expansion.hpp"
#include "altro/problem/costfunction.hpp"
#include "altro/problem/dynamics.hpp"

namespace altro {
namespace ilqr {

template <int n, int m>
class KnotPointFunctions {
 public:
  KnotPointFunctions(std::shared_ptr<problem::DiscreteDynamics> dynamics, std::shared_ptr<problem::CostFunction> costfun)
      : model_(std::move(dynamics)), costfun_(std::move(costfun)), n_(model_->StateDimension()),
        m_(model_->ControlDimension()), cost_(n_, m_), dyn_(n_, m_), action_(n_, m_), K_(MatrixXd::Zero(m_, n_)),
        d_(VectorXd::Zero(m_)), P_(MatrixXd::Zero(n_, n_)), p_(VectorXd::Zero(n_)) {}

  std::shared_ptr<problem::DiscreteDynamics> GetModelPtr() { return model_; }
  std::shared_ptr<problem::CostFunction> GetCostFunPtr() { return costfun_; }

  CostExpansion<n, m>& GetCostExpansion() { return cost_; }
  DynamicsExpansion<n, m>& GetDynamicsExpansion() { return dyn_; }
  CostExpansion<n, m>& GetActionValueExpansion() { return action_; }
  MatrixXd& GetCostToGoHessian() { return P_; }
  VectorXd& GetCostToGoGradient() { return p_; }
  MatrixXd& GetFeedbackGain() { return K_; }
  VectorXd& GetFeedforwardGain() { return d_; }
  int StateDimension() const { return n_; }
  int ControlDimension() const { return m_; }

 private:
  std::shared_ptr<problem::DiscreteDynamics> model_;
  std::shared_ptr<problem::CostFunction> costfun_;
  int n_, m_;
  CostExpansion<n, m> cost_;
  DynamicsExpansion<n, m> dyn_;
  CostExpansion<n, m> action_;
  MatrixXd K_;
  VectorXd d_;
  MatrixXd P_;
  VectorXd p_;
};

}  // namespace ilqr
}  // namespace altro
