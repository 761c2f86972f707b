// altro/ilqr/dynamics_expansion.hpp (B200 host mirror) — discrete dynamics Jacobian [A | B]
// (altro/ilqr/dynamics_expansion.hpp:19 there); host copy of the device record's A, B fields.
#pragma once

#include <memory>
#include <stdexcept>

#include "altro/common/knotpoint.hpp"
#include "altro/common/state_control_sized.hpp"
#include "altro/eigentypes.hpp"
#include "altro/problem/dynamics.hpp"

namespace altro {
namespace ilqr {

template <int n, int m>
class DynamicsExpansion : public StateControlSized<n, m> {
 public:
  DynamicsExpansion(int state_dim, int control_dim)
      : StateControlSized<n, m>(state_dim, control_dim), jac_(MatrixXd::Zero(state_dim, state_dim + control_dim)) {}
  MatrixXd& GetJacobian() { return jac_; }
  const MatrixXd& GetJacobian() const { return jac_; }
  void SetZero() { jac_.setZero(); }
  // writable views of the two blocks (assigning to them fills the Jacobian)
  Eigen::Ref<MatrixXd> GetA() { return jac_.topLeftCorner(this->n_, this->n_); }
  Eigen::Ref<MatrixXd> GetB() { return jac_.topRightCorner(this->n_, this->m_); }
  MatrixXd GetA() const { return MatrixXd(const_cast<MatrixXd&>(jac_).topLeftCorner(this->n_, this->n_)); }
  MatrixXd GetB() const { return MatrixXd(const_cast<MatrixXd&>(jac_).topRightCorner(this->n_, this->m_)); }
  // Host-side evaluation through the model's virtual Jacobian (debugging aid; the solve differentiates on the
  // device).  Only discrete dynamics have a step to expand: anything else is refused.
  void CalcExpansion(const std::shared_ptr<FunctionBase>& model, const VectorXdRef& x, const VectorXdRef& u, float t, float h) {
    const std::shared_ptr<problem::DiscreteDynamics> discrete = std::dynamic_pointer_cast<problem::DiscreteDynamics>(model);
    if (!discrete) throw std::runtime_error("DynamicsExpansion::CalcExpansion needs a discrete dynamics model.");
    discrete->Jacobian(x, u, t, h, jac_);
  }
  template <class Model, int n2, int m2, class T>
  void CalcExpansion(const std::shared_ptr<Model>& model, const KnotPoint<n2, m2, T>& z) {
    CalcExpansion(std::static_pointer_cast<FunctionBase>(model), z.State(), z.Control(), z.GetTime(), z.GetStep());
  }

 private:
  MatrixXd jac_;
};

}  // namespace ilqr
}  // namespace altro
