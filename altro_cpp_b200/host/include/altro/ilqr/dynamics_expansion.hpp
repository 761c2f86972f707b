// altro/ilqr/dynamics_expansion.hpp (B200 host mirror) — discrete dynamics Jacobian [A | B]
// (altro/ilqr/dynamics_expansion.hpp:19 there); host copy of the device record's A, B fields.
#pragma once

#include "altro/common/state_control_sized.hpp"
#include "altro/eigentypes.hpp"

namespace altro {
namespace ilqr {

template <int n, int m>
class DynamicsExpansion : public StateControlSized<n, m> {
 public:
  DynamicsExpansion(int state_dim, int control_dim)
      : StateControlSized<n, m>(state_dim, control_dim), jac_(MatrixXd::Zero(state_dim, state_dim + control_dim)) {}
  MatrixXd& GetJacobian() { return jac_; }
  const MatrixXd& GetJacobian() const { return jac_; }
  MatrixXd GetA() const { return MatrixXd(const_cast<MatrixXd&>(jac_).topLeftCorner(this->n_, this->n_)); }
  MatrixXd GetB() const { return MatrixXd(const_cast<MatrixXd&>(jac_).topRightCorner(this->n_, this->m_)); }

 private:
  MatrixXd jac_;
};

}  // namespace ilqr
}  // namespace altro
