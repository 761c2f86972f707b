// altro/ilqr/cost_expansion.hpp (B200 host mirror) — second-order expansion of a scalar function of
// (x, u) (altro/ilqr/cost_expansion.hpp:27 there): dxdx, dxdu, dudu, dx, du.  On the device these
// are fields of the per-knot record (csrc/kernels.cuh k_update_expansions); this is the host copy
// GetKnotPointFunction(k) hands out.
#pragma once

#include <memory>

#include "altro/common/knotpoint.hpp"
#include "altro/common/state_control_sized.hpp"
#include "altro/eigentypes.hpp"
#include "altro/problem/costfunction.hpp"

namespace altro {
namespace ilqr {

template <int n, int m>
class CostExpansion : public StateControlSized<n, m> {
 public:
  CostExpansion(int state_dim, int control_dim)
      : StateControlSized<n, m>(state_dim, control_dim), xx_(MatrixXd::Zero(state_dim, state_dim)),
        xu_(MatrixXd::Zero(state_dim, control_dim)), uu_(MatrixXd::Zero(control_dim, control_dim)),
        x_(VectorXd::Zero(state_dim)), u_(VectorXd::Zero(control_dim)) {}
  MatrixXd& dxdx() { return xx_; }
  MatrixXd& dxdu() { return xu_; }
  MatrixXd& dudu() { return uu_; }
  VectorXd& dx() { return x_; }
  VectorXd& du() { return u_; }
  const MatrixXd& dxdx() const { return xx_; }
  const MatrixXd& dxdu() const { return xu_; }
  const MatrixXd& dudu() const { return uu_; }
  const VectorXd& dx() const { return x_; }
  const VectorXd& du() const { return u_; }
  // joint (n+m) x (n+m) Hessian and (n+m) gradient
  MatrixXd GetExpansion() const {
    const int N = this->n_, M = this->m_;
    MatrixXd H = MatrixXd::Zero(N + M, N + M);
    H.topLeftCorner(N, N) = xx_;
    H.topRightCorner(N, M) = xu_;
    H.bottomLeftCorner(M, N) = xu_.transpose();
    H.bottomRightCorner(M, M) = uu_;
    return H;
  }
  VectorXd GetGradient() const {
    VectorXd g = VectorXd::Zero(this->n_ + this->m_);
    g.head(this->n_) = x_;
    g.tail(this->m_) = u_;
    return g;
  }
  void SetZero() {
    xx_.setZero(); xu_.setZero(); uu_.setZero(); x_.setZero(); u_.setZero();
  }
  // Host-side evaluation through the functor's virtuals (what the reference's solver does every iteration; here a
  // debugging aid — the solve expands on the device, k_update_expansions)
  void CalcExpansion(const std::shared_ptr<problem::CostFunction>& costfun, const VectorXdRef& x, const VectorXdRef& u) {
    costfun->Gradient(x, u, x_, u_);
    costfun->Hessian(x, u, xx_, xu_, uu_);
  }
  template <int n2, int m2, class T>
  void CalcExpansion(const std::shared_ptr<problem::CostFunction>& costfun, const KnotPoint<n2, m2, T>& z) {
    CalcExpansion(costfun, z.State(), z.Control());
  }

 private:
  MatrixXd xx_, xu_, uu_;
  VectorXd x_, u_;
};

}  // namespace ilqr
}  // namespace altro
