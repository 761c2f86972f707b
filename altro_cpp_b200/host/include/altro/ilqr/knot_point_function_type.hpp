// altro/ilqr/knot_point_function_type.hpp (B200 host mirror) — KnotPointFunctions<n, m>, what
// iLQR::GetKnotPointFunction(k) exposes (altro/ilqr/knot_point_function_type.hpp:243-268 there): gains, cost-to-go,
// cost / dynamics / action-value expansions of one knot point.
//
// In a solve these are RESULTS: riccati_step / update_expansions in csrc/device.cuh compute them for every knot
// and instance, and iLQR::GetKnotPointFunction refreshes this host copy from the device.  The Calc* members below
// are the reference's public single-knot arithmetic (:107-230 there) restated on the host for callers who drive
// one knot point by hand (its unit tests do); no solver method of this repo calls them.
#pragma once

#include <array>
#include <cmath>
#include <memory>
#include <utility>

#include "altro/eigentypes.hpp"
#include "altro/ilqr/cost_expansion.hpp"
#include "altro/ilqr/dynamics_expansion.hpp"
#include "altro/problem/costfunction.hpp"
#include "altro/problem/dynamics.hpp"
#include "altro/utils/assert.hpp"

namespace altro {
namespace ilqr {

// where the backward-pass regularisation is added; the device path has the same single mode (rho on Quu)
enum class BackwardPassRegularization { kControlOnly };

template <int n, int m>
class KnotPointFunctions {
  using DynamicsPtr = std::shared_ptr<problem::DiscreteDynamics>;
  using CostFunPtr = std::shared_ptr<problem::CostFunction>;

 public:
  KnotPointFunctions(DynamicsPtr dynamics, CostFunPtr costfun)
      : model_(Checked(std::move(dynamics))), costfun_(Checked(std::move(costfun))), n_(model_->StateDimension()),
        m_(model_->ControlDimension()), cost_(n_, m_), dyn_(n_, m_), action_(n_, m_), action_reg_(n_, m_),
        K_(MatrixXd::Zero(m_, n_)), d_(VectorXd::Zero(m_)), P_(MatrixXd::Zero(n_, n_)), p_(VectorXd::Zero(n_)) {}
  // the terminal knot point of a problem without a model there
  KnotPointFunctions(int state_dim, int control_dim, CostFunPtr costfun)
      : costfun_(Checked(std::move(costfun))), n_(state_dim), m_(control_dim), cost_(n_, m_), dyn_(n_, m_),
        action_(n_, m_), action_reg_(n_, m_), K_(MatrixXd::Zero(m_, n_)), d_(VectorXd::Zero(m_)),
        P_(MatrixXd::Zero(n_, n_)), p_(VectorXd::Zero(n_)) {}

  DynamicsPtr GetModelPtr() { return model_; }
  CostFunPtr GetCostFunPtr() { return costfun_; }
  int StateDimension() const { return n_; }
  int ControlDimension() const { return m_; }

  // ---- results (of the device after a solver phase, or of the Calc* calls below)
  CostExpansion<n, m>& GetCostExpansion() { return cost_; }
  DynamicsExpansion<n, m>& GetDynamicsExpansion() { return dyn_; }
  CostExpansion<n, m>& GetActionValueExpansion() { return action_; }
  CostExpansion<n, m>& GetActionValueExpansionRegularized() { return action_reg_; }
  MatrixXd& GetCostToGoHessian() { return P_; }
  VectorXd& GetCostToGoGradient() { return p_; }
  MatrixXd& GetFeedbackGain() { return K_; }
  VectorXd& GetFeedforwardGain() { return d_; }
  // expected change of the cost-to-go for a step of length alpha: alpha d'Qu + alpha^2 d'Quu d / 2
  double GetCostToGoDelta(const double alpha = 1.0) { return alpha * dV_[0] + alpha * alpha * dV_[1]; }
  void AddCostToGo(std::array<double, 2>* const deltaV) const { AddCostToGo(deltaV->data()); }
  void AddCostToGo(double* const deltaV) const {
    deltaV[0] += dV_[0];
    deltaV[1] += dV_[1];
  }

  // ---- single-knot evaluation through the user's functors
  double Cost(const VectorXdRef& x, const VectorXdRef& u) const { return costfun_->Evaluate(x, u); }
  void Dynamics(const VectorXdRef& x, const VectorXdRef& u, float t, float h, Eigen::Ref<VectorXd> xnext) const {
    model_->Evaluate(x, u, t, h, xnext);
  }
  void CalcCostExpansion(const VectorXdRef& x, const VectorXdRef& u) {
    cost_.SetZero();
    cost_.CalcExpansion(costfun_, x, u);
  }
  void CalcDynamicsExpansion(const VectorXdRef& x, const VectorXdRef& u, const float t, const float h) {
    if (!model_) return;
    dyn_.SetZero();
    dyn_.CalcExpansion(model_, x, u, t, h);
  }

  // ---- one step of the Riccati recursion, from the expansions held above
  void CalcTerminalCostToGo() {
    P_ = cost_.dxdx();
    p_ = cost_.dx();
  }
  // Q = l + [A B]' S [A B], S / s the cost-to-go of the next knot point
  void CalcActionValueExpansion(const Eigen::Ref<const MatrixXd>& ctg_hessian, const Eigen::Ref<const MatrixXd>& ctg_gradient) {
    const MatrixXd& J = dyn_.GetJacobian();  // [A B], n x (n + m)
    MatrixXd SJ = MatrixXd::Zero(n_, n_ + m_);
    for (int i = 0; i < n_; ++i)
      for (int j = 0; j < n_ + m_; ++j) {
        double acc = 0.0;
        for (int r = 0; r < n_; ++r) acc += ctg_hessian(i, r) * J(r, j);
        SJ(i, j) = acc;
      }
    for (int a = 0; a < n_ + m_; ++a) {
      double g = 0.0;
      for (int r = 0; r < n_; ++r) g += J(r, a) * ctg_gradient(r, 0);
      if (a < n_) action_.dx()(a) = cost_.dx()(a) + g; else action_.du()(a - n_) = cost_.du()(a - n_) + g;
      for (int b = a < n_ ? 0 : n_; b < n_ + m_; ++b) {
        double h = 0.0;
        for (int r = 0; r < n_; ++r) h += J(r, a) * SJ(r, b);
        if (a < n_ && b < n_) action_.dxdx()(a, b) = cost_.dxdx()(a, b) + h;
        else if (a < n_) action_.dxdu()(a, b - n_) = cost_.dxdu()(a, b - n_) + h;
        else action_.dudu()(a - n_, b - n_) = cost_.dudu()(a - n_, b - n_) + h;
      }
    }
  }
  void RegularizeActionValue(const double rho, BackwardPassRegularization reg_type = BackwardPassRegularization::kControlOnly) {
    action_reg_ = action_;
    if (reg_type == BackwardPassRegularization::kControlOnly)
      for (int i = 0; i < m_; ++i) action_reg_.dudu()(i, i) += rho;
  }
  // K = -Quu^-1 Qux, d = -Quu^-1 Qu by Cholesky of the regularised Quu; gains untouched when it fails
  Eigen::ComputationInfo CalcGains() {
    Eigen::LLT<MatrixXd> chol;
    chol.compute(action_reg_.dudu());
    if (chol.info() != Eigen::Success) return chol.info();
    MatrixXd Qux = MatrixXd::Zero(m_, n_);
    for (int i = 0; i < m_; ++i)
      for (int j = 0; j < n_; ++j) Qux(i, j) = action_reg_.dxdu()(j, i);
    K_ = chol.solve(Qux);
    K_ *= -1;
    d_ = chol.solve(action_reg_.du());
    d_ *= -1;
    return Eigen::Success;
  }
  // S = Qxx + K'Quu K + K'Qux + Qxu K,  s = Qx + K'Quu d + K'Qu + Qxu d,  and the two parts of the expected change
  void CalcCostToGo() {
    const MatrixXd& Qxx = action_.dxdx();
    const MatrixXd& Qxu = action_.dxdu();
    const MatrixXd& Quu = action_.dudu();
    const VectorXd& Qx = action_.dx();
    const VectorXd& Qu = action_.du();
    MatrixXd QuuK = MatrixXd::Zero(m_, n_);
    VectorXd Quud = VectorXd::Zero(m_);
    for (int i = 0; i < m_; ++i) {
      for (int j = 0; j < n_; ++j) {
        double acc = 0.0;
        for (int r = 0; r < m_; ++r) acc += Quu(i, r) * K_(r, j);
        QuuK(i, j) = acc;
      }
      double acc = 0.0;
      for (int r = 0; r < m_; ++r) acc += Quu(i, r) * d_(r);
      Quud(i) = acc;
    }
    for (int i = 0; i < n_; ++i) {
      double g = Qx(i);
      for (int r = 0; r < m_; ++r) g += K_(r, i) * (Quud(r) + Qu(r)) + Qxu(i, r) * d_(r);
      p_(i) = g;
      for (int j = 0; j < n_; ++j) {
        double h = Qxx(i, j);
        for (int r = 0; r < m_; ++r) h += K_(r, i) * (QuuK(r, j) + Qxu(j, r)) + Qxu(i, r) * K_(r, j);
        P_(i, j) = h;
      }
    }
    dV_[0] = 0.0;
    dV_[1] = 0.0;
    for (int r = 0; r < m_; ++r) {
      dV_[0] += d_(r) * Qu(r);
      dV_[1] += d_(r) * Quud(r);
    }
    dV_[1] *= 0.5;
  }

 private:
  static DynamicsPtr Checked(DynamicsPtr dynamics) {
    ALTRO_ASSERT(dynamics != nullptr, "Cannot provide a null dynamics pointer.");
    return dynamics;
  }
  static CostFunPtr Checked(CostFunPtr costfun) {
    ALTRO_ASSERT(costfun != nullptr, "Cannot provide a null cost function pointer.");
    return costfun;
  }

  DynamicsPtr model_;
  CostFunPtr costfun_;
  int n_, m_;
  CostExpansion<n, m> cost_;
  DynamicsExpansion<n, m> dyn_;
  CostExpansion<n, m> action_;
  CostExpansion<n, m> action_reg_;
  MatrixXd K_;
  VectorXd d_;
  MatrixXd P_;
  VectorXd p_;
  std::array<double, 2> dV_ = {{0.0, 0.0}};
};

}  // namespace ilqr
}  // namespace altro
