// altro/constraints/constraint_values.hpp (B200 host mirror) — ConstraintValues<n, m, ConType>: one constraint
// of one knot point together with its multipliers and penalty
// (altro/constraints/constraint_values.hpp:24 there).
//
// Two lives, one type:
//   * standalone (constructed from a constraint pointer, or by ALCost(prob, k)): multipliers, penalty and the
//     augmented-Lagrangian value / gradient / Gauss-Newton Hessian at a point (x, u) the caller passes are kept
//     and computed here, on the host, through the constraint's own virtuals.  This is the user-facing
//     single-point arithmetic of the reference (:113-176 there); no solver runs through it.
//   * bound to a device solver (handed out by AugmentedLagrangianiLQR::GetALCost(k) and by the knot point
//     functions of an AL problem): multipliers, penalty and constraint values of the solver's trajectory live on
//     the device (csrc/device.cuh al_value / al_expansion, LAM, one penalty per instance — SURVEY.md Q9) and every
//     getter reads them from there; SetPenalty and writes through GetDuals() go back to the device before its
//     next phase.
#pragma once

#include <algorithm>
#include <cmath>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "altro/common/state_control_sized.hpp"
#include "altro/constraints/constraint.hpp"
#include "altro/device_solver.hpp"
#include "altro/eigentypes.hpp"

namespace altro {
namespace constraints {

template <int n, int m, class ConType>
class ConstraintValues : public Constraint<ConType> {
  using DualCone = typename ConType::DualCone;

 public:
  static constexpr double kDefaultPenaltyScaling = 10.0;
  static constexpr int NStates = n;
  static constexpr int NControls = m;

  ConstraintValues(int state_dim, int control_dim, ConstraintPtr<ConType> con)
      : n_(state_dim), m_(control_dim), con_(std::move(con)), p_(con_->OutputDimension()),
        lambda_(std::make_shared<VectorXd>(VectorXd::Zero(p_))), c_(VectorXd::Zero(p_)), rho_(VectorXd::Ones(p_)),
        viol_(VectorXd::Zero(p_)), shifted_(VectorXd::Zero(p_)), lambda_bar_(VectorXd::Zero(p_)),
        jac_(MatrixXd::Zero(p_, n_ + m_)), cone_jac_(MatrixXd::Zero(p_, p_)), active_jac_(MatrixXd::Zero(p_, n_ + m_)) {}

  // rows [row0, row0 + p) of knot k of `core` (ALCost row order: equalities, then inequalities)
  void BindDevice(std::shared_ptr<altro::detail::DeviceSolver> core, int k, int row0) {
    core_ = std::move(core);
    k_ = k;
    row0_ = row0;
  }
  bool OnDevice() const { return core_ != nullptr && core_->Ready(); }

  // ---- Constraint<ConType>: forwarded to the wrapped constraint
  int StateDimension() const override { return n_; }
  int ControlDimension() const override { return m_; }
  int OutputDimension() const override { return p_; }
  std::string GetLabel() const override { return con_->GetLabel(); }
  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> c) override { con_->Evaluate(x, u, c); }
  void Jacobian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> jac) override { con_->Jacobian(x, u, jac); }
  ConstraintPtr<ConType> GetConstraint() { return con_; }

  // ---- state
  // Writable.  Bound: starts as the device's multipliers and is sent back right before the next device phase
  // (every instance of a batched solver receives the same values).
  VectorXd& GetDuals() {
    if (OnDevice()) {
      *lambda_ = DeviceRows(core_->Duals(k_, 0));
      core_->PushDualsBeforeNextRun(k_, row0_, lambda_);
    }
    return *lambda_;
  }
  VectorXd& GetPenalty() {
    PullPenalty();
    return rho_;
  }
  // standalone: c at the point of the last evaluation; bound: c(x_k, u_k) of the solver's trajectory
  VectorXd& GetConstraintValue() {
    if (OnDevice()) c_ = DeviceRows(core_->ConstraintValues(k_, 0));
    return c_;
  }
  double GetPenaltyScaling() const { return phi_; }
  // c - Pi_K(c): c itself for an equality, max(0, c) for an inequality
  VectorXd& GetViolation() {
    GetConstraintValue();
    ConType::Projection(c_, viol_);
    for (int i = 0; i < p_; ++i) viol_(i) = c_(i) - viol_(i);
    return viol_;
  }
  ConstraintInfo GetConstraintInfo() { return ConstraintInfo{con_->GetLabel(), k_, GetViolation(), con_->GetConstraintType()}; }

  void SetPenalty(double rho) {
    ALTRO_ASSERT(rho >= 0, "Penalty must be positive.");
    rho_.setConstant(rho);
    if (core_) core_->SetPenalty(rho);  // one penalty per instance on the device
  }
  void SetPenaltyScaling(double phi) {
    ALTRO_ASSERT(phi >= 1, "Penalty must be greater than 1.");
    phi_ = phi;
    if (core_) core_->SetPenaltyScaling(phi);
  }

  // ---- the augmented Lagrangian of this constraint at (x, u):
  //      (|Pi_K*(lambda - rho c)|^2 - |lambda|^2) / (2 rho),   rho = the first row's penalty
  double AugLag(const VectorXdRef& x, const VectorXdRef& u) {
    const double rho = Shift(x, u);
    double proj2 = 0.0, lam2 = 0.0;
    for (int i = 0; i < p_; ++i) {
      proj2 += lambda_bar_(i) * lambda_bar_(i);
      lam2 += (*lambda_)(i) * (*lambda_)(i);
    }
    return (proj2 - lam2) / (2 * rho);
  }
  // -(dPi C)^T Pi_K*(lambda - rho c), split into its state and control parts
  void AugLagGradient(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> dx, Eigen::Ref<VectorXd> du) {
    Shift(x, u);
    ThroughCone(x, u);
    for (int j = 0; j < n_ + m_; ++j) {
      double g = 0.0;
      for (int i = 0; i < p_; ++i) g -= active_jac_(i, j) * lambda_bar_(i);
      if (j < n_) dx(j) = g; else du(j - n_) = g;
    }
  }
  // Gauss-Newton: rho (dPi C)^T (dPi C); the second-order constraint term does not exist in the reference either
  void AugLagHessian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> dxdx, Eigen::Ref<MatrixXd> dxdu,
                     Eigen::Ref<MatrixXd> dudu, const bool full_newton) {
    if (full_newton) throw std::runtime_error("Second-order constraint terms are not yet supported.");
    const double rho = Shift(x, u);
    ThroughCone(x, u);
    for (int a = 0; a < n_ + m_; ++a)
      for (int b = a < n_ ? 0 : n_; b < n_ + m_; ++b) {
        double h = 0.0;
        for (int i = 0; i < p_; ++i) h += active_jac_(i, a) * active_jac_(i, b);
        h *= rho;
        if (a < n_ && b < n_) dxdx(a, b) = h;
        else if (a < n_) dxdu(a, b - n_) = h;
        else dudu(a - n_, b - n_) = h;
      }
  }

  // ---- outer-loop updates of a standalone object.  A solver's updates are one launch over all knot points
  // (AugmentedLagrangianiLQR::UpdateDuals / UpdatePenalties -> altro_b200_update_duals / _penalties).
  void UpdateDuals() {  // lambda <- Pi_K*(lambda - rho .* c), with the constraint value of the last evaluation
    GetDuals();
    PullPenalty();
    GetConstraintValue();
    for (int i = 0; i < p_; ++i) shifted_(i) = (*lambda_)(i) - rho_(i) * c_(i);
    DualCone::Projection(shifted_, *lambda_);
  }
  void UpdatePenalties() {
    if (core_) throw DeviceError(ALTRO_B200_ERR_UNSUPPORTED,
                                 "the device keeps one penalty per instance: use AugmentedLagrangianiLQR::UpdatePenalties()");
    rho_ *= phi_;
  }
  void ResetDualVariables() {
    GetDuals();
    lambda_->setZero();
  }
  void CalcExpansion(const VectorXdRef& x, const VectorXdRef& u) {
    con_->Evaluate(x, u, c_);
    con_->Jacobian(x, u, jac_);
  }

  template <int norm = Eigen::Infinity>
  double MaxViolation() {
    return GetViolation().template lpNorm<norm>();
  }
  double MaxPenalty() {
    PullPenalty();
    return rho_.maxCoeff();
  }

 private:
  VectorXd DeviceRows(const std::vector<double>& all) const {
    VectorXd out = VectorXd::Zero(p_);
    for (int i = 0; i < p_; ++i) out(i) = all.at(static_cast<size_t>(row0_ + i));
    return out;
  }
  void PullPenalty() {
    if (OnDevice()) rho_.setConstant(core_->MaxPenalty(0));
  }
  // c(x,u), lambda - rho c and its projection onto the dual cone; returns rho
  double Shift(const VectorXdRef& x, const VectorXdRef& u) {
    if (OnDevice()) *lambda_ = DeviceRows(core_->Duals(k_, 0));
    PullPenalty();
    const double rho = rho_(0);
    con_->Evaluate(x, u, c_);
    for (int i = 0; i < p_; ++i) shifted_(i) = (*lambda_)(i) - rho * c_(i);
    DualCone::Projection(shifted_, lambda_bar_);
    return rho;
  }
  // (Jacobian of the dual-cone projection at lambda - rho c) * (constraint Jacobian)
  void ThroughCone(const VectorXdRef& x, const VectorXdRef& u) {
    con_->Jacobian(x, u, jac_);
    cone_jac_.setZero();
    DualCone::Jacobian(shifted_, cone_jac_);
    for (int i = 0; i < p_; ++i)
      for (int j = 0; j < n_ + m_; ++j) {
        double s = 0.0;
        for (int r = 0; r < p_; ++r) s += cone_jac_(i, r) * jac_(r, j);
        active_jac_(i, j) = s;
      }
  }

  const int n_, m_;
  ConstraintPtr<ConType> con_;
  const int p_;
  std::shared_ptr<VectorXd> lambda_;  // shared with the device solver's pending-write list when bound
  VectorXd c_, rho_, viol_, shifted_, lambda_bar_;
  MatrixXd jac_, cone_jac_, active_jac_;
  double phi_ = kDefaultPenaltyScaling;

  std::shared_ptr<altro::detail::DeviceSolver> core_;
  int k_ = 0, row0_ = 0;
};

}  // namespace constraints
}  // namespace altro
