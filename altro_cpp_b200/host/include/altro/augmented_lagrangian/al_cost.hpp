// altro/augmented_lagrangian/al_cost.hpp (B200 host mirror) — ALCost<n,m>: the cost of one knot point plus the
// augmented-Lagrangian terms of its constraints (altro/augmented_lagrangian/al_cost.hpp:37 there), and the object
// behind AugmentedLagrangianiLQR::GetALCost(k).
//
// The multipliers and penalties belong to its ConstraintValues (altro/constraints/constraint_values.hpp).  Built
// from a problem (or empty, then filled with SetCostFunction / Set...Constraints) it is a self-contained host
// object: Evaluate / Gradient / Hessian at a point (x, u) the caller passes.  Built by a solver it is bound to
// that solver's device state: the same calls use the device's multipliers and penalty, MaxViolation() and the
// constraint getters report the solver's current trajectory.  The solve itself never comes through here — the
// kernels add these terms (csrc/device.cuh al_value / al_expansion).
#pragma once

#include <algorithm>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "altro/constraints/constraint.hpp"
#include "altro/constraints/constraint_values.hpp"
#include "altro/device_solver.hpp"
#include "altro/problem/costfunction.hpp"
#include "altro/problem/problem.hpp"

namespace altro {
namespace augmented_lagrangian {

template <int n, int m>
class ALCost : public problem::CostFunction {
 public:
  template <class ConType>
  using ConstraintValueVec = std::vector<std::shared_ptr<constraints::ConstraintValues<n, m, ConType>>>;

  // no cost, no constraints yet
  ALCost(int state_dim, int control_dim) : n_(state_dim), m_(control_dim) { Scratch(); }
  // cost and constraints of knot k of `prob`; multipliers zero, penalties one
  ALCost(const problem::Problem& prob, int k) : n_(prob.GetDynamics(k)->StateDimension()),
                                                m_(prob.GetDynamics(k)->ControlDimension()), k_(k) {
    Scratch();
    SetCostFunction(prob.GetCostFunction(k));
    SetEqualityConstraints(prob.GetEqualityConstraints().at(k).begin(), prob.GetEqualityConstraints().at(k).end());
    SetInequalityConstraints(prob.GetInequalityConstraints().at(k).begin(), prob.GetInequalityConstraints().at(k).end());
  }
  // knot k of a device solver: multipliers, penalty and constraint values are the device's
  ALCost(std::shared_ptr<detail::DeviceSolver> core, int k) : ALCost(core->GetProblem(), k) {
    core_ = std::move(core);
    int row = 0;  // ALCost row order: equalities, then inequalities
    for (auto& v : eq_) {
      v->BindDevice(core_, k, row);
      row += v->OutputDimension();
    }
    for (auto& v : ineq_) {
      v->BindDevice(core_, k, row);
      row += v->OutputDimension();
    }
  }

  // ---- contents
  void SetCostFunction(const std::shared_ptr<problem::CostFunction>& costfun) {
    ALTRO_ASSERT(costfun != nullptr, "Cost function cannot be a nullptr.");
    costfun_ = costfun;
  }
  template <class Iterator>
  void SetEqualityConstraints(const Iterator& begin, const Iterator& end) {
    Wrap(begin, end, &eq_);
  }
  template <class Iterator>
  void SetInequalityConstraints(const Iterator& begin, const Iterator& end) {
    Wrap(begin, end, &ineq_);
  }
  std::shared_ptr<problem::CostFunction> GetCostFunction() { return costfun_; }
  ConstraintValueVec<constraints::Equality>& GetEqualityConstraints() { return eq_; }
  ConstraintValueVec<constraints::Inequality>& GetInequalityConstraints() { return ineq_; }
  int NumConstraints() {
    int rows = 0;
    for (const auto& v : eq_) rows += v->OutputDimension();
    for (const auto& v : ineq_) rows += v->OutputDimension();
    return rows;
  }
  template <class ConType>
  int NumConstraintFunctions() {
    return static_cast<int>(Of<ConType>().size());
  }
  void GetConstraintInfo(std::vector<constraints::ConstraintInfo>* coninfo) {
    for (auto& v : eq_) coninfo->emplace_back(v->GetConstraintInfo());
    for (auto& v : ineq_) coninfo->emplace_back(v->GetConstraintInfo());
  }

  // ---- penalties: one constraint function, or all of one cone
  template <class ConType>
  void SetPenalty(const double rho, const int i) {
    ALTRO_ASSERT(0 <= i && i < NumConstraintFunctions<ConType>(), "Invalid constraint index.");
    Of<ConType>().at(i)->SetPenalty(rho);
  }
  template <class ConType>
  void SetPenalty(const double rho) {
    for (auto& v : Of<ConType>()) v->SetPenalty(rho);
  }
  template <class ConType>
  void SetPenaltyScaling(const double phi, const int i) {
    ALTRO_ASSERT(0 <= i && i < NumConstraintFunctions<ConType>(), "Invalid constraint index.");
    Of<ConType>().at(i)->SetPenaltyScaling(phi);
  }
  template <class ConType>
  void SetPenaltyScaling(const double phi) {
    for (auto& v : Of<ConType>()) v->SetPenaltyScaling(phi);
  }

  // ---- CostFunction interface: cost + sum of the constraints' augmented-Lagrangian terms at (x, u)
  using problem::CostFunction::Gradient;
  using problem::CostFunction::Hessian;
  int StateDimension() const override { return n_; }
  int ControlDimension() const override { return m_; }
  double Evaluate(const VectorXdRef& x, const VectorXdRef& u) override {
    ALTRO_ASSERT(costfun_ != nullptr, "Cost function must be set before evaluating.");
    double J = costfun_->Evaluate(x, u);
    for (auto& v : eq_) J += v->AugLag(x, u);
    for (auto& v : ineq_) J += v->AugLag(x, u);
    return J;
  }
  void Gradient(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> dx, Eigen::Ref<VectorXd> du) override {
    ALTRO_ASSERT(costfun_ != nullptr, "Cost function must be set before evaluating.");
    costfun_->Gradient(x, u, dx, du);
    auto add = [&](auto& v) {
      v->AugLagGradient(x, u, gx_, gu_);
      for (int i = 0; i < n_; ++i) dx(i) += gx_(i);
      for (int i = 0; i < m_; ++i) du(i) += gu_(i);
    };
    for (auto& v : eq_) add(v);
    for (auto& v : ineq_) add(v);
  }
  void Hessian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> dxdx, Eigen::Ref<MatrixXd> dxdu,
               Eigen::Ref<MatrixXd> dudu) override {
    ALTRO_ASSERT(costfun_ != nullptr, "Cost function must be set before evaluating.");
    costfun_->Hessian(x, u, dxdx, dxdu, dudu);
    auto add = [&](auto& v) {
      v->AugLagHessian(x, u, hxx_, hxu_, huu_, /*full_newton=*/false);
      for (int i = 0; i < n_; ++i) {
        for (int j = 0; j < n_; ++j) dxdx(i, j) += hxx_(i, j);
        for (int j = 0; j < m_; ++j) dxdu(i, j) += hxu_(i, j);
      }
      for (int i = 0; i < m_; ++i)
        for (int j = 0; j < m_; ++j) dudu(i, j) += huu_(i, j);
    };
    for (auto& v : eq_) add(v);
    for (auto& v : ineq_) add(v);
  }

  // ---- outer-loop steps of a standalone object (a solver updates all knot points in one launch)
  void UpdateDuals() {
    for (auto& v : eq_) v->UpdateDuals();
    for (auto& v : ineq_) v->UpdateDuals();
  }
  void UpdatePenalties() {
    for (auto& v : eq_) v->UpdatePenalties();
    for (auto& v : ineq_) v->UpdatePenalties();
  }
  void ResetDualVariables() {
    for (auto& v : eq_) v->ResetDualVariables();
    for (auto& v : ineq_) v->ResetDualVariables();
  }
  template <int norm = Eigen::Infinity>
  double MaxViolation() {
    double worst = 0.0;  // the norm over constraints is taken as a maximum for every `norm`, as there
    for (auto& v : eq_) worst = std::max(worst, v->template MaxViolation<norm>());
    for (auto& v : ineq_) worst = std::max(worst, v->template MaxViolation<norm>());
    return worst;
  }
  double MaxPenalty() {
    double worst = 0.0;
    for (auto& v : eq_) worst = std::max(worst, v->MaxPenalty());
    for (auto& v : ineq_) worst = std::max(worst, v->MaxPenalty());
    return worst;
  }

 private:
  template <class Iterator, class ConType>
  void Wrap(const Iterator& begin, const Iterator& end, ConstraintValueVec<ConType>* values) {
    ALTRO_ASSERT(values != nullptr, "Must provide a pointer to a valid collection.");
    values->clear();
    for (Iterator it = begin; it != end; ++it)
      values->emplace_back(std::make_shared<constraints::ConstraintValues<n, m, ConType>>(n_, m_, *it));
  }
  // the collection of one cone, selected by type
  ConstraintValueVec<constraints::Equality>& Pick(const constraints::Equality*) { return eq_; }
  ConstraintValueVec<constraints::Inequality>& Pick(const constraints::Inequality*) { return ineq_; }
  template <class ConType>
  ConstraintValueVec<ConType>& Of() {
    return Pick(static_cast<const ConType*>(nullptr));
  }
  void Scratch() {
    gx_ = VectorXd::Zero(n_);
    gu_ = VectorXd::Zero(m_);
    hxx_ = MatrixXd::Zero(n_, n_);
    hxu_ = MatrixXd::Zero(n_, m_);
    huu_ = MatrixXd::Zero(m_, m_);
  }

  int n_, m_;
  int k_ = 0;
  std::shared_ptr<problem::CostFunction> costfun_;
  ConstraintValueVec<constraints::Equality> eq_;
  ConstraintValueVec<constraints::Inequality> ineq_;
  VectorXd gx_, gu_;       // one constraint's contribution, before it is added
  MatrixXd hxx_, hxu_, huu_;
  std::shared_ptr<detail::DeviceSolver> core_;
};

}  // namespace augmented_lagrangian
}  // namespace altro
