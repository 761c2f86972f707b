// altro/problem/problem.hpp (B200 host mirror) — the container a user fills (reference:
// altro/problem/problem.hpp:65): per-knot shared pointers to dynamics, cost and constraints plus
// the shared initial state.
#pragma once

#include <memory>
#include <stdexcept>
#include <vector>

#include "altro/constraints/constraint.hpp"
#include "altro/problem/costfunction.hpp"
#include "altro/problem/dynamics.hpp"

namespace altro {
namespace problem {

class Problem {
 public:
  explicit Problem(int N)
      : N_(N), initial_state_(std::make_shared<VectorXd>()), costs_(N + 1), models_(N + 1), eq_(N + 1),
        ineq_(N + 1) {}

  void SetInitialState(const VectorXd& x0) { *initial_state_ = x0; }
  void SetCostFunction(std::shared_ptr<CostFunction> cost, int k) { costs_.at(k) = std::move(cost); }
  void SetDynamics(std::shared_ptr<DiscreteDynamics> model, int k) {
    if (k < 0 || k >= N_) throw std::out_of_range("Cannot set dynamics at the terminal knot point.");
    models_.at(k) = std::move(model);
  }
  void SetConstraint(std::shared_ptr<constraints::Constraint<constraints::Equality>> con, int k) {
    eq_.at(k).emplace_back(std::move(con));
  }
  void SetConstraint(std::shared_ptr<constraints::Constraint<constraints::Inequality>> con, int k) {
    ineq_.at(k).emplace_back(std::move(con));
  }

  int NumSegments() const { return N_; }
  int NumConstraints(int k) const {
    int p = 0;
    for (const auto& c : eq_.at(k)) p += c->OutputDimension();
    for (const auto& c : ineq_.at(k)) p += c->OutputDimension();
    return p;
  }
  int NumConstraints() const {
    int p = 0;
    for (int k = 0; k <= N_; ++k) p += NumConstraints(k);
    return p;
  }
  bool IsFullyDefined() const {
    if (initial_state_->size() == 0) return false;
    for (int k = 0; k <= N_; ++k)
      if (!costs_[k] || (k < N_ && !models_[k])) return false;
    return true;
  }
  // Set by augmented_lagrangian::BuildAugLagProblem: an iLQR solver built from this problem
  // minimises the augmented Lagrangian of its constraints (al_problem.hpp:24 there) instead of
  // ignoring them.
  void MarkAugmentedLagrangian(bool v) { auglag_ = v; }
  bool IsAugmentedLagrangian() const { return auglag_; }
  const VectorXd& GetInitialState() const { return *initial_state_; }
  std::shared_ptr<VectorXd> GetInitialStatePointer() const { return initial_state_; }
  std::shared_ptr<CostFunction> GetCostFunction(int k) const { return costs_.at(k); }
  std::shared_ptr<DiscreteDynamics> GetDynamics(int k) const { return models_.at(k); }
  const std::vector<std::vector<constraints::ConstraintPtr<constraints::Equality>>>& GetEqualityConstraints() const {
    return eq_;
  }
  const std::vector<std::vector<constraints::ConstraintPtr<constraints::Inequality>>>& GetInequalityConstraints()
      const {
    return ineq_;
  }

 private:
  int N_;
  bool auglag_ = false;
  std::shared_ptr<VectorXd> initial_state_;
  std::vector<std::shared_ptr<CostFunction>> costs_;
  std::vector<std::shared_ptr<DiscreteDynamics>> models_;
  std::vector<std::vector<constraints::ConstraintPtr<constraints::Equality>>> eq_;
  std::vector<std::vector<constraints::ConstraintPtr<constraints::Inequality>>> ineq_;
};

}  // namespace problem
}  // namespace altro
