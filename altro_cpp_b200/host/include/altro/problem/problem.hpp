// altro/problem/problem.hpp (B200 host mirror) — the container a user fills
// (altro/problem/problem.hpp:65 there): per-knot shared pointers to dynamics, cost and constraints
// and the shared initial state.  Copies share the functors and the initial-state vector.
#pragma once

#include <memory>
#include <vector>

#include "altro/constraints/constraint.hpp"
#include "altro/eigentypes.hpp"
#include "altro/problem/costfunction.hpp"
#include "altro/problem/dynamics.hpp"

namespace altro {
namespace problem {

// x+ = x: installed at the terminal knot by SetDynamics (problem.hpp:30,161-164 there)
class IdentityDynamics : public DiscreteDynamics {
 public:
  using DiscreteDynamics::Evaluate;
  IdentityDynamics(int n, int m) : n_(n), m_(m) {
    ALTRO_ASSERT(n > 0, "State dimension must be greater than zero.");
    ALTRO_ASSERT(m > 0, "Control dimension must be greater than zero.");
  }
  int StateDimension() const override { return n_; }
  int ControlDimension() const override { return m_; }
  void Evaluate(const VectorXdRef& x, const VectorXdRef&, float, float, Eigen::Ref<VectorXd> xnext) override { xnext = x; }
  void Jacobian(const VectorXdRef&, const VectorXdRef&, float, float, Eigen::Ref<MatrixXd> jac) override { jac.setIdentity(); }
  void Hessian(const VectorXdRef&, const VectorXdRef&, float, float, const VectorXdRef&, Eigen::Ref<MatrixXd> hess) override {
    hess.setZero();
  }
  bool HasHessian() const override { return true; }

 private:
  int n_, m_;
};

class Problem {
  template <class ConType>
  using ConstraintSet = std::vector<constraints::ConstraintPtr<ConType>>;

 public:
  explicit Problem(int N, std::shared_ptr<VectorXd> initial_state = std::make_shared<VectorXd>(0))
      : N_(N), initial_state_(std::move(initial_state)), costfuns_(N + 1), models_(N + 1), eq_(N + 1), ineq_(N + 1) {}

  void SetInitialState(const VectorXdRef& x0) { *initial_state_ = x0; }

  void SetCostFunction(std::shared_ptr<CostFunction> costfun, int k) {
    ALTRO_ASSERT(k >= 0 && k <= N_, "Invalid knot point index.");
    costfuns_.at(k) = std::move(costfun);
  }
  template <class CostFun>
  void SetCostFunction(const std::vector<std::shared_ptr<CostFun>>& costfuns, int k_start = 0) {
    for (size_t i = 0; i < costfuns.size(); ++i) SetCostFunction(costfuns[i], static_cast<int>(i) + k_start);
  }
  void SetDynamics(std::shared_ptr<DiscreteDynamics> model, int k) {
    ALTRO_ASSERT(model != nullptr, "Cannot pass a nullptr for the dynamics.");
    ALTRO_ASSERT(k >= 0 && k < N_, "Invalid knot point index.");
    if (k == N_ - 1) models_.at(N_) = std::make_shared<IdentityDynamics>(model->StateDimension(), model->ControlDimension());
    models_.at(k) = std::move(model);
  }
  template <class Dynamics>
  void SetDynamics(const std::vector<std::shared_ptr<Dynamics>>& models, int k_start = 0) {
    for (size_t i = 0; i < models.size(); ++i) SetDynamics(models[i], static_cast<int>(i) + k_start);
  }
  template <class ConstraintObject>
  void SetConstraint(std::shared_ptr<ConstraintObject> con, int k) {
    using ConType = typename ConstraintObject::ConstraintType;
    constraints::ConstraintPtr<ConType> ptr = con;
    ALTRO_ASSERT(ptr != nullptr, "Must provide a valid constraint pointer.");
    ALTRO_ASSERT(ptr->OutputDimension() > 0, "Constraint must have a length greater than zero.");
    AddConstraint(std::move(ptr), k);
  }

  int NumConstraints(int k) const {
    ALTRO_ASSERT(0 <= k && k <= N_, "k outside valid knot point indices.");
    int cnt = 0;
    for (const auto& con : eq_.at(k)) cnt += con->OutputDimension();
    for (const auto& con : ineq_.at(k)) cnt += con->OutputDimension();
    return cnt;
  }
  int NumConstraints() const {
    int cnt = 0;
    for (int k = 0; k <= N_; ++k) cnt += NumConstraints(k);
    return cnt;
  }
  const VectorXd& GetInitialState() const { return *initial_state_; }
  std::shared_ptr<VectorXd> GetInitialStatePointer() const { return initial_state_; }
  std::shared_ptr<CostFunction> GetCostFunction(int k) const { return costfuns_.at(k); }
  std::shared_ptr<DiscreteDynamics> GetDynamics(int k) const {
    ALTRO_ASSERT(models_.at(k) != nullptr, "Dynamics have not been defined at this knot point.");
    return models_.at(k);
  }
  const std::vector<ConstraintSet<constraints::Equality>>& GetEqualityConstraints() const { return eq_; }
  const std::vector<ConstraintSet<constraints::Inequality>>& GetInequalityConstraints() const { return ineq_; }
  int NumSegments() const { return N_; }

  bool IsFullyDefined(bool verbose = false) const {
    bool ok = true;
    if (initial_state_->size() == 0) {
      if (verbose) std::cerr << "Initial state is not set." << std::endl;
      ok = false;
    } else if (models_[0] && models_[0]->StateDimension() != initial_state_->size()) {
      if (verbose) std::cerr << "The initial state has " << initial_state_->size() << " entries, the first model has "
                             << models_[0]->StateDimension() << " states." << std::endl;
      ok = false;
    }
    for (int k = 0; k <= N_; ++k) {
      const bool has_model = k == N_ || models_[k] != nullptr;
      if (!costfuns_[k] || !has_model) {
        if (verbose) std::cerr << "Knot point " << k << " is missing its " << (costfuns_[k] ? "dynamics" : "cost function") << std::endl;
        ok = false;
      }
    }
    return ok;
  }

  // Set by augmented_lagrangian::BuildAugLagProblem on the problem it returns (costs wrapped in ALCost objects, no
  // constraints of its own): `constrained` is the problem it was built from.  A solver built from the marked
  // problem describes `constrained` to the device and asks the kernels for the augmented-Lagrangian terms.
  void MarkAugmentedLagrangian(std::shared_ptr<const Problem> constrained) { auglag_source_ = std::move(constrained); }
  bool IsAugmentedLagrangian() const { return auglag_source_ != nullptr; }
  const Problem& ConstrainedProblem() const { return auglag_source_ ? *auglag_source_ : *this; }

 private:
  void AddConstraint(constraints::ConstraintPtr<constraints::Equality> con, int k) { eq_.at(k).emplace_back(std::move(con)); }
  void AddConstraint(constraints::ConstraintPtr<constraints::Inequality> con, int k) { ineq_.at(k).emplace_back(std::move(con)); }

  int N_;
  std::shared_ptr<const Problem> auglag_source_;
  std::shared_ptr<VectorXd> initial_state_;
  std::vector<std::shared_ptr<CostFunction>> costfuns_;
  std::vector<std::shared_ptr<DiscreteDynamics>> models_;
  std::vector<ConstraintSet<constraints::Equality>> eq_;
  std::vector<ConstraintSet<constraints::Inequality>> ineq_;
};

}  // namespace problem
}  // namespace altro
