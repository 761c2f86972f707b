// altro/augmented_lagrangian/al_cost.hpp (B200 host mirror) — ALCost<n,m>, the per-knot object
// behind AugmentedLagrangianiLQR::GetALCost(k) (altro/augmented_lagrangian/al_cost.hpp:37 there).
//
// In the reference this object evaluates cost + augmented-Lagrangian terms and owns the duals and
// penalties of the knot's constraints (ConstraintValues, altro/constraints/constraint_values.hpp:24).
// Here those live on the device (csrc/device.cuh al_value / al_expansion; duals in LAM, one penalty
// per instance) and this is the host-side view of them: constraint counts, duals, penalty, constraint
// values and violations of the current trajectory, fetched from the device on request.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "altro/constraints/constraint.hpp"
#include "altro/device_solver.hpp"
#include "altro/problem/costfunction.hpp"

namespace altro {
namespace augmented_lagrangian {

// one constraint of one knot (the reference's ConstraintValues<n, m, ConType>)
class ConstraintValuesView {
 public:
  ConstraintValuesView(std::shared_ptr<detail::DeviceSolver> core, int k, int row0, int p, std::string label, std::string type,
                       bool equality)
      : core_(std::move(core)), k_(k), row0_(row0), p_(p), label_(std::move(label)), type_(std::move(type)),
        equality_(equality) {}
  int OutputDimension() const { return p_; }
  std::string GetLabel() const { return label_; }
  VectorXd GetDuals() const { return Slice(core_->Duals(k_, 0)); }
  // writable: the returned vector starts as the device's duals and is sent back right before the next device
  // phase (every instance of a batched solver receives the same values)
  VectorXd& GetDuals() {
    if (!edited_) edited_ = std::make_shared<VectorXd>();
    *edited_ = Slice(core_->Duals(k_, 0));
    core_->PushDualsBeforeNextRun(k_, row0_, edited_);
    return *edited_;
  }
  // the device keeps ONE penalty per instance, uniform over all constraints: this sets it
  void SetPenalty(double rho) { core_->SetPenalty(rho); }
  VectorXd GetConstraintValue() const { return Slice(core_->ConstraintValues(k_, 0)); }
  // the penalty is one scalar per instance on the device, uniform over the rows
  VectorXd GetPenalty() const { return VectorXd::Constant(p_, core_->MaxPenalty(0)); }
  double MaxPenalty() const { return core_->MaxPenalty(0); }
  // c - Pi_K(c): |c| for equalities, max(0, c) for inequalities (constraint_values.hpp:216-221 there)
  VectorXd GetViolation() const {
    VectorXd v = GetConstraintValue();
    for (int i = 0; i < p_; ++i) v(i) = equality_ ? v(i) : (v(i) > 0.0 ? v(i) : 0.0);
    return v;
  }
  double MaxViolation() const {
    const VectorXd v = GetViolation();
    double r = 0.0;
    for (int i = 0; i < p_; ++i) r = std::max(r, std::fabs(v(i)));
    return r;
  }
  constraints::ConstraintInfo GetConstraintInfo() const { return constraints::ConstraintInfo{label_, k_, GetViolation(), type_}; }

 private:
  VectorXd Slice(const std::vector<double>& all) const {
    VectorXd out = VectorXd::Zero(p_);
    for (int i = 0; i < p_; ++i) out(i) = all.at(static_cast<size_t>(row0_ + i));
    return out;
  }
  std::shared_ptr<detail::DeviceSolver> core_;
  int k_, row0_, p_;
  std::string label_, type_;
  bool equality_;
  std::shared_ptr<VectorXd> edited_;
};

}  // namespace augmented_lagrangian

namespace constraints {
// the reference's name and template signature for the per-constraint state of an ALCost
// (altro/constraints/constraint_values.hpp:24 there); here a typed handle on the device-side state
template <int n, int m, class ConType>
class ConstraintValues : public augmented_lagrangian::ConstraintValuesView {
 public:
  using augmented_lagrangian::ConstraintValuesView::ConstraintValuesView;
};
}  // namespace constraints

namespace augmented_lagrangian {

// As a CostFunction this object evaluates the PLAIN cost of its knot (it forwards to the user's functor): the
// augmented-Lagrangian terms it stands for are added on the device (csrc/device.cuh al_value / al_expansion),
// and iLQR::Cost() / GetCostExpansion() of an AL problem report them from there.
template <int n, int m>
class ALCost : public problem::CostFunction {
  using EqValues = constraints::ConstraintValues<n, m, constraints::Equality>;
  using IneqValues = constraints::ConstraintValues<n, m, constraints::Inequality>;

 public:
  ALCost(std::shared_ptr<detail::DeviceSolver> core, int k) : core_(std::move(core)), k_(k) {
    const problem::Problem& prob = core_->GetProblem();
    base_ = prob.GetCostFunction(k);
    int row = 0;  // ALCost order: equalities, then inequalities (al_cost.hpp:264-273 there)
    for (const auto& con : prob.GetEqualityConstraints()[k]) {
      eq_.emplace_back(std::make_shared<EqValues>(core_, k, row, con->OutputDimension(), con->GetLabel(),
                                                  con->GetConstraintType(), true));
      row += con->OutputDimension();
    }
    for (const auto& con : prob.GetInequalityConstraints()[k]) {
      ineq_.emplace_back(std::make_shared<IneqValues>(core_, k, row, con->OutputDimension(), con->GetLabel(),
                                                      con->GetConstraintType(), false));
      row += con->OutputDimension();
    }
    p_ = row;
  }
  int NumConstraints() const { return p_; }
  const std::vector<std::shared_ptr<EqValues>>& GetEqualityConstraints() const { return eq_; }
  const std::vector<std::shared_ptr<IneqValues>>& GetInequalityConstraints() const { return ineq_; }
  std::shared_ptr<problem::CostFunction> GetCostFunction() const { return base_; }
  double MaxViolation() const {
    double r = 0.0;
    for (const auto& c : eq_) r = std::max(r, c->MaxViolation());
    for (const auto& c : ineq_) r = std::max(r, c->MaxViolation());
    return r;
  }
  double MaxPenalty() const { return p_ > 0 ? core_->MaxPenalty(0) : 0.0; }
  void GetConstraintInfo(std::vector<constraints::ConstraintInfo>* coninfo) const {
    for (const auto& c : eq_) coninfo->emplace_back(c->GetConstraintInfo());
    for (const auto& c : ineq_) coninfo->emplace_back(c->GetConstraintInfo());
  }

  // ---- CostFunction interface: the plain cost of this knot
  using problem::CostFunction::Gradient;
  using problem::CostFunction::Hessian;
  int StateDimension() const override { return base_->StateDimension(); }
  int ControlDimension() const override { return base_->ControlDimension(); }
  double Evaluate(const VectorXdRef& x, const VectorXdRef& u) override { return base_->Evaluate(x, u); }
  void Gradient(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> dx, Eigen::Ref<VectorXd> du) override {
    base_->Gradient(x, u, dx, du);
  }
  void Hessian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> dxdx, Eigen::Ref<MatrixXd> dxdu,
               Eigen::Ref<MatrixXd> dudu) override {
    base_->Hessian(x, u, dxdx, dxdu, dudu);
  }

 private:
  std::shared_ptr<detail::DeviceSolver> core_;
  std::shared_ptr<problem::CostFunction> base_;
  int k_;
  int p_ = 0;
  std::vector<std::shared_ptr<EqValues>> eq_;
  std::vector<std::shared_ptr<IneqValues>> ineq_;
};

}  // namespace augmented_lagrangian
}  // namespace altro
