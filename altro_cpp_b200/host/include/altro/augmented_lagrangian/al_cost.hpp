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
};

template <int n, int m>
class ALCost {
 public:
  ALCost(std::shared_ptr<detail::DeviceSolver> core, int k) : core_(std::move(core)), k_(k) {
    const problem::Problem& prob = core_->GetProblem();
    int row = 0;  // ALCost order: equalities, then inequalities (al_cost.hpp:264-273 there)
    for (const auto& con : prob.GetEqualityConstraints()[k]) {
      eq_.emplace_back(std::make_shared<ConstraintValuesView>(core_, k, row, con->OutputDimension(), con->GetLabel(),
                                                              con->GetConstraintType(), true));
      row += con->OutputDimension();
    }
    for (const auto& con : prob.GetInequalityConstraints()[k]) {
      ineq_.emplace_back(std::make_shared<ConstraintValuesView>(core_, k, row, con->OutputDimension(), con->GetLabel(),
                                                                con->GetConstraintType(), false));
      row += con->OutputDimension();
    }
    p_ = row;
  }
  int NumConstraints() const { return p_; }
  const std::vector<std::shared_ptr<ConstraintValuesView>>& GetEqualityConstraints() const { return eq_; }
  const std::vector<std::shared_ptr<ConstraintValuesView>>& GetInequalityConstraints() const { return ineq_; }
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

 private:
  std::shared_ptr<detail::DeviceSolver> core_;
  int k_;
  int p_ = 0;
  std::vector<std::shared_ptr<ConstraintValuesView>> eq_, ineq_;
};

}  // namespace augmented_lagrangian
}  // namespace altro
