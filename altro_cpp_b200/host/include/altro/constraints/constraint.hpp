// altro/constraints/constraint.hpp (B200 host mirror) — convex cones and the Constraint<ConType>
// ABC (altro/constraints/constraint.hpp:28,65,98,174 there).  The projections below are the
// host-side statement of what csrc/device.cuh applies per row (neg_part, the identity dual cone).
#pragma once

#include <algorithm>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>

#include "altro/common/functionbase.hpp"
#include "altro/eigentypes.hpp"
#include "altro/utils/utils.hpp"

namespace altro {
namespace constraints {

class IdentityCone;

// g(x,u) = 0: projects onto {0}; its dual cone is the whole space
class ZeroCone {
 public:
  ZeroCone() = delete;
  using DualCone = IdentityCone;
  static void Projection(const VectorXdRef& x, Eigen::Ref<VectorXd> x_proj) {
    ALTRO_UNUSED(x);
    x_proj.setZero();
  }
  static void Jacobian(const VectorXdRef& x, Eigen::Ref<MatrixXd> jac) {
    ALTRO_UNUSED(x);
    jac.setZero();
  }
  static void Hessian(const VectorXdRef& x, const VectorXdRef& b, Eigen::Ref<MatrixXd> hess) {
    ALTRO_UNUSED(x);
    ALTRO_UNUSED(b);
    hess.setZero();
  }
};
using Equality = ZeroCone;

class IdentityCone {
 public:
  IdentityCone() = delete;
  using DualCone = ZeroCone;
  static void Projection(const VectorXdRef& x, Eigen::Ref<VectorXd> x_proj) { x_proj = x; }
  static void Jacobian(const VectorXdRef& x, Eigen::Ref<MatrixXd> jac) {
    ALTRO_UNUSED(x);
    jac.setIdentity();
  }
  static void Hessian(const VectorXdRef& x, const VectorXdRef& b, Eigen::Ref<MatrixXd> hess) {
    ALTRO_UNUSED(x);
    ALTRO_UNUSED(b);
    hess.setZero();
  }
};

// h(x,u) <= 0: self-dual, projection min(0, .), Jacobian diag(x_i > 0 ? 0 : 1) — an exact zero counts
// as active
class NegativeOrthant {
 public:
  NegativeOrthant() = delete;
  using DualCone = NegativeOrthant;
  static void Projection(const VectorXdRef& x, Eigen::Ref<VectorXd> x_proj) {
    for (int i = 0; i < x.size(); ++i) x_proj(i) = std::min(0.0, x(i));
  }
  static void Jacobian(const VectorXdRef& x, Eigen::Ref<MatrixXd> jac) {
    for (int i = 0; i < x.size(); ++i) jac(i, i) = x(i) > 0 ? 0 : 1;
  }
  static void Hessian(const VectorXdRef& x, const VectorXdRef& b, Eigen::Ref<MatrixXd> hess) {
    ALTRO_UNUSED(x);
    ALTRO_UNUSED(b);
    hess.setZero();
  }
};
using Inequality = NegativeOrthant;

// one entry of AugmentedLagrangianiLQR::GetConstraintInfo()
struct ConstraintInfo {
  std::string label;
  int index;           // knot point
  VectorXd violation;  // c - Pi_K(c)
  std::string type;
  std::string ToString(int precision = 4) const {
    std::ostringstream os;
    os << type << " at index " << index << ": " << label << " [";
    os << std::setprecision(precision);
    for (int i = 0; i < violation.size(); ++i) os << (i ? ", " : "") << violation(i);
    os << "]";
    return os.str();
  }
};
inline std::ostream& operator<<(std::ostream& os, const ConstraintInfo& info) { return os << info.ToString(); }

template <class ConType>
class Constraint : public FunctionBase {
 public:
  using ConstraintType = ConType;
  int StateDimension() const override {
    ALTRO_ASSERT(false, "StateDimension hasn't been defined for this constraint.");
    return -1;
  }
  int ControlDimension() const override {
    ALTRO_ASSERT(false, "ControlDimension hasn't been defined for this constraint.");
    return -1;
  }
  bool HasHessian() const override { return false; }
  virtual std::string GetLabel() const { return GetConstraintType(); }
  std::string GetConstraintType() const {
    if (std::is_same<ConType, Equality>::value) return "Equality Constraint";
    if (std::is_same<ConType, Inequality>::value) return "Inequality Constraint";
    return "Undefined Constraint Type";
  }
};

template <class ConType>
using ConstraintPtr = std::shared_ptr<Constraint<ConType>>;

}  // namespace constraints
}  // namespace altro
