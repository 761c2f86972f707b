// altro/constraints/constraint.hpp (B200 host mirror) — the convex cones a constraint value is required to
// lie in, and the Constraint<ConType> interface users derive from
// (altro/constraints/constraint.hpp:28,65,98,174 there).
//
// The projections below are the host-side statement of what csrc/device.cuh applies row by row when it
// forms the augmented-Lagrangian terms: neg_part for the negative orthant, the identity for the dual of
// the zero cone.  A user's constraint never runs on the device through these virtuals — it is recognised
// from their answers when a solver is built (altro/device_registry.hpp).
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

namespace detail {
// What the three cones have in common: they are tags (never instantiated) and their projections are
// piecewise linear, so the second-order term of the projection vanishes identically.
struct PiecewiseLinearCone {
  PiecewiseLinearCone() = delete;
  static void Hessian(const VectorXdRef& /*x*/, const VectorXdRef& /*b*/, Eigen::Ref<MatrixXd> hess) { hess.setZero(); }
};
}  // namespace detail

class ZeroCone;
class IdentityCone;
class NegativeOrthant;

// K = {0} (equality constraints g(x,u) = 0).  Everything projects to the origin.
class ZeroCone : public detail::PiecewiseLinearCone {
 public:
  using DualCone = IdentityCone;
  static void Projection(const VectorXdRef& /*x*/, Eigen::Ref<VectorXd> x_proj) { x_proj.setZero(); }
  static void Jacobian(const VectorXdRef& /*x*/, Eigen::Ref<MatrixXd> jac) { jac.setZero(); }
};

// K = R^p, the dual of the zero cone: the projection is the identity map.
class IdentityCone : public detail::PiecewiseLinearCone {
 public:
  using DualCone = ZeroCone;
  static void Projection(const VectorXdRef& x, Eigen::Ref<VectorXd> x_proj) { x_proj = x; }
  static void Jacobian(const VectorXdRef& /*x*/, Eigen::Ref<MatrixXd> jac) { jac.setIdentity(); }
};

// K = {y : y <= 0} (inequality constraints h(x,u) <= 0), self-dual.  Row i projects to min(0, y_i); the
// projection's Jacobian is diagonal with 1 where y_i <= 0 — an exact zero counts as active —
// and only the diagonal is written, like the reference.
class NegativeOrthant : public detail::PiecewiseLinearCone {
 public:
  using DualCone = NegativeOrthant;
  static void Projection(const VectorXdRef& x, Eigen::Ref<VectorXd> x_proj) {
    const int p = static_cast<int>(x.size());
    for (int i = 0; i < p; ++i) x_proj(i) = std::min(0.0, x(i));
  }
  static void Jacobian(const VectorXdRef& x, Eigen::Ref<MatrixXd> jac) {
    const int p = static_cast<int>(x.size());
    for (int i = 0; i < p; ++i) jac(i, i) = (x(i) > 0) ? 0.0 : 1.0;
  }
};

using Equality = ZeroCone;
using Inequality = NegativeOrthant;

// One entry of AugmentedLagrangianiLQR::GetConstraintInfo(): which constraint, at which knot point, how far
// outside its cone (c - Pi_K(c), one number per row).
struct ConstraintInfo {
  std::string label;
  int index = 0;
  VectorXd violation;
  std::string type;

  std::string ToString(int precision = 4) const {
    std::ostringstream text;
    text << type << " at index " << index << ": " << label << " [" << std::setprecision(precision);
    for (int i = 0; i < violation.size(); ++i) text << (i ? ", " : "") << violation(i);
    text << "]";
    return text.str();
  }
};
inline std::ostream& operator<<(std::ostream& os, const ConstraintInfo& info) { return os << info.ToString(); }

// The interface a constraint implements: FunctionBase's Evaluate / Jacobian with OutputDimension() rows,
// tagged with the cone its value must lie in.
template <class ConType>
class Constraint : public FunctionBase {
  static constexpr bool kIsEquality = std::is_same<ConType, Equality>::value;
  static constexpr bool kIsInequality = std::is_same<ConType, Inequality>::value;

 public:
  using ConstraintType = ConType;

  std::string GetConstraintType() const {
    return kIsEquality ? "Equality Constraint" : kIsInequality ? "Inequality Constraint" : "Undefined Constraint Type";
  }
  virtual std::string GetLabel() const { return GetConstraintType(); }
  bool HasHessian() const override { return false; }

  // a constraint that can be sized statically says so by overriding these; asking one that does not is a
  // contract violation, as in the reference
  int StateDimension() const override { return Undefined("StateDimension hasn't been defined for this constraint."); }
  int ControlDimension() const override { return Undefined("ControlDimension hasn't been defined for this constraint."); }

 private:
  static int Undefined(const char* what) {
    ALTRO_ASSERT(false, what);
    ALTRO_UNUSED(what);
    return -1;
  }
};

template <class ConType>
using ConstraintPtr = std::shared_ptr<Constraint<ConType>>;

}  // namespace constraints
}  // namespace altro
