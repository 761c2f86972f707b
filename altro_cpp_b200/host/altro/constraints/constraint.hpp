// altro/constraints/constraint.hpp (B200 host mirror) — cone tags and the Constraint<ConType>
// ABC (reference: altro/constraints/constraint.hpp:28,98,174).  The cone projections and their
// Jacobians are evaluated in device code (altro_cpp_b200/csrc/device.cuh al_value, al_expansion).
#pragma once

#include <memory>
#include <sstream>
#include <string>
#include <type_traits>

#include "altro/device_descriptor.hpp"
#include "altro/eigentypes.hpp"

namespace altro {
namespace constraints {

class ZeroCone {};
class NegativeOrthant {};
using Equality = ZeroCone;
using Inequality = NegativeOrthant;

// One entry of AugmentedLagrangianiLQR::GetConstraintInfo() (constraint.hpp:134-141 there)
struct ConstraintInfo {
  std::string label;
  int index;           // knot point
  VectorXd violation;  // c - Pi_K(c)
  std::string type;
  std::string ToString(int precision = 4) const {
    std::ostringstream os;
    os << label << " at index " << index << ": [";
    os.precision(precision);
    for (int i = 0; i < violation.size(); ++i) os << (i ? ", " : "") << violation(i);
    os << "]";
    return os.str();
  }
};

template <class ConType>
class Constraint {
 public:
  using ConstraintType = ConType;
  virtual ~Constraint() = default;
  virtual int OutputDimension() const = 0;
  virtual std::string GetLabel() const { return GetConstraintType(); }
  std::string GetConstraintType() const {
    return std::is_same<ConType, Equality>::value ? "Equality Constraint" : "Inequality Constraint";
  }
  virtual bool Describe(device::ConstraintDesc*) const { return false; }
};

template <class ConType>
using ConstraintPtr = std::shared_ptr<Constraint<ConType>>;

}  // namespace constraints
}  // namespace altro
