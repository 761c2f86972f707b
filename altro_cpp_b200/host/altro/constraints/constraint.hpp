// altro/constraints/constraint.hpp (B200 host mirror) — cone tags and the Constraint<ConType>
// ABC (reference: altro/constraints/constraint.hpp:28,98,174).  The cone projections and their
// Jacobians are evaluated in device code (altro_cpp_b200/csrc/device.cuh al_value, al_expansion).
#pragma once

#include <memory>
#include <string>

#include "altro/device_descriptor.hpp"
#include "altro/eigentypes.hpp"

namespace altro {
namespace constraints {

class ZeroCone {};
class NegativeOrthant {};
using Equality = ZeroCone;
using Inequality = NegativeOrthant;

template <class ConType>
class Constraint {
 public:
  using ConstraintType = ConType;
  virtual ~Constraint() = default;
  virtual int OutputDimension() const = 0;
  virtual std::string GetLabel() const { return "Constraint"; }
  virtual bool Describe(device::ConstraintDesc*) const { return false; }
};

template <class ConType>
using ConstraintPtr = std::shared_ptr<Constraint<ConType>>;

}  // namespace constraints
}  // namespace altro
