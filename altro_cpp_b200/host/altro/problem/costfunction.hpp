// altro/problem/costfunction.hpp (B200 host mirror) — CostFunction ABC (reference:
// altro/problem/costfunction.hpp:52).
#pragma once

#include "altro/device_descriptor.hpp"
#include "altro/eigentypes.hpp"

namespace altro {
namespace problem {

class CostFunction {
 public:
  virtual ~CostFunction() = default;
  virtual int StateDimension() const = 0;
  virtual int ControlDimension() const = 0;
  virtual bool Describe(device::CostDesc*) const { return false; }
};

}  // namespace problem
}  // namespace altro
