// altro/problem/dynamics.hpp (B200 host mirror) — the plugin ABCs of the reference
// (altro/problem/dynamics.hpp:59,148).  Evaluation happens on the device; a model takes part by
// describing itself (altro/device_descriptor.hpp).
#pragma once

#include "altro/device_descriptor.hpp"
#include "altro/eigentypes.hpp"

namespace altro {
namespace problem {

class ContinuousDynamics {
 public:
  virtual ~ContinuousDynamics() = default;
  virtual int StateDimension() const = 0;
  virtual int ControlDimension() const = 0;
  virtual bool HasHessian() const { return false; }
  virtual bool Describe(device::ModelDesc*) const { return false; }
};

class DiscreteDynamics {
 public:
  virtual ~DiscreteDynamics() = default;
  virtual int StateDimension() const = 0;
  virtual int ControlDimension() const = 0;
  virtual bool HasHessian() const { return false; }
  virtual bool Describe(device::ModelDesc*) const { return false; }
};

}  // namespace problem
}  // namespace altro
