// altro/device_descriptor.hpp — how a user-level functor tells the solver what to run on the GPU.
//
// The reference's extension mechanism is subclassing with virtual Evaluate/Jacobian calls; a
// virtual call cannot run on the device.  In this mirror every dynamics / cost /
// constraint class of the closed registry overrides one `Describe(...)` method that writes its
// parameters into a POD; the solver forwards the POD through the C ABI (include/altro_b200.h).
// A functor that does not override it makes the solver constructor throw — there is no host
// fallback.
#pragma once

#include <vector>

#include "altro_b200.h"

namespace altro {
namespace device {

struct ModelDesc {
  int model = -1;  // altro_b200_model
  std::vector<double> params;
};

struct CostDesc {
  std::vector<double> Q, R, H, q, r;  // column-major
  double c = 0.0;
};

struct ConstraintDesc {
  enum Kind { kNone = -1, kGoal = 0, kControlBound = 1, kCircle = 2 } kind = kNone;
  std::vector<double> a, b, c;  // goal: a = xf | bound: a = lb, b = ub | circle: cx, cy, r
  int xi = 0, yi = 1;
};

}  // namespace device
}  // namespace altro
