// examples/unicycle.hpp (B200 host mirror) — the unicycle model of the reference
// (examples/unicycle.cpp:12-33); its f and df/d(x,u) are the device functor `Unicycle` in
// altro_cpp_b200/csrc/device.cuh.
#pragma once

#include "altro/problem/dynamics.hpp"

namespace altro {
namespace examples {

class Unicycle : public problem::ContinuousDynamics {
 public:
  static constexpr int NStates = 3;
  static constexpr int NControls = 2;
  int StateDimension() const override { return NStates; }
  int ControlDimension() const override { return NControls; }
  bool HasHessian() const override { return true; }
  bool Describe(device::ModelDesc* d) const override {
    d->model = ALTRO_B200_MODEL_UNICYCLE;
    d->params.clear();
    return true;
  }
};

}  // namespace examples
}  // namespace altro
