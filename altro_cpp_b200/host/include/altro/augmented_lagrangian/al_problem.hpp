// altro/augmented_lagrangian/al_problem.hpp (B200 host mirror) — BuildAugLagProblem
// (altro/augmented_lagrangian/al_problem.hpp:24-63 there).  On the device the ALCost of a knot point
// is not a separate functor: the constraint rows travel with the problem and the kernels add the
// augmented-Lagrangian terms when the solver is created with use_constraints = 1.  The returned
// problem therefore shares everything with `prob` and only carries that request.
#pragma once

#include <memory>
#include <vector>

#include "altro/problem/problem.hpp"

namespace altro {
namespace augmented_lagrangian {

template <int n, int m>
class ALCost;

template <int n, int m>
problem::Problem BuildAugLagProblem(const problem::Problem& prob,
                                    std::vector<std::shared_ptr<ALCost<n, m>>>* costs = nullptr) {
  ALTRO_UNUSED(costs);  // per-knot views are created by the solver once the device state exists (GetALCost)
  problem::Problem out = prob;
  out.MarkAugmentedLagrangian(true);
  return out;
}

}  // namespace augmented_lagrangian
}  // namespace altro
