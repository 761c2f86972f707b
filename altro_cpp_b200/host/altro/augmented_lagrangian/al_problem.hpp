// altro/augmented_lagrangian/al_problem.hpp (B200 host mirror) — BuildAugLagProblem (reference:
// altro/augmented_lagrangian/al_problem.hpp:24-63).  On the device the ALCost of a knot point is
// not a separate object: the constraint rows travel with the problem and the kernels add the
// augmented-Lagrangian terms when the solver is created with use_constraints = 1.  The returned
// problem therefore only carries that request.
#pragma once

#include "altro/problem/problem.hpp"

namespace altro {
namespace augmented_lagrangian {

template <int n, int m>
problem::Problem BuildAugLagProblem(const problem::Problem& prob) {
  problem::Problem out = prob;
  out.MarkAugmentedLagrangian(true);
  return out;
}

}  // namespace augmented_lagrangian
}  // namespace altro
