// altro/augmented_lagrangian/al_problem.hpp (B200 host mirror) — BuildAugLagProblem
// (altro/augmented_lagrangian/al_problem.hpp:24-63 there): an unconstrained problem whose cost functions are the
// ALCost objects of the constrained one.  The returned problem evaluates like the reference's on the host
// (GetCostFunction(k)->Evaluate adds the augmented-Lagrangian terms, multipliers zero, penalties one) and carries a
// pointer to `prob`: an iLQR solver built from it sends `prob`'s cost and constraint rows to the device and the
// kernels form the same terms there (use_constraints = 1).
#pragma once

#include <memory>
#include <vector>

#include "altro/augmented_lagrangian/al_cost.hpp"
#include "altro/problem/problem.hpp"

namespace altro {
namespace augmented_lagrangian {

template <int n, int m>
problem::Problem BuildAugLagProblem(const problem::Problem& prob,
                                    std::vector<std::shared_ptr<ALCost<n, m>>>* costs = nullptr) {
  ALTRO_ASSERT(prob.IsFullyDefined(), "Expected problem to be fully defined.");
  const int N = prob.NumSegments();
  problem::Problem out(N, prob.GetInitialStatePointer());
  for (int k = 0; k < N; ++k) out.SetDynamics(prob.GetDynamics(k), k);
  for (int k = 0; k <= N; ++k) {
    std::shared_ptr<ALCost<n, m>> alcost = std::make_shared<ALCost<n, m>>(prob, k);
    if (costs != nullptr) costs->emplace_back(alcost);
    out.SetCostFunction(alcost, k);
  }
  out.MarkAugmentedLagrangian(std::make_shared<const problem::Problem>(prob));
  return out;
}

}  // namespace augmented_lagrangian
}  // namespace altro
