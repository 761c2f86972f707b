// perf/benchmark_triple_integrator.cpp (B200 host mirror) — the constrained triple-integrator
// benchmark of the reference (perf/benchmark_triple_integrator.cpp:14-56 there; BASELINE config C3
// in its batched form).
//
//   benchmark_triple_integrator [N] [batch]
#include <chrono>
#include <cstdio>
#include <exception>
#include <memory>
#include <string>
#include <vector>

#include "examples/problems/triple_integrator.hpp"

int main(int argc, char* argv[]) {
  const int N = argc > 1 ? std::stoi(argv[1]) : 100;
  const int batch = argc > 2 ? std::stoi(argv[2]) : 1;
  using Clock = std::chrono::high_resolution_clock;
  try {
    altro::problems::TripleIntegratorProblem<2> def;
    def.N = N;
    constexpr int n = altro::problems::TripleIntegratorProblem<2>::NStates;
    constexpr int m = altro::problems::TripleIntegratorProblem<2>::NControls;
    altro::augmented_lagrangian::BatchedAugmentedLagrangianiLQR<n, m> solver(def.MakeProblem(true), batch);
    solver.SetTrajectory(std::make_shared<altro::Trajectory<n, m>>(def.InitialTrajectory()));
    std::vector<altro::VectorXd> x0(batch, def.x0);
    for (int b = 1; b < batch; ++b) x0[b](0) += 1e-3 * (b % 97);
    solver.SetInitialStates(x0);
    for (int run = 0; run < 3; ++run) {
      const auto start = Clock::now();
      solver.Solve();
      const double ms = std::chrono::duration<double, std::milli>(Clock::now() - start).count();
      int solved = 0;
      for (int b = 0; b < batch; ++b) solved += solver.GetStatus(b) == altro::SolverStatus::kSolved;
      std::printf("Run %d: N = %d, batch = %d, solved = %d, nominal iters = %d, viol = %.3e, Time = %.3f ms\n", run, N,
                  batch, solved, solver.GetIterations(0), solver.GetMaxViolation(0), ms);
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "benchmark_triple_integrator: %s\n", e.what());
    return 2;
  }
  return 0;
}
