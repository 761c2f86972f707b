// perf/benchmark_unicycle.cpp (B200 host mirror) — the reference's headline benchmark program
// (perf/benchmark_unicycle.cpp:18-96 there) written against this repo's host API, plus the
// batched form the device exists for.
//
//   benchmark_unicycle [nruns] [batch]
//
// nruns  : how many times the three-obstacle unicycle problem is solved (default 1)
// batch  : 1 = the reference's single-instance AugmentedLagrangianiLQR<3,2>;
//          >1 = BatchedAugmentedLagrangianiLQR<3,2> over `batch` perturbed initial states.
// Exit code 2 = the device library reported an error (e.g. no usable CUDA device).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <exception>
#include <memory>
#include <string>
#include <vector>

#include "altro/augmented_lagrangian/al_solver.hpp"
#include "altro/common/solver_options.hpp"
#include "altro/ilqr/ilqr.hpp"
#include "examples/problems/unicycle.hpp"

namespace {

using altro::problems::UnicycleProblem;
constexpr int NStates = UnicycleProblem::NStates;
constexpr int NControls = UnicycleProblem::NControls;
using Clock = std::chrono::high_resolution_clock;

double Millis(Clock::time_point a, Clock::time_point b) {
  return std::chrono::duration<double, std::milli>(b - a).count();
}

// one instance, solved `nruns` times from the same initial guess
int SolveSingle(int nruns) {
  UnicycleProblem prob_def;
  prob_def.SetScenario(UnicycleProblem::kThreeObstacles);
  altro::problem::Problem prob = prob_def.MakeProblem(true);

  altro::augmented_lagrangian::AugmentedLagrangianiLQR<NStates, NControls> solver(prob);
  auto traj = std::make_shared<altro::Trajectory<NStates, NControls>>(prob_def.InitialTrajectory());
  solver.SetTrajectory(traj);

  for (int run = 0; run < nruns; ++run) {
    solver.SetPenalty(10.0);
    solver.GetOptions().verbose = altro::LogLevel::kSilent;
    *traj = prob_def.InitialTrajectory();

    const auto start = Clock::now();
    solver.Solve();
    const auto stop = Clock::now();
    std::printf("Iteration %d: Cost = %.10g, iters = %d, outer = %d, status = %d, viol = %.3e, Time = %.3f ms\n", run,
                solver.GetiLQRSolver().Cost(), solver.GetStats().iterations_total, solver.GetStats().iterations_outer,
                static_cast<int>(solver.GetStatus()), solver.MaxViolation(), Millis(start, stop));
  }
  const altro::VectorXd& xN = traj->State(prob_def.N);
  std::printf("Final state: %.6f %.6f %.6f\n", xN(0), xN(1), xN(2));
  return 0;
}

// splitmix64 -> uniform in [-1, 1): the same stream altro_cpp_b200/problems.py::uniform_batch draws
double Uniform(uint64_t* state) {
  uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<double>(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

// `batch` instances that differ in initial state, solved together
int SolveBatch(int nruns, int batch) {
  UnicycleProblem prob_def;
  prob_def.SetScenario(UnicycleProblem::kThreeObstacles);
  altro::problem::Problem prob = prob_def.MakeProblem(true);

  altro::augmented_lagrangian::BatchedAugmentedLagrangianiLQR<NStates, NControls> solver(prob, batch);
  solver.SetTrajectory(std::make_shared<altro::Trajectory<NStates, NControls>>(prob_def.InitialTrajectory()));

  std::vector<altro::VectorXd> x0(batch, prob_def.x0);
  uint64_t seed = 2021;
  const double scale[NStates] = {0.1, 0.1, 0.2};
  for (int b = 1; b < batch; ++b)  // instance 0 stays the nominal problem
    for (int i = 0; i < NStates; ++i) x0[b](i) += scale[i] * Uniform(&seed);
  solver.SetInitialStates(x0);

  for (int run = 0; run < nruns; ++run) {
    solver.SetPenalty(10.0);
    const auto start = Clock::now();
    solver.Solve();
    const auto stop = Clock::now();
    int solved = 0;
    int64_t iters = 0;
    for (int b = 0; b < batch; ++b) {
      solved += solver.GetStatus(b) == altro::SolverStatus::kSolved;
      iters += solver.GetIterations(b);
    }
    const double ms = Millis(start, stop);
    std::printf("Run %d: batch = %d, solved = %d, mean iters = %.2f, nominal iters = %d, Time = %.3f ms, %.1f solves/s\n",
                run, batch, solved, static_cast<double>(iters) / batch, solver.GetIterations(0), ms,
                1e3 * batch / ms);
  }
  return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
  const int nruns = argc > 1 ? std::stoi(argv[1]) : 1;
  const int batch = argc > 2 ? std::stoi(argv[2]) : 1;
  try {
    return batch > 1 ? SolveBatch(nruns, batch) : SolveSingle(nruns);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "benchmark_unicycle: %s\n", e.what());
    return 2;
  }
}
