// oracle/ref_shim/ref_driver.cpp — a C entry point on top of the REFERENCE's own classes, compiled by
// oracle/build_ref.py together with the reference's unmodified sources (read where they lie, /root/reference) into
// oracle/_ref/libaltro_ref.so.  Test infrastructure: tests/test_oracle_vs_reference_build.py solves the same
// problems with this library and with the oracle restatement (oracle/altro_oracle.cpp) and compares them.
//
// What it is and is not: the control flow, the problem definitions (examples/problems/*.hpp) and every formula are
// the reference's own code; Eigen is absent from this image, so the linear algebra underneath is this repo's
// Eigen stand-in (altro_cpp_b200/host/include/eigen3/Eigen/Dense: eager evaluation, plain triple loops, unblocked
// LLT).  Results therefore carry the reference's logic with the stand-in's rounding order.
#include <cstring>
#include <memory>

#include "altro/augmented_lagrangian/al_solver.hpp"
#include "altro/ilqr/ilqr.hpp"
#include "examples/problems/triple_integrator.hpp"
#include "examples/problems/unicycle.hpp"

namespace {

struct Outputs {
  double* X;      // [(N+1) * n]
  double* U;      // [N * m]
  double* scalars;  // cost, max violation, max penalty, initial cost
  int* counters;    // status, iterations_inner, iterations_outer, iterations_total
};

template <int n, int m>
void Export(altro::Trajectory<n, m>& Z, int N, const Outputs& out) {
  for (int k = 0; k <= N; ++k)
    for (int i = 0; i < n; ++i) out.X[k * n + i] = Z.State(k)(i);
  for (int k = 0; k < N; ++k)
    for (int i = 0; i < m; ++i) out.U[k * m + i] = Z.Control(k)(i);
}

// options: [0] constraint_tolerance, [1] SetPenalty value, [2] initial_penalty, [3] max_iterations_total,
//          [4] max_iterations_inner, [5] max_iterations_outer; a negative entry keeps the reference's default
void Apply(altro::SolverOptions& o, const double* options) {
  o.verbose = altro::LogLevel::kSilent;
  if (options == nullptr) return;
  if (options[0] >= 0) o.constraint_tolerance = options[0];
  if (options[2] >= 0) o.initial_penalty = options[2];
  if (options[3] >= 0) o.max_iterations_total = static_cast<int>(options[3]);
  if (options[4] >= 0) o.max_iterations_inner = static_cast<int>(options[4]);
  if (options[5] >= 0) o.max_iterations_outer = static_cast<int>(options[5]);
}

template <int n, int m>
void SolveConstrained(const altro::problem::Problem& prob, std::shared_ptr<altro::Trajectory<n, m>> Z, const double* options,
                      const Outputs& out) {
  altro::augmented_lagrangian::AugmentedLagrangianiLQR<n, m> solver(prob);
  solver.SetTrajectory(Z);
  Apply(solver.GetOptions(), options);
  if (options != nullptr && options[1] >= 0) solver.SetPenalty(options[1]);
  solver.Solve();
  Export<n, m>(*Z, prob.NumSegments(), out);
  out.scalars[1] = solver.GetMaxViolation();  // of the constraint values the solve left behind: before Cost() refreshes them
  out.scalars[2] = solver.GetMaxPenalty();
  out.scalars[0] = solver.GetiLQRSolver().Cost();
  out.scalars[3] = solver.GetStats().initial_cost;
  out.counters[0] = static_cast<int>(solver.GetStatus());
  out.counters[1] = solver.GetStats().iterations_inner;
  out.counters[2] = solver.GetStats().iterations_outer;
  out.counters[3] = solver.GetStats().iterations_total;
}

template <int n, int m>
void SolveUnconstrained(const altro::problem::Problem& prob, std::shared_ptr<altro::Trajectory<n, m>> Z, const double* options,
                        const Outputs& out) {
  altro::ilqr::iLQR<n, m> solver(prob);
  solver.SetTrajectory(Z);
  Apply(solver.GetOptions(), options);
  solver.Solve();
  Export<n, m>(*Z, prob.NumSegments(), out);
  out.scalars[0] = solver.Cost();
  out.scalars[1] = 0.0;
  out.scalars[2] = 0.0;
  out.scalars[3] = solver.GetStats().initial_cost;
  out.counters[0] = static_cast<int>(solver.GetStatus());
  out.counters[1] = solver.GetStats().iterations_inner;
  out.counters[2] = 0;
  out.counters[3] = solver.GetStats().iterations_total;
}

}  // namespace

extern "C" {

// examples/problems/unicycle.hpp: scenario 0 = kTurn90, 1 = kThreeObstacles; N = 100
int altro_ref_unicycle(int scenario, int constrained, const double* x0, const double* options, double* X, double* U,
                       double* scalars, int* counters) {
  altro::problems::UnicycleProblem def;
  def.SetScenario(scenario == 1 ? altro::problems::UnicycleProblem::kThreeObstacles
                                : altro::problems::UnicycleProblem::kTurn90);
  altro::problem::Problem prob = def.MakeProblem(constrained != 0);  // a scenario may set its own x0 in there
  if (x0 != nullptr) prob.SetInitialState(Eigen::Vector3d(x0[0], x0[1], x0[2]));
  auto Z = std::make_shared<altro::Trajectory<3, 2>>(def.InitialTrajectory());
  const Outputs out{X, U, scalars, counters};
  if (constrained != 0) SolveConstrained<3, 2>(prob, Z, options, out);
  else SolveUnconstrained<3, 2>(prob, Z, options, out);
  return def.N;
}

// examples/problems/triple_integrator.hpp with dof = 2 (n = 6, m = 2): N knots of step 0.1
int altro_ref_triple_integrator(int N, int constrained, const double* x0, const double* options, double* X, double* U,
                                double* scalars, int* counters) {
  altro::problems::TripleIntegratorProblem<2> def;
  def.N = N;
  if (x0 != nullptr)
    for (int i = 0; i < 6; ++i) def.x0(i) = x0[i];
  altro::problem::Problem prob = def.MakeProblem(constrained != 0);
  prob.SetInitialState(def.x0);
  auto Z = std::make_shared<altro::Trajectory<6, 2>>(def.InitialTrajectory());
  const Outputs out{X, U, scalars, counters};
  if (constrained != 0) SolveConstrained<6, 2>(prob, Z, options, out);
  else SolveUnconstrained<6, 2>(prob, Z, options, out);
  return def.N;
}

}  // extern "C"
