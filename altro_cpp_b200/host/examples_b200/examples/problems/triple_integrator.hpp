// examples/problems/triple_integrator.hpp (shipped with the B200 host mirror) — the triple-integrator problem of
// the reference (examples/problems/triple_integrator.hpp:22-105): dof chains, LQR cost towards
// xf = (1..dof, 0...), optional control bounds and goal constraint.
#pragma once

#include <memory>
#include <vector>

#include "altro/augmented_lagrangian/al_problem.hpp"
#include "altro/augmented_lagrangian/al_solver.hpp"
#include "altro/common/trajectory.hpp"
#include "altro/ilqr/ilqr.hpp"
#include "altro/problem/discretized_model.hpp"
#include "altro/problem/problem.hpp"
#include "examples/basic_constraints.hpp"
#include "examples/quadratic_cost.hpp"
#include "examples/triple_integrator.hpp"

namespace altro {
namespace problems {

template <int dof = 2>
class TripleIntegratorProblem {
 public:
  static constexpr int NStates = 3 * dof;
  static constexpr int NControls = dof;
  using ModelType = problem::DiscretizedModel<examples::TripleIntegrator>;

  TripleIntegratorProblem() : model(examples::TripleIntegrator(dof)) {
    Q = MatrixXd::Zero(n, n);
    R = MatrixXd::Zero(m, m);
    Qf = MatrixXd::Zero(n, n);
    Q.diagonal().setConstant(1.0);
    R.diagonal().setConstant(0.001);
    Qf.diagonal().setConstant(1e5);
    xf = VectorXd::Zero(n);
    x0 = VectorXd::Zero(n);
    u0 = VectorXd::Zero(m);
    uref = VectorXd::Zero(m);
    for (int i = 0; i < dof; ++i) {
      xf(i) = i + 1;
      x0(i) = -(i + 1);
      lb.push_back(-100.0 * (i + 1));
      ub.push_back(+100.0 * (i + 1));
    }
  }

  const int n = NStates;
  const int m = NControls;
  int N = 10;
  float h = 0.1f;
  ModelType model;
  MatrixXd Q, R, Qf;
  VectorXd xf, x0, u0, uref;
  std::vector<double> lb, ub;

  float GetTimeStep() const { return h; }

  problem::Problem MakeProblem(bool add_constraints = false) {
    problem::Problem prob(N);
    auto qcost = std::make_shared<examples::QuadraticCost>(examples::QuadraticCost::LQRCost(Q, R, xf, uref));
    auto qterm = std::make_shared<examples::QuadraticCost>(examples::QuadraticCost::LQRCost(Qf, R * 0, xf, uref, true));
    auto dynamics = std::make_shared<ModelType>(model);
    for (int k = 0; k < N; ++k) {
      prob.SetCostFunction(qcost, k);
      prob.SetDynamics(dynamics, k);
    }
    prob.SetCostFunction(qterm, N);
    if (add_constraints) {
      auto bound = std::make_shared<examples::ControlBound>(lb, ub);
      for (int k = 0; k < N; ++k) prob.SetConstraint(bound, k);
      prob.SetConstraint(std::make_shared<examples::GoalConstraint>(xf), N);
    }
    prob.SetInitialState(x0);
    return prob;
  }

  template <int n_size = NStates, int m_size = NControls>
  Trajectory<n_size, m_size> InitialTrajectory() {
    Trajectory<n_size, m_size> Z(n, m, N);
    for (int k = 0; k < N; ++k) Z.Control(k) = u0;
    Z.SetUniformStep(GetTimeStep());
    return Z;
  }

  template <int n_size = NStates, int m_size = NControls>
  ilqr::iLQR<n_size, m_size> MakeSolver(bool alcost = false) {
    problem::Problem prob = MakeProblem(alcost);
    if (alcost) prob = augmented_lagrangian::BuildAugLagProblem<n_size, m_size>(prob);
    ilqr::iLQR<n_size, m_size> solver(prob);
    solver.SetTrajectory(std::make_shared<Trajectory<n_size, m_size>>(InitialTrajectory<n_size, m_size>()));
    solver.Rollout();
    return solver;
  }

  template <int n_size = NStates, int m_size = NControls>
  augmented_lagrangian::AugmentedLagrangianiLQR<n_size, m_size> MakeALSolver() {
    augmented_lagrangian::AugmentedLagrangianiLQR<n_size, m_size> solver_al(MakeProblem(true));
    solver_al.SetTrajectory(std::make_shared<Trajectory<n_size, m_size>>(InitialTrajectory<n_size, m_size>()));
    solver_al.GetiLQRSolver().Rollout();
    return solver_al;
  }
};

}  // namespace problems
}  // namespace altro
