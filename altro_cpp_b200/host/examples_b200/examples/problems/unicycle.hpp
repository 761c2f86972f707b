// examples/problems/unicycle.hpp (shipped with the B200 host mirror) — the unicycle problem definitions the
// reference's tests and perf/ programs are built on (examples/problems/unicycle.hpp:25-120,
// unicycle.cpp:12-88): scenario kTurn90 (control bounds + goal) and kThreeObstacles (BASELINE
// config C1/C2: three keep-out circles, control bounds, goal).  Header-only; the numbers are the
// ones altro_cpp_b200/problems.py::unicycle_problem uses, so C++ and Python callers describe the
// same device problem.
#pragma once

#include <cmath>
#include <memory>
#include <vector>

#include "altro/augmented_lagrangian/al_problem.hpp"
#include "altro/augmented_lagrangian/al_solver.hpp"
#include "altro/common/trajectory.hpp"
#include "altro/ilqr/ilqr.hpp"
#include "altro/problem/discretized_model.hpp"
#include "altro/problem/problem.hpp"
#include "examples/basic_constraints.hpp"
#include "examples/obstacle_constraints.hpp"
#include "examples/quadratic_cost.hpp"
#include "examples/unicycle.hpp"

namespace altro {
namespace problems {

class UnicycleProblem {
 public:
  static constexpr int NStates = 3;
  static constexpr int NControls = 2;
  using ModelType = problem::DiscretizedModel<examples::Unicycle>;
  using CostFunType = examples::QuadraticCost;
  enum Scenario { kTurn90, kThreeObstacles };

  UnicycleProblem() { Load(); }

  const int n = NStates;
  const int m = NControls;
  int N = 100;
  ModelType model = ModelType(examples::Unicycle());
  MatrixXd Q, R, Qf;
  VectorXd xf, x0, u0, uref;
  std::shared_ptr<examples::QuadraticCost> qcost, qterm;
  double v_bnd = 1.5;  // linear velocity bound
  double w_bnd = 1.5;  // angular velocity bound
  VectorXd cx, cy, cr;  // obstacle centres and radii
  std::vector<double> lb, ub;

  void SetScenario(Scenario scenario) {
    scenario_ = scenario;
    Load();
  }
  float GetTimeStep() const { return tf_ / N; }  // float on purpose

  problem::Problem MakeProblem(bool add_constraints = true) {
    Load();
    const float h = GetTimeStep();
    problem::Problem prob(N);

    // unicycle.cpp:52-60: the circles go in whenever the scenario has them, and before the bounds
    if (scenario_ == kThreeObstacles) {
      auto obstacles = std::make_shared<examples::CircleConstraint>();
      for (int i = 0; i < cx.size(); ++i) obstacles->AddObstacle(cx(i), cy(i), cr(i));
      for (int k = 1; k < N; ++k) prob.SetConstraint(obstacles, k);
    }
    // stage costs are scaled by the step, the terminal cost is not and has R = 0 (unicycle.cpp:63-69)
    qcost = std::make_shared<examples::QuadraticCost>(examples::QuadraticCost::LQRCost(Q * h, R * h, xf, uref));
    qterm = std::make_shared<examples::QuadraticCost>(examples::QuadraticCost::LQRCost(Qf, R * 0, xf, uref, true));
    for (int k = 0; k < N; ++k) prob.SetCostFunction(qcost, k);
    prob.SetCostFunction(qterm, N);

    auto dynamics = std::make_shared<ModelType>(model);
    for (int k = 0; k < N; ++k) prob.SetDynamics(dynamics, k);

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
    problem::Problem prob = MakeProblem();
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

 private:
  // scenario data, unicycle.cpp:12-50
  void Load() {
    uref = VectorXd::Zero(NControls);
    x0 = VectorXd::Zero(NStates);
    Q = MatrixXd::Zero(NStates, NStates);
    R = MatrixXd::Zero(NControls, NControls);
    Qf = MatrixXd::Zero(NStates, NStates);
    if (scenario_ == kTurn90) {
      tf_ = 3.0f;
      Q.diagonal().setConstant(1e-2);
      R.diagonal().setConstant(1e-2);
      Qf.diagonal().setConstant(100.0);
      xf = Eigen::Vector3d(1.5, 1.5, M_PI / 2);
      u0 = VectorXd::Constant(NControls, 0.1);
      lb = {-v_bnd, -w_bnd};
      ub = {+v_bnd, +w_bnd};
      cx = cy = cr = VectorXd();
    } else {
      tf_ = 5.0f;
      Q.diagonal().setConstant(1.0);
      R.diagonal().setConstant(0.5);
      Qf.diagonal().setConstant(10.0);
      xf = Eigen::Vector3d(3.0, 3.0, 0.0);
      u0 = VectorXd::Constant(NControls, 0.01);
      const double scaling = 3.0;
      cx = Eigen::Vector3d(0.25, 0.5, 0.75);
      cx *= scaling;
      cy = cx;
      cr = VectorXd::Constant(3, 0.425);
      lb = {0.0, -3.0};
      ub = {3.0, 3.0};
    }
  }

  Scenario scenario_ = kTurn90;
  float tf_ = 3.0f;
};

}  // namespace problems
}  // namespace altro
