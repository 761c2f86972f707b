// Reference-style tests written against the host mirror (altro_cpp_b200/host/include): they read like
// test/ilqr/unicycle_ilqr_test.cpp, test/augmented_lagrangian/auglag_test.cpp and
// test/examples/example_*_test.cpp of the reference and expect the same golden numbers, but every
// solver method runs on the device.
//
//   host_mirror_test cpu   checks that need no GPU (descriptor plumbing, loud failure)
//   host_mirror_test gpu   everything
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

#include "examples/problems/triple_integrator.hpp"
#include "examples/problems/unicycle.hpp"

namespace {

int failures = 0;
#define EXPECT(cond)                                                          \
  do {                                                                        \
    if (!(cond)) {                                                            \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);             \
      ++failures;                                                             \
    }                                                                         \
  } while (0)

using altro::SolverStatus;
using TripleIntegratorProblem = altro::problems::TripleIntegratorProblem<2>;
using altro::problems::UnicycleProblem;

using altro::MatrixXd;
using altro::VectorXd;
using altro::VectorXdRef;

// a user-defined model that is not in the device registry: the reference would call its virtual
// Evaluate on the host; here the solver must refuse it
class HostOnlyModel : public altro::problem::ContinuousDynamics {
 public:
  static constexpr int NStates = 3;
  static constexpr int NControls = 2;
  int StateDimension() const override { return 3; }
  int ControlDimension() const override { return 2; }
  bool HasHessian() const override { return false; }
  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, float, Eigen::Ref<VectorXd> xdot) override {
    xdot(0) = u(0) * x(1);
    xdot(1) = -x(0);
    xdot(2) = u(1);
  }
  void Jacobian(const VectorXdRef&, const VectorXdRef&, float, Eigen::Ref<MatrixXd> jac) override { jac.setZero(); }
  void Hessian(const VectorXdRef&, const VectorXdRef&, float, const VectorXdRef&, Eigen::Ref<MatrixXd> hess) override {
    hess.setZero();
  }
};

// Functors that only have the reference's virtual interface (no Describable, no accessors), like the
// classes of an unmodified altro-cpp checkout: the registry must recognise them by probing.
class PlainCircles : public altro::constraints::Constraint<altro::constraints::Inequality> {
 public:
  int OutputDimension() const override { return 2; }
  void Evaluate(const VectorXdRef& x, const VectorXdRef&, Eigen::Ref<VectorXd> c) override {
    c(0) = -(std::pow(x(1) - 0.75, 2) + std::pow(x(0) - 0.5, 2) - std::pow(0.425, 2));  // position = states (1, 0)
    c(1) = -(std::pow(x(1) - 2.25, 2) + std::pow(x(0) + 1.0, 2) - std::pow(0.3, 2));
  }
  void Jacobian(const VectorXdRef& x, const VectorXdRef&, Eigen::Ref<MatrixXd> jac) override {
    jac(0, 0) = 2 * (0.75 - x(1));
    jac(0, 1) = 2 * (0.5 - x(0));
    jac(1, 0) = 2 * (2.25 - x(1));
    jac(1, 1) = 2 * (-1.0 - x(0));
  }
};
class PlainBound : public altro::constraints::Constraint<altro::constraints::Inequality> {
 public:
  int OutputDimension() const override { return 3; }  // lb_0, ub_0, ub_1 finite
  void Evaluate(const VectorXdRef&, const VectorXdRef& u, Eigen::Ref<VectorXd> c) override {
    c(0) = -0.25 - u(0);
    c(1) = u(0) - 3.0;
    c(2) = u(1) - 1.5;
  }
  void Jacobian(const VectorXdRef&, const VectorXdRef&, Eigen::Ref<MatrixXd> jac) override {
    jac.setZero();
    jac(0, 3) = -1;
    jac(1, 3) = 1;
    jac(2, 4) = 1;
  }
};
class PlainCubicCost : public altro::problem::CostFunction {  // not a quadratic form
 public:
  int StateDimension() const override { return 3; }
  int ControlDimension() const override { return 2; }
  double Evaluate(const VectorXdRef& x, const VectorXdRef&) override { return x(0) * x(0) * x(0); }
  void Gradient(const VectorXdRef& x, const VectorXdRef&, Eigen::Ref<VectorXd> dx, Eigen::Ref<VectorXd> du) override {
    dx.setZero();
    du.setZero();
    dx(0) = 3 * x(0) * x(0);
  }
  void Hessian(const VectorXdRef& x, const VectorXdRef&, Eigen::Ref<MatrixXd> dxdx, Eigen::Ref<MatrixXd> dxdu,
               Eigen::Ref<MatrixXd> dudu) override {
    dxdx.setZero();
    dxdu.setZero();
    dudu.setZero();
    dxdx(0, 0) = 6 * x(0);
  }
};

void TestDescriptors() {
  UnicycleProblem def;
  def.SetScenario(UnicycleProblem::kThreeObstacles);
  altro::problem::Problem prob = def.MakeProblem(true);
  EXPECT(prob.IsFullyDefined());
  EXPECT(prob.NumSegments() == 100);
  EXPECT(prob.NumConstraints(0) == 4);    // control bounds only
  EXPECT(prob.NumConstraints(1) == 7);    // 3 circles + 4 bound rows
  EXPECT(prob.NumConstraints(100) == 3);  // goal
  EXPECT(def.GetTimeStep() == 5.0f / 100);

  std::string why;
  altro::device::ConstraintDesc d;
  EXPECT(altro::device::DescribeConstraint(*prob.GetInequalityConstraints()[1][0], 3, 2, &d, &why));
  EXPECT(d.kind == altro::device::ConstraintDesc::kCircle && d.a.size() == 3 && d.c[0] == 0.425 * 0.425);
  altro::device::CostDesc c;
  EXPECT(altro::device::DescribeCost(*prob.GetCostFunction(100), 3, 2, &c, &why));
  EXPECT(c.Q[0] == 10.0 && c.q[0] == -30.0 && c.R[0] == 0.0);

  // recognition by probing: the same functors seen ONLY through their virtual interface give the same
  // descriptions, bit for bit (what an unmodified altro-cpp checkout's example classes go through)
  {
    struct Veil : altro::constraints::Constraint<altro::constraints::Inequality> {  // hides Describable
      std::shared_ptr<altro::constraints::Constraint<altro::constraints::Inequality>> in;
      int OutputDimension() const override { return in->OutputDimension(); }
      void Evaluate(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> o) override { in->Evaluate(x, u, o); }
      void Jacobian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> j) override { in->Jacobian(x, u, j); }
    };
    Veil veiled;
    veiled.in = prob.GetInequalityConstraints()[1][0];
    altro::device::ConstraintDesc p;
    EXPECT(altro::device::DescribeConstraint(veiled, 3, 2, &p, &why));
    EXPECT(p.kind == d.kind && p.a == d.a && p.b == d.b && p.c == d.c && p.xi == 0 && p.yi == 1);
    veiled.in = prob.GetInequalityConstraints()[1][1];  // the control bound
    altro::device::ConstraintDesc pb, db;
    EXPECT(altro::device::DescribeConstraint(*veiled.in, 3, 2, &db, &why));
    EXPECT(altro::device::DescribeConstraint(veiled, 3, 2, &pb, &why));
    EXPECT(pb.kind == altro::device::ConstraintDesc::kControlBound && pb.a == db.a && pb.b == db.b);

    PlainCircles circles;
    EXPECT(altro::device::DescribeConstraint(circles, 3, 2, &p, &why));
    EXPECT(p.kind == altro::device::ConstraintDesc::kCircle && p.xi == 1 && p.yi == 0);
    EXPECT(p.a[0] == 0.75 && p.b[0] == 0.5 && p.c[0] == std::pow(0.425, 2) && p.a[1] == 2.25 && p.b[1] == -1.0);
    PlainBound bound;
    EXPECT(altro::device::DescribeConstraint(bound, 3, 2, &p, &why));
    EXPECT(p.kind == altro::device::ConstraintDesc::kControlBound && p.a[0] == -0.25 && std::isinf(p.a[1]) &&
           p.b[0] == 3.0 && p.b[1] == 1.5);
    PlainCubicCost cubic;
    altro::device::CostDesc pc;
    EXPECT(!altro::device::DescribeCost(cubic, 3, 2, &pc, &why));
    EXPECT(why.find("not a quadratic form") != std::string::npos);
  }

  // ControlBound drops infinite rows (basic_constraints.hpp:136-143 there)
  altro::examples::ControlBound half(2);
  half.SetUpperBound({1.0, 2.0});
  EXPECT(half.OutputDimension() == 2);
  bool threw = false;
  try {
    altro::examples::ControlBound bad({1.0}, {0.0});
  } catch (const std::invalid_argument&) {
    threw = true;
  }
  EXPECT(threw);

  // a functor that cannot be described is rejected before any device work
  altro::problem::Problem custom(10);
  auto model = std::make_shared<altro::problem::DiscretizedModel<HostOnlyModel>>(HostOnlyModel());
  auto cost = std::make_shared<altro::examples::QuadraticCost>(altro::examples::QuadraticCost::LQRCost(
      altro::MatrixXd::Identity(3, 3), altro::MatrixXd::Identity(2, 2), altro::VectorXd::Zero(3),
      altro::VectorXd::Zero(2)));
  for (int k = 0; k <= 10; ++k) custom.SetCostFunction(cost, k);
  for (int k = 0; k < 10; ++k) custom.SetDynamics(model, k);
  custom.SetInitialState(altro::VectorXd::Zero(3));
  altro::ilqr::iLQR<3, 2> solver(custom);
  auto Z = std::make_shared<altro::Trajectory<3, 2>>(3, 2, 10);
  Z->SetUniformStep(0.1f);
  int code = 0;
  try {
    solver.SetTrajectory(Z);
  } catch (const altro::DeviceError& e) {
    code = e.code;
  }
  EXPECT(code == ALTRO_B200_ERR_UNSUPPORTED);
}

// The Eigen stand-in's factorisations (the reference's tests compare gains against ldlt().solve(); QuadraticCost
// validation reads the signs of vectorD()): residuals on a matrix that needs pivoting, and on an SPD one.
void TestStandInFactorisations() {
  MatrixXd A(3, 3);
  A << 1e-8, 2.0, 0.0,  //
      2.0, 1.0, 3.0,    //
      0.0, 3.0, -4.0;
  MatrixXd rhs(3, 2);
  rhs << 1.0, -2.0, 0.5, 4.0, -3.0, 0.25;
  const auto ldlt = A.ldlt();
  const MatrixXd x = ldlt.solve(rhs);
  EXPECT((A * x - rhs).norm() < 1e-12);
  int negative = 0;
  for (int i = 0; i < 3; ++i) negative += ldlt.vectorD()(i) < 0.0;
  EXPECT(negative == 2);  // det(A) > 0 and a negative diagonal entry: inertia (1, 2)

  MatrixXd S = A.transpose() * A + MatrixXd::Identity(3, 3);
  EXPECT((S * S.llt().solve(rhs) - rhs).norm() < 1e-10);
  EXPECT((S * S.ldlt().solve(rhs) - rhs).norm() < 1e-10);
  EXPECT(S.llt().info() == Eigen::Success && A.llt().info() != Eigen::Success);
}

// The AL terms of one knot evaluated on the host (ALCost built from a problem, not from a solver) against the
// closed form for a goal constraint: J = l(x,u) + lambda'(x - xf)... with the sign convention of the reference,
// (|lambda - rho c|^2 - |lambda|^2) / (2 rho).
void TestStandaloneALCost() {
  UnicycleProblem def;
  altro::problem::Problem prob = def.MakeProblem();
  const int N = prob.NumSegments();
  altro::augmented_lagrangian::ALCost<3, 2> term(prob, N);
  EXPECT(term.NumConstraints() == 3 && term.GetEqualityConstraints().size() == 1);
  const double rho = 7.0;
  term.SetPenalty<altro::constraints::Equality>(rho);
  VectorXd& lambda = term.GetEqualityConstraints()[0]->GetDuals();
  lambda << 0.3, -0.2, 0.1;
  VectorXd x(3), u(2);
  x << 0.4, 1.2, 0.9;
  u << 0.0, 0.0;
  VectorXd c(3);
  prob.GetEqualityConstraints()[N][0]->Evaluate(x, u, c);
  double expected = prob.GetCostFunction(N)->Evaluate(x, u);
  for (int i = 0; i < 3; ++i) expected += (std::pow(lambda(i) - rho * c(i), 2) - lambda(i) * lambda(i)) / (2 * rho);
  EXPECT(std::fabs(term.Evaluate(x, u) - expected) < 1e-12);
  // the problem BuildAugLagProblem returns evaluates the same terms and has no constraints of its own
  altro::problem::Problem prob_al = altro::augmented_lagrangian::BuildAugLagProblem<3, 2>(prob);
  EXPECT(prob_al.NumConstraints() == 0 && prob.NumConstraints() > 0 && prob_al.IsAugmentedLagrangian());
  double unit = prob.GetCostFunction(N)->Evaluate(x, u);
  for (int i = 0; i < 3; ++i) unit += 0.5 * c(i) * c(i);  // multipliers zero, penalty one
  EXPECT(std::fabs(prob_al.GetCostFunction(N)->Evaluate(x, u) - unit) < 1e-12);
}

// without a GPU every solver reports the CUDA error instead of computing on the host
void TestNoDeviceIsLoud() {
  UnicycleProblem def;
  int code = 0;
  std::string what;
  try {
    auto solver = def.MakeSolver();
  } catch (const altro::DeviceError& e) {
    code = e.code;
    what = e.what();
  }
  EXPECT(code == ALTRO_B200_ERR_CUDA);
  EXPECT(what.find("no CPU fallback") != std::string::npos);
}

// test/ilqr/unicycle_ilqr_test.cpp:28-100
void TestUnicycleILQR() {
  UnicycleProblem def;
  auto solver = def.MakeSolver();
  const double J0 = solver.Cost();
  EXPECT(std::fabs(J0 - 259.27636137767087) < 1e-5);
  solver.UpdateExpansions();
  solver.BackwardPass();
  solver.ForwardPass();
  EXPECT(solver.Cost() < J0);

  // unicycle_ilqr_test.cpp:56-65: the first line search accepts alpha = 1/16, logged in row 0 of the stats
  EXPECT(solver.GetStats().alpha.size() == 1 && solver.GetStats().alpha[0] == 0.0625);
  solver.UpdateConvergenceStatistics();
  EXPECT(solver.GetStats().iterations_inner == 1 && solver.GetStats().cost_decrease.size() == 2);
  // without Solve() nobody set stats.initial_cost: dJ = 0 - J (ilqr.hpp:573-574), exactly like the reference
  EXPECT(solver.GetStats().gradient[0] > 0.0);
  EXPECT(std::fabs(solver.GetStats().cost_decrease[0] - (solver.GetStats().initial_cost - solver.Cost())) <
         1e-9 * (1.0 + std::fabs(solver.Cost())));

  // unicycle_ilqr_test.cpp:40-54: cost-to-go gradient and feedforward gain at knot 0
  auto step = def.MakeSolver();
  step.UpdateExpansions();
  step.BackwardPass();
  auto& kpf = step.GetKnotPointFunction(0);
  const double p0[3] = {0.024904637422419617, -0.46496022574032614, -0.0573096310550007};
  const double d0[2] = {-2.565783457444465, 5.514158930898376};
  for (int i = 0; i < 3; ++i) EXPECT(std::fabs(kpf.GetCostToGoGradient()(i) - p0[i]) < 1e-5 * 0.47);
  for (int i = 0; i < 2; ++i) EXPECT(std::fabs(kpf.GetFeedforwardGain()(i) - d0[i]) < 1e-5 * 6.1);
  EXPECT(kpf.GetDynamicsExpansion().GetA().rows() == 3 && kpf.GetDynamicsExpansion().GetB().cols() == 2);
  EXPECT(kpf.GetDynamicsExpansion().GetA()(0, 0) == 1.0);  // d x_next / d x = 1 for the unicycle
  EXPECT(kpf.GetCostExpansion().dudu()(0, 0) > 0.0);
  EXPECT(step.NumThreads() == 0 && step.GetTaskAssignment().back() == def.N + 1);  // no pool for one thread

  // the problem's initial state is shared with the solver (ilqr_class_test.cpp:84-96)
  {
    altro::problem::Problem prob = def.MakeProblem();
    altro::ilqr::iLQR<3, 2> late(def.N);
    late.InitializeFromProblem(prob);
    auto Zl = std::make_shared<altro::Trajectory<3, 2>>(def.InitialTrajectory());
    late.SetTrajectory(Zl);
    altro::VectorXd moved = prob.GetInitialState();
    moved(0) = 0.25;
    prob.SetInitialState(moved);
    EXPECT((*late.GetInitialState())(0) == 0.25);
    late.Rollout();
    EXPECT(Zl->State(0)(0) == 0.25);
  }

  auto fresh = def.MakeSolver();
  fresh.Solve();
  EXPECT(fresh.GetStatus() == SolverStatus::kSolved);
  EXPECT(fresh.GetStats().iterations_inner == 9);
  EXPECT(std::fabs(fresh.Cost() - 0.0387016567) < 1e-5);
  {  // unicycle_ilqr_test.cpp:96-99 and solver_stats.hpp:54-61: one row per iteration + the open slot
    const altro::SolverStats& st = fresh.GetStats();
    EXPECT(st.cost.size() == 10 && st.alpha.size() == 10 && st.gradient.size() == 10 && st.regularization.size() == 10);
    EXPECT(st.alpha[0] == 0.0625);
    EXPECT(st.cost_decrease.back() < fresh.GetOptions().cost_tolerance);
    EXPECT(st.gradient.back() < fresh.GetOptions().gradient_tolerance);
    EXPECT(std::fabs(st.cost[8] - 0.0387016567) < 1e-5 && st.cost[9] == st.cost[8]);
    for (int i = 1; i < 9; ++i) EXPECT(st.cost[i] <= st.cost[i - 1]);
    EXPECT(std::fabs(st.initial_cost - 259.27636137767087) < 1e-5);
    EXPECT(std::fabs((st.initial_cost - st.cost[0]) - st.cost_decrease[0]) < 1e-9 * st.initial_cost);
  }
  const altro::VectorXd& xN = fresh.GetTrajectory()->State(def.N);
  EXPECT(std::fabs(xN(0) - 1.5) < 1e-2 && std::fabs(xN(1) - 1.5) < 1e-2);
  auto& g = fresh.GetKnotPointFunction(0);
  EXPECT(g.GetFeedbackGain().rows() == 2 && g.GetFeedbackGain().cols() == 3);
}

// a trajectory with per-knot steps (Trajectory::SetStep / SetTime) is solved on its own time grid
void TestNonUniformSteps() {
  UnicycleProblem def;
  auto uniform = def.MakeALSolver();
  uniform.Solve();
  const double J_uniform = uniform.GetiLQRSolver().Cost();
  const int iters_uniform = uniform.GetStats().iterations_total;

  // the same grid written knot by knot travels to the device unchanged: identical solve
  {
    auto solver = def.MakeALSolver();
    auto Z = solver.GetiLQRSolver().GetTrajectory();
    const int N = Z->NumSegments();
    const float h = Z->GetStep(0);
    for (int k = 0; k <= N; ++k) {
      Z->SetTime(k, k < N ? static_cast<float>(k) * h : static_cast<float>(h) * N);
      Z->SetStep(k, k < N ? h : 0.0F);
    }
    solver.Solve();
    EXPECT(solver.GetiLQRSolver().Cost() == J_uniform);
    EXPECT(solver.GetStats().iterations_total == iters_uniform);
  }
  // a stretched grid (steps growing from 0.8 h to 1.2 h) is a different discretisation: different optimum
  {
    auto solver = def.MakeALSolver();
    auto Z = solver.GetiLQRSolver().GetTrajectory();
    const int N = Z->NumSegments();
    const float h0 = Z->GetStep(0);
    float t = 0.0F;
    for (int k = 0; k <= N; ++k) {
      const float h = k < N ? h0 * (0.8F + 0.4F * static_cast<float>(k) / static_cast<float>(N)) : 0.0F;
      Z->SetTime(k, t);
      Z->SetStep(k, h);
      t += h;
    }
    EXPECT(Z->CheckTimeConsistency());
    solver.Solve();
    const double J = solver.GetiLQRSolver().Cost();
    EXPECT(std::isfinite(J) && J > 0.0);
    EXPECT(std::fabs(J - J_uniform) > 1e-9 * J_uniform);
    if (solver.GetStatus() != SolverStatus::kSolved)
      std::printf("note: stretched grid ended with status %d after %d iterations, J = %.6g (uniform: %.6g)\n",
                  static_cast<int>(solver.GetStatus()), solver.GetStats().iterations_total, J, J_uniform);
  }
}

// test/augmented_lagrangian/auglag_test.cpp:326-380
void TestUnicycleAugLag() {
  UnicycleProblem def;
  auto solver = def.MakeALSolver();
  solver.GetOptions().constraint_tolerance = 1e-6;
  for (int repeat = 0; repeat < 2; ++repeat) {
    *solver.GetiLQRSolver().GetTrajectory() = def.InitialTrajectory();
    solver.Solve();
    EXPECT(solver.GetStatus() == SolverStatus::kSolved);
    EXPECT(solver.GetStats().iterations_total == 14);
    EXPECT(solver.GetStats().iterations_outer == 5);
    {  // example_unicycle_test.cpp:87-88 + the AL log into the open slot (al_solver.hpp:361-362)
      const altro::SolverStats& st = solver.GetStats();
      EXPECT(st.cost.size() == 15 && st.violations.size() == 15 && st.max_penalty.size() == 15);
      EXPECT(st.cost_decrease.back() < solver.GetOptions().cost_tolerance);
      EXPECT(st.gradient.back() < solver.GetOptions().gradient_tolerance);
      EXPECT(st.violations.back() < 1e-6 && st.violations.back() >= 0.0);
      EXPECT(st.max_penalty.back() == solver.GetMaxPenalty() && st.max_penalty.back() > st.max_penalty.front());
    }
    EXPECT(solver.MaxViolation() < 1e-6);
    const double J = solver.GetiLQRSolver().Cost();
    EXPECT(std::fabs(J - 0.03893465058924039) / 0.03893465058924039 < 1e-9);
  }
}

// iLQR::Solve() spelled out with the public step methods (ilqr.hpp:284-316 there) walks the same path as the
// one-launch Solve(): verdict, iteration count, cost, trajectory; GetCosts() adds up to Cost().
void TestInnerLoopByHand() {
  UnicycleProblem def;
  auto whole = def.MakeSolver();
  whole.Solve();
  const double J_whole = whole.Cost();

  auto steps = def.MakeSolver();
  steps.SolveSetup();
  steps.Rollout();
  steps.GetStats().initial_cost = steps.Cost();
  EXPECT(std::fabs(steps.GetStats().initial_cost - 259.27636137767087) < 1e-5);
  int iterations = 0;
  for (; iterations < steps.GetOptions().max_iterations_inner; ++iterations) {
    steps.UpdateExpansions();
    steps.BackwardPass();
    steps.ForwardPass();
    steps.UpdateConvergenceStatistics();
    if (iterations == 0) {  // the first decrease is measured against the initial cost the caller assigned
      const altro::SolverStats& st = steps.GetStats();
      EXPECT(std::fabs(st.cost_decrease[0] - (st.initial_cost - steps.Cost())) < 1e-9 * st.initial_cost);
      const double grad = steps.NormalizedFeedforwardGain();
      EXPECT(std::fabs(grad - st.gradient[0]) < 1e-12 * (1.0 + grad));
    }
    if (steps.IsDone()) break;
  }
  steps.WrapUp();
  EXPECT(steps.GetStatus() == SolverStatus::kSolved && whole.GetStatus() == SolverStatus::kSolved);
  EXPECT(steps.GetStats().iterations_inner == whole.GetStats().iterations_inner);
  EXPECT(iterations + 1 == whole.GetStats().iterations_inner);
  const double J_steps = steps.Cost();
  EXPECT(J_steps == J_whole);
  if (J_steps != J_whole) std::printf("  by hand %.17g, whole solve %.17g\n", J_steps, J_whole);
  double worst = 0.0;
  for (int k = 0; k <= def.N; ++k)
    worst = std::max(worst, (whole.GetTrajectory()->State(k) - steps.GetTrajectory()->State(k)).norm());
  EXPECT(worst == 0.0);
  const VectorXd& costs = steps.GetCosts();
  EXPECT(costs.size() == def.N + 1);
  double sum = 0.0;
  for (int k = 0; k <= def.N; ++k) sum += costs(k);
  EXPECT(sum == J_steps);
}

// The outer loop driven by the caller with the reference's public step methods (al_solver.hpp:304-334 there) must
// walk the same path as Solve(), which runs it in one launch: same verdict, iteration counts, cost and trajectory.
void TestOuterLoopByHand() {
  UnicycleProblem def;
  auto whole = def.MakeALSolver();
  whole.GetOptions().constraint_tolerance = 1e-6;
  *whole.GetiLQRSolver().GetTrajectory() = def.InitialTrajectory();
  whole.Solve();
  const double J_whole = whole.GetiLQRSolver().Cost();

  auto steps = def.MakeALSolver();
  steps.GetOptions().constraint_tolerance = 1e-6;
  *steps.GetiLQRSolver().GetTrajectory() = def.InitialTrajectory();
  steps.Init();
  EXPECT(steps.GetStats().iterations_outer == 0 && steps.GetStats().violations.size() == 1);
  EXPECT(steps.GetStatus() == SolverStatus::kUnsolved);
  int passes = 0;
  for (; passes < steps.GetOptions().max_iterations_outer; ++passes) {
    steps.GetiLQRSolver().Solve();
    steps.UpdateDuals();
    steps.UpdateConvergenceStatistics();
    if (steps.IsDone()) break;
    steps.UpdatePenalties();
  }
  EXPECT(steps.GetStatus() == whole.GetStatus() && steps.GetStatus() == SolverStatus::kSolved);
  EXPECT(passes + 1 == whole.GetStats().iterations_outer);
  EXPECT(steps.GetStats().iterations_outer == whole.GetStats().iterations_outer);
  EXPECT(steps.GetStats().iterations_total == whole.GetStats().iterations_total);
  EXPECT(steps.GetMaxPenalty() == whole.GetMaxPenalty());
  const double J_steps = steps.GetiLQRSolver().Cost();
  EXPECT(J_steps == J_whole);
  if (J_steps != J_whole) std::printf("  by hand %.17g, whole solve %.17g\n", J_steps, J_whole);
  const auto Za = whole.GetiLQRSolver().GetTrajectory(), Zb = steps.GetiLQRSolver().GetTrajectory();
  double worst = 0.0;
  for (int k = 0; k <= def.N; ++k) worst = std::max(worst, (Za->State(k) - Zb->State(k)).norm());
  EXPECT(worst == 0.0);

  // ResetDualVariables: every multiplier back to zero, the penalty untouched
  const double pen = steps.GetMaxPenalty();
  steps.ResetDualVariables();
  EXPECT(steps.GetDuals(def.N).norm() == 0.0 && steps.GetDuals(1).norm() == 0.0);
  EXPECT(steps.GetMaxPenalty() == pen);
  // MaxViolation(Z) of another trajectory: the initial guess violates the goal constraint
  EXPECT(steps.MaxViolation(def.InitialTrajectory()) > 1e-2);
}

// test/examples/example_unicycle_test.cpp:69-89 (BASELINE config C1)
void TestThreeObstacles() {
  UnicycleProblem def;
  def.SetScenario(UnicycleProblem::kThreeObstacles);
  altro::augmented_lagrangian::AugmentedLagrangianiLQR<3, 2> solver(def.MakeProblem(true));
  auto Z = std::make_shared<altro::Trajectory<3, 2>>(def.InitialTrajectory());
  solver.SetTrajectory(Z);
  solver.SetPenalty(10.0);
  solver.Solve();
  EXPECT(solver.GetStatus() == SolverStatus::kSolved);
  EXPECT(solver.GetStats().iterations_total == 50);
  EXPECT(solver.GetStats().iterations_outer == 5);
  EXPECT(solver.MaxViolation() < 1e-4);
  for (int i = 0; i < 3; ++i) {
    altro::examples::Circle c(def.cx(i), def.cy(i), def.cr(i));
    for (int k = 0; k <= def.N; ++k) EXPECT(c.Distance(Z->State(k)(0), Z->State(k)(1)) > -1e-3);
  }
  for (int i = 0; i < 3; ++i) EXPECT(std::fabs(Z->State(def.N)(i) - def.xf(i)) < 1e-4);

  // al_solver.hpp:68-104: one ConstraintInfo per constraint and knot point
  EXPECT(solver.NumConstraints() == 4 * 100 + 3 * 99 + 3);
  EXPECT(solver.NumConstraints(1) == 7);
  auto coninfo = solver.GetConstraintInfo();
  EXPECT(coninfo.size() == 100 + 99 + 1);
  EXPECT(coninfo.front().index == 0 && coninfo.front().label == "Control Bound");
  EXPECT(coninfo[1].label == "Circle Constraint" && coninfo[1].index == 1 && coninfo[1].violation.size() == 3);
  EXPECT(coninfo.back().label == "Goal Constraint" && coninfo.back().type == "Equality Constraint");
  double worst = 0.0;
  for (const auto& info : coninfo)
    for (int i = 0; i < info.violation.size(); ++i) worst = std::fmax(worst, std::fabs(info.violation(i)));
  EXPECT(std::fabs(worst - solver.MaxViolation()) < 1e-15);  // two device code sites, FMA contraction may differ
  auto sorted = solver.GetConstraintInfo(true);
  double first = 0.0;
  for (int i = 0; i < sorted.front().violation.size(); ++i) first = std::fmax(first, std::fabs(sorted.front().violation(i)));
  EXPECT(std::fabs(first - worst) < 1e-15);
  EXPECT(sorted.front().ToString().find(" at index ") != std::string::npos);
  EXPECT(solver.GetMaxPenalty() >= 1.0 && solver.GetMaxPenalty() <= 1e8);
  EXPECT(solver.GetDuals(def.N).size() == 3 && solver.GetDuals(0).size() == 4);

  // the batched solver gives the nominal instance the same answer
  altro::augmented_lagrangian::BatchedAugmentedLagrangianiLQR<3, 2> batched(def.MakeProblem(true), 64);
  batched.SetTrajectory(std::make_shared<altro::Trajectory<3, 2>>(def.InitialTrajectory()));
  std::vector<altro::VectorXd> x0(64, def.x0);
  for (int b = 1; b < 64; ++b) x0[b](1) += 0.002 * b;
  batched.SetInitialStates(x0);
  batched.SetPenalty(10.0);
  batched.Solve();
  EXPECT(batched.GetStatus(0) == SolverStatus::kSolved);
  EXPECT(batched.GetIterations(0) == 50 && batched.GetOuterIterations(0) == 5);
  altro::Trajectory<3, 2> Z0 = batched.GetTrajectory(0);
  double diff = 0.0;
  for (int k = 0; k <= def.N; ++k)
    for (int i = 0; i < 3; ++i) diff = std::fmax(diff, std::fabs(Z0.State(k)(i) - Z->State(k)(i)));
  EXPECT(diff == 0.0);
  EXPECT(batched.KernelLaunches() > 0);

  // the same batch cut into two slices (two solvers + host threads; here both on device 0): every
  // instance gets the same bits wherever it is solved
  altro::augmented_lagrangian::BatchedAugmentedLagrangianiLQR<3, 2> sharded(def.MakeProblem(true), 64, std::vector<int>{0, 0});
  sharded.SetTrajectory(std::make_shared<altro::Trajectory<3, 2>>(def.InitialTrajectory()));
  sharded.SetInitialStates(x0);
  sharded.GetOptions().initial_penalty = 1.0;
  sharded.Solve();
  for (int b : {0, 31, 32, 63}) {
    EXPECT(sharded.GetIterations(b) == batched.GetIterations(b) && sharded.GetStatus(b) == batched.GetStatus(b));
    EXPECT(sharded.GetCost(b) == batched.GetCost(b));
    altro::Trajectory<3, 2> Za = batched.GetTrajectory(b), Zb = sharded.GetTrajectory(b);
    double dz = 0.0;
    for (int k = 0; k <= def.N; ++k)
      for (int i = 0; i < 3; ++i) dz = std::fmax(dz, std::fabs(Za.State(k)(i) - Zb.State(k)(i)));
    EXPECT(dz == 0.0);
  }
}

// test/ilqr/ilqr_test.cpp:304-336, test/examples/example_triple_integrator_test.cpp:16-70
void TestTripleIntegrator() {
  TripleIntegratorProblem def;
  auto solver = def.MakeSolver();
  solver.Solve();
  EXPECT(solver.GetStatus() == SolverStatus::kSolved);
  EXPECT(solver.GetStats().iterations_inner == 2);

  auto al = def.MakeALSolver();
  al.Solve();
  EXPECT(al.GetStatus() == SolverStatus::kSolved);
  EXPECT(al.MaxViolation() < 1e-4);
  auto Z = al.GetiLQRSolver().GetTrajectory();
  for (int i = 0; i < 6; ++i) EXPECT(std::fabs(Z->State(10)(i) - def.xf(i)) < 1e-4);
  EXPECT(std::fabs(Z->Control(0)(0) - 100.0) < 1e-4 && std::fabs(Z->Control(0)(1) - 200.0) < 1e-4);
}

}  // namespace

int main(int argc, char* argv[]) {
  const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
  try {
    TestDescriptors();
    TestStandInFactorisations();
    TestStandaloneALCost();
    if (!gpu) {
      TestNoDeviceIsLoud();
    } else {
      TestUnicycleILQR();
      TestInnerLoopByHand();
      TestUnicycleAugLag();
      TestOuterLoopByHand();
      TestNonUniformSteps();
      TestThreeObstacles();
      TestTripleIntegrator();
    }
  } catch (const std::exception& e) {
    std::printf("FAIL unexpected exception: %s\n", e.what());
    return 3;
  }
  std::printf("%s: %d failure(s)\n", gpu ? "gpu" : "cpu", failures);
  return failures ? 1 : 0;
}
