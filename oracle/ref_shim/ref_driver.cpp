// oracle/ref_shim/ref_driver.cpp — a C entry point on top of the REFERENCE's own classes, compiled by
// oracle/build_ref.py together with the reference's unmodified sources (read where they lie, /root/reference) into
// oracle/_ref/libaltro_ref.so.  Test infrastructure: tests/test_oracle_vs_reference_build.py solves the same
// problems with this library and with the oracle restatement (oracle/altro_oracle.cpp) and compares them.
//
// What it is and is not: the control flow, the problem definitions (examples/problems/*.hpp) and every formula are
// the reference's own code; Eigen is absent from this image, so the linear algebra underneath is this repo's
// Eigen stand-in (altro_cpp_b200/host/include/eigen3/Eigen/Dense: eager evaluation, plain triple loops, unblocked
// LLT).  Results therefore carry the reference's logic with the stand-in's rounding order.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "altro/augmented_lagrangian/al_solver.hpp"
#include "altro/ilqr/ilqr.hpp"
#include "altro/problem/discretized_model.hpp"
#include "examples/basic_constraints.hpp"
#include "examples/obstacle_constraints.hpp"
#include "examples/problems/triple_integrator.hpp"
#include "examples/problems/unicycle.hpp"
#include "examples/quadratic_cost.hpp"
#include "examples/triple_integrator.hpp"
#include "examples/unicycle.hpp"

namespace {

struct Outputs {
  double* X;      // [(N+1) * n]
  double* U;      // [N * m]
  double* scalars;  // cost, max violation, max penalty, initial cost
  int* counters;    // status, iterations_inner, iterations_outer, iterations_total
  // optional (nullptr = not wanted)
  double* K = nullptr;     // [N][m * n] column-major, the gains the solve left behind
  double* d = nullptr;     // [N][m]
  double* history = nullptr;  // [history_cap][8]: cost, alpha, improvement_ratio, gradient, cost_decrease,
                              //                  regularization, violations, max_penalty (SolverStats vectors)
  int history_cap = 0;
  int* history_rows = nullptr;
};

template <class Solver>
void ExportGains(Solver& ilqr, int N, const Outputs& out) {
  if (out.K == nullptr && out.d == nullptr) return;
  for (int k = 0; k < N; ++k) {
    auto& K = ilqr.GetKnotPointFunction(k).GetFeedbackGain();
    auto& d = ilqr.GetKnotPointFunction(k).GetFeedforwardGain();
    const int mm = static_cast<int>(K.rows()), nn = static_cast<int>(K.cols());
    if (out.K != nullptr)
      for (int j = 0; j < nn; ++j)
        for (int i = 0; i < mm; ++i) out.K[(k * nn + j) * mm + i] = K(i, j);
    if (out.d != nullptr)
      for (int i = 0; i < mm; ++i) out.d[k * mm + i] = d(i);
  }
}
void ExportHistory(const altro::SolverStats& stats, const Outputs& out) {
  if (out.history == nullptr) return;
  const std::vector<double>* cols[8] = {&stats.cost, &stats.alpha, &stats.improvement_ratio, &stats.gradient,
                                        &stats.cost_decrease, &stats.regularization, &stats.violations, &stats.max_penalty};
  int rows = 0;
  for (const std::vector<double>* c : cols) rows = std::max(rows, static_cast<int>(c->size()));
  rows = std::min(rows, out.history_cap);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < 8; ++c)
      out.history[r * 8 + c] = r < static_cast<int>(cols[c]->size()) ? (*cols[c])[static_cast<std::size_t>(r)] : 0.0;
  if (out.history_rows != nullptr) *out.history_rows = rows;
}

template <int n, int m>
void Export(altro::Trajectory<n, m>& Z, int N, const Outputs& out) {
  const int nx = static_cast<int>(Z.State(0).size()), nu = static_cast<int>(Z.Control(0).size());  // n, m may be Dynamic
  for (int k = 0; k <= N; ++k)
    for (int i = 0; i < nx; ++i) out.X[k * nx + i] = Z.State(k)(i);
  for (int k = 0; k < N; ++k)
    for (int i = 0; i < nu; ++i) out.U[k * nu + i] = Z.Control(k)(i);
}

// options: [0] constraint_tolerance, [1] SetPenalty value, [2] initial_penalty, [3] max_iterations_total,
//          [4] max_iterations_inner, [5] max_iterations_outer; a negative entry keeps the reference's default
// further option overrides, set once for all following solves (altro_ref_set_extra_options; not thread-safe — the
// tests that use it solve one problem at a time): state_max, control_max, bp_reg_max, bp_reg_fail_threshold,
// line_search_max_iterations, cost_tolerance, gradient_tolerance, bp_reg_initial; negative = the reference's default
double g_extra[8] = {-1, -1, -1, -1, -1, -1, -1, -1};

void Apply(altro::SolverOptions& o, const double* options) {
  o.verbose = altro::LogLevel::kSilent;
  if (g_extra[0] >= 0) o.state_max = g_extra[0];
  if (g_extra[1] >= 0) o.control_max = g_extra[1];
  if (g_extra[2] >= 0) o.bp_reg_max = g_extra[2];
  if (g_extra[3] >= 0) o.bp_reg_fail_threshold = static_cast<int>(g_extra[3]);
  if (g_extra[4] >= 0) o.line_search_max_iterations = static_cast<int>(g_extra[4]);
  if (g_extra[5] >= 0) o.cost_tolerance = g_extra[5];
  if (g_extra[6] >= 0) o.gradient_tolerance = g_extra[6];
  if (g_extra[7] >= 0) o.bp_reg_initial = g_extra[7];
  if (options == nullptr) return;
  if (options[0] >= 0) o.constraint_tolerance = options[0];
  if (options[2] >= 0) o.initial_penalty = options[2];
  if (options[3] >= 0) o.max_iterations_total = static_cast<int>(options[3]);
  if (options[4] >= 0) o.max_iterations_inner = static_cast<int>(options[4]);
  if (options[5] >= 0) o.max_iterations_outer = static_cast<int>(options[5]);
}

// warm_solves: that many further Solve() calls on the same solver, each starting from the previous solution with
// reset_duals = false and initial_penalty = 0 (the MPC-style re-solve of docs/Overview.dox:49-54 there)
template <int n, int m>
void SolveConstrained(const altro::problem::Problem& prob, std::shared_ptr<altro::Trajectory<n, m>> Z, const double* options,
                      const Outputs& out, int warm_solves = 0) {
  altro::augmented_lagrangian::AugmentedLagrangianiLQR<n, m> solver(prob);
  solver.SetTrajectory(Z);
  Apply(solver.GetOptions(), options);
  if (options != nullptr && options[1] >= 0) solver.SetPenalty(options[1]);
  solver.Solve();
  for (int again = 0; again < warm_solves; ++again) {
    solver.GetOptions().reset_duals = false;
    solver.GetOptions().initial_penalty = 0.0;
    solver.Solve();
  }
  Export<n, m>(*Z, prob.NumSegments(), out);
  out.scalars[1] = solver.GetMaxViolation();  // of the constraint values the solve left behind: before Cost() refreshes them
  out.scalars[2] = solver.GetMaxPenalty();
  out.scalars[0] = solver.GetiLQRSolver().Cost();
  out.scalars[3] = solver.GetStats().initial_cost;
  out.counters[0] = static_cast<int>(solver.GetStatus());
  out.counters[1] = solver.GetStats().iterations_inner;
  out.counters[2] = solver.GetStats().iterations_outer;
  out.counters[3] = solver.GetStats().iterations_total;
  ExportGains(solver.GetiLQRSolver(), prob.NumSegments(), out);
  ExportHistory(solver.GetStats(), out);
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
  ExportGains(solver, prob.NumSegments(), out);
  ExportHistory(solver.GetStats(), out);
}

}  // namespace

extern "C" {

void altro_ref_set_extra_options(const double* extra) {
  for (int i = 0; i < 8; ++i) g_extra[i] = extra != nullptr ? extra[i] : -1.0;
}

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

// ---------------------------------------------------------------------------------------------------------------
// A problem assembled call by call, with the builder calls of altro_cpp_b200/problems.py ProblemSpec (the same
// calls build the oracle's and the device's problem): the reference's Problem filled with the reference's own
// QuadraticCost / GoalConstraint / ControlBound / CircleConstraint and DiscretizedModel<Unicycle | TripleIntegrator>.
// Two models of this repo's BASELINE configs do not exist in the reference and are supplied here as user functors
// against its ABCs, exactly as a user of the reference would: the cart-pole (C4) and a discrete linear system (C5).
// ---------------------------------------------------------------------------------------------------------------
namespace {

using altro::MatrixXd;
using altro::VectorXd;
using altro::VectorXdRef;

// frictionless cart-pole, state (x, theta, xd, thetad), theta = 0 hanging down; params mc, mp, l, g
// (the closed form of oracle/altro_oracle.hpp ModelEvaluate / ModelJacobian, same order of operations)
class CartpoleFunctor : public altro::problem::ContinuousDynamics {
 public:
  static constexpr int NStates = 4;
  static constexpr int NControls = 1;
  explicit CartpoleFunctor(const double* p) : mc_(p[0]), mp_(p[1]), l_(p[2]), g_(p[3]) {}
  int StateDimension() const override { return 4; }
  int ControlDimension() const override { return 1; }
  bool HasHessian() const override { return false; }
  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, const float, Eigen::Ref<VectorXd> xdot) override {
    const double th = x(1), xd = x(2), thd = x(3);
    const double s = std::sin(th), c = std::cos(th);
    const double den = mc_ + mp_ * s * s;
    const double F = u(0);
    const double xdd = (F + mp_ * s * (l_ * thd * thd + g_ * c)) / den;
    const double thdd = (-F * c - mp_ * l_ * thd * thd * c * s - (mc_ + mp_) * g_ * s) / (l_ * den);
    xdot(0) = xd;
    xdot(1) = thd;
    xdot(2) = xdd;
    xdot(3) = thdd;
  }
  void Jacobian(const VectorXdRef& x, const VectorXdRef& u, const float, Eigen::Ref<MatrixXd> jac) override {
    const double th = x(1), thd = x(3);
    const double s = std::sin(th), c = std::cos(th);
    const double den = mc_ + mp_ * s * s;
    const double F = u(0);
    const double numx = F + mp_ * s * (l_ * thd * thd + g_ * c);
    const double numt = -F * c - mp_ * l_ * thd * thd * c * s - (mc_ + mp_) * g_ * s;
    const double dden = 2.0 * mp_ * s * c;
    const double dnumx = mp_ * c * (l_ * thd * thd + g_ * c) - mp_ * s * g_ * s;
    const double dnumt = F * s - mp_ * l_ * thd * thd * (c * c - s * s) - (mc_ + mp_) * g_ * c;
    jac.setZero();
    jac(0, 2) = 1.0;
    jac(1, 3) = 1.0;
    jac(2, 1) = (dnumx * den - numx * dden) / (den * den);
    jac(2, 3) = (2.0 * mp_ * s * l_ * thd) / den;
    jac(2, 4) = 1.0 / den;
    jac(3, 1) = (dnumt * den - numt * dden) / (l_ * den * den);
    jac(3, 3) = (-2.0 * mp_ * l_ * thd * c * s) / (l_ * den);
    jac(3, 4) = -c / (l_ * den);
  }
  void Hessian(const VectorXdRef&, const VectorXdRef&, const float, const VectorXdRef&, Eigen::Ref<MatrixXd> hess) override {
    hess.setZero();
  }

 private:
  double mc_, mp_, l_, g_;
};

// x+ = A x + B u; params = A (n x n, column-major) then B (n x m)
class LinearFunctor : public altro::problem::DiscreteDynamics {
 public:
  LinearFunctor(int n, int m, const double* p) : n_(n), m_(m), ab_(p, p + n * (n + m)) {}
  int StateDimension() const override { return n_; }
  int ControlDimension() const override { return m_; }
  bool HasHessian() const override { return false; }
  using altro::problem::DiscreteDynamics::Evaluate;
  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, const float, const float, Eigen::Ref<VectorXd> xnext) override {
    const double* A = ab_.data();
    const double* B = A + n_ * n_;
    for (int i = 0; i < n_; ++i) {
      double acc = 0.0;
      for (int j = 0; j < n_; ++j) acc += A[i + j * n_] * x(j);
      double accb = 0.0;
      for (int j = 0; j < m_; ++j) accb += B[i + j * n_] * u(j);
      xnext(i) = acc + accb;
    }
  }
  void Jacobian(const VectorXdRef&, const VectorXdRef&, const float, const float, Eigen::Ref<MatrixXd> jac) override {
    for (int j = 0; j < n_ + m_; ++j)
      for (int i = 0; i < n_; ++i) jac(i, j) = ab_[i + j * n_];
  }
  void Hessian(const VectorXdRef&, const VectorXdRef&, const float, const float, const VectorXdRef&,
               Eigen::Ref<MatrixXd> hess) override {
    hess.setZero();
  }

 private:
  int n_, m_;
  std::vector<double> ab_;
};

struct Assembled {
  int n, m, N;
  altro::problem::Problem prob;
  std::vector<float> t, h;
  VectorXd x0;
  Assembled(int n_, int m_, int N_) : n(n_), m(m_), N(N_), prob(N_), t(N_ + 1, 0.0F), h(N_ + 1, 0.0F), x0(VectorXd::Zero(n_)) {}
};

MatrixXd ColMajor(const double* p, int rows, int cols) {
  MatrixXd M = MatrixXd::Zero(rows, cols);
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i) M(i, j) = p[i + j * rows];
  return M;
}
VectorXd Vec(const double* p, int len) {
  VectorXd v = VectorXd::Zero(len);
  for (int i = 0; i < len; ++i) v(i) = p[i];
  return v;
}

template <int n, int m>
void SolveAssembled(Assembled& a, bool constrained, const double* U0, const double* options, const Outputs& out,
                    int warm_solves = 0) {
  auto Z = std::make_shared<altro::Trajectory<n, m>>(a.n, a.m, a.N);
  for (int k = 0; k <= a.N; ++k) {
    Z->SetStep(k, a.h[k]);
    Z->SetTime(k, a.t[k]);
  }
  for (int k = 0; k < a.N; ++k)
    for (int i = 0; i < a.m; ++i) Z->Control(k)(i) = U0 != nullptr ? U0[k * a.m + i] : 0.0;
  a.prob.SetInitialState(a.x0);
  if (constrained) SolveConstrained<n, m>(a.prob, Z, options, out, warm_solves);
  else SolveUnconstrained<n, m>(a.prob, Z, options, out);
}

}  // namespace

extern "C" {

int altro_refb_problem_create(int n, int m, int N, void** handle) {
  *handle = new Assembled(n, m, N);
  return 0;
}
int altro_refb_problem_destroy(void* handle) {
  delete static_cast<Assembled*>(handle);
  return 0;
}
// kinds as in include/altro_b200.h: 0 unicycle, 1 triple integrator (dof = m), 2 cart-pole, 3 discrete linear
int altro_refb_problem_set_model(void* handle, int kind, const double* params, int nparams) {
  Assembled& a = *static_cast<Assembled*>(handle);
  std::shared_ptr<altro::problem::DiscreteDynamics> model;
  if (kind == 0) {
    model = std::make_shared<altro::problem::DiscretizedModel<altro::examples::Unicycle>>(altro::examples::Unicycle());
  } else if (kind == 1) {
    using Rk4 = altro::problem::RungeKutta4<Eigen::Dynamic, Eigen::Dynamic>;
    model = std::make_shared<altro::problem::DiscretizedModel<altro::examples::TripleIntegrator, Rk4>>(
        altro::examples::TripleIntegrator(a.m));
  } else if (kind == 2 && nparams == 4) {
    model = std::make_shared<altro::problem::DiscretizedModel<CartpoleFunctor>>(CartpoleFunctor(params));
  } else if (kind == 3 && nparams == a.n * (a.n + a.m)) {
    model = std::make_shared<LinearFunctor>(a.n, a.m, params);
  } else {
    return -1;
  }
  for (int k = 0; k < a.N; ++k) a.prob.SetDynamics(model, k);  // the models keep no per-call state
  return 0;
}
int altro_refb_problem_set_uniform_step(void* handle, float hs) {
  Assembled& a = *static_cast<Assembled*>(handle);
  altro::Trajectory<Eigen::Dynamic, Eigen::Dynamic> Z(a.n, a.m, a.N);
  Z.SetUniformStep(hs);  // the reference's own rounding of t_k and of the terminal knot
  for (int k = 0; k <= a.N; ++k) {
    a.h[k] = Z.GetStep(k);
    a.t[k] = static_cast<float>(Z.GetTime(k));
  }
  return 0;
}
int altro_refb_problem_set_steps(void* handle, const float* t, const float* h) {
  Assembled& a = *static_cast<Assembled*>(handle);
  for (int k = 0; k <= a.N; ++k) {
    a.t[k] = t[k];
    a.h[k] = h[k];
  }
  return 0;
}
// knots k0 .. k1-1; matrices column-major.  A range that ends after knot N holds the terminal cost.
int altro_refb_problem_set_cost(void* handle, int k0, int k1, const double* Q, const double* R, const double* H,
                                const double* q, const double* r, double c) {
  Assembled& a = *static_cast<Assembled*>(handle);
  for (int k = k0; k < k1; ++k) {
    const bool terminal = k == a.N;
    a.prob.SetCostFunction(std::make_shared<altro::examples::QuadraticCost>(ColMajor(Q, a.n, a.n), ColMajor(R, a.m, a.m),
                                                                          ColMajor(H, a.n, a.m), Vec(q, a.n),
                                                                          Vec(r, a.m), c, terminal),
                           k);
  }
  return 0;
}
int altro_refb_problem_add_goal(void* handle, int k, const double* xf) {
  Assembled& a = *static_cast<Assembled*>(handle);
  altro::constraints::ConstraintPtr<altro::constraints::Equality> goal =
      std::make_shared<altro::examples::GoalConstraint>(Vec(xf, a.n));
  a.prob.SetConstraint(goal, k);
  return 0;
}
int altro_refb_problem_add_control_bound(void* handle, int k, const double* lb, const double* ub) {
  Assembled& a = *static_cast<Assembled*>(handle);
  altro::constraints::ConstraintPtr<altro::constraints::Inequality> bnd = std::make_shared<altro::examples::ControlBound>(
      std::vector<double>(lb, lb + a.m), std::vector<double>(ub, ub + a.m));
  a.prob.SetConstraint(bnd, k);
  return 0;
}
int altro_refb_problem_add_circles(void* handle, int k, int count, const double* cx, const double* cy, const double* cr,
                                   int xi, int yi) {
  Assembled& a = *static_cast<Assembled*>(handle);
  if (xi != 0 || yi != 1) return -1;  // the reference's CircleConstraint reads the position from states 0 and 1
  std::shared_ptr<altro::examples::CircleConstraint> circles = std::make_shared<altro::examples::CircleConstraint>();
  for (int i = 0; i < count; ++i) circles->AddObstacle(cx[i], cy[i], cr[i]);
  altro::constraints::ConstraintPtr<altro::constraints::Inequality> con = circles;
  a.prob.SetConstraint(con, k);
  return 0;
}
int altro_refb_problem_set_initial_state(void* handle, const double* x0) {
  Assembled& a = *static_cast<Assembled*>(handle);
  a.x0 = Vec(x0, a.n);
  return 0;
}
// one solve of the assembled problem from initial state x0 (nullptr: the problem's) and controls U0 [N][m];
// K, d, history, history_rows may be nullptr
int altro_refb_solve_ex(void* handle, int constrained, const double* x0, const double* U0, const double* options, double* X,
                        double* U, double* scalars, int* counters, double* K, double* d, double* history, int history_cap,
                        int* history_rows) {
  Assembled& a = *static_cast<Assembled*>(handle);
  if (x0 != nullptr) a.x0 = Vec(x0, a.n);
  Outputs out{X, U, scalars, counters};
  out.K = K;
  out.d = d;
  out.history = history;
  out.history_cap = history_cap;
  out.history_rows = history_rows;
  const bool al = constrained != 0;
  if (a.n == 3 && a.m == 2) SolveAssembled<3, 2>(a, al, U0, options, out);
  else if (a.n == 6 && a.m == 2) SolveAssembled<6, 2>(a, al, U0, options, out);
  else if (a.n == 4 && a.m == 1) SolveAssembled<4, 1>(a, al, U0, options, out);
  else SolveAssembled<Eigen::Dynamic, Eigen::Dynamic>(a, al, U0, options, out);
  return a.N;
}
// an AL solve followed by `warm_solves` warm-started re-solves on the same solver; outputs are those of the last one
int altro_refb_solve_warm(void* handle, const double* x0, const double* U0, const double* options, int warm_solves,
                          double* X, double* U, double* scalars, int* counters) {
  Assembled& a = *static_cast<Assembled*>(handle);
  if (x0 != nullptr) a.x0 = Vec(x0, a.n);
  const Outputs out{X, U, scalars, counters};
  if (a.n == 3 && a.m == 2) SolveAssembled<3, 2>(a, true, U0, options, out, warm_solves);
  else if (a.n == 6 && a.m == 2) SolveAssembled<6, 2>(a, true, U0, options, out, warm_solves);
  else SolveAssembled<Eigen::Dynamic, Eigen::Dynamic>(a, true, U0, options, out, warm_solves);
  return a.N;
}
int altro_refb_solve(void* handle, int constrained, const double* x0, const double* U0, const double* options, double* X,
                     double* U, double* scalars, int* counters) {
  return altro_refb_solve_ex(handle, constrained, x0, U0, options, X, U, scalars, counters, nullptr, nullptr, nullptr, 0,
                             nullptr);
}

}  // extern "C"
