// include/altro_b200.hpp — header-only C++ host API over the C ABI (altro_b200.h).
//
// The reference's public API is Eigen-typed (altro/eigentypes.hpp:8-27) and Eigen is not part
// of this image, so this wrapper speaks plain column-major `double` arrays; the class and
// method names follow the reference (`Problem`, `SetCostFunction`, `SetConstraint`,
// `SetInitialState`, `Solve`, `GetStatus`, ...) so that host code written against
// altro::problem::Problem / AugmentedLagrangianiLQR<n,m> maps one to one.  INTEGRATION.md shows
// the Eigen-typed shim a reference maintainer adds on top of it.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "altro_b200.h"

namespace altro_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(int rc, const char* what) {
  if (rc != 0) throw Error(rc, std::string(what) + ": " + altro_b200_last_error());
}

// altro::problem::Problem (altro/problem/problem.hpp:65) for the device-capable functors.
class Problem {
 public:
  Problem(int n, int m, int N) : n_(n), m_(m), N_(N) { check(altro_b200_problem_create(n, m, N, &p_), "Problem"); }
  ~Problem() { altro_b200_problem_destroy(p_); }
  Problem(const Problem&) = delete;
  Problem& operator=(const Problem&) = delete;

  int NumSegments() const { return N_; }
  // SetDynamics(std::make_shared<DiscretizedModel<Model>>(model), k) for every k
  void SetDynamics(altro_b200_model model, const std::vector<double>& params = {}) {
    check(altro_b200_problem_set_model(p_, model, params.data(), static_cast<int>(params.size())), "SetDynamics");
  }
  void SetUniformStep(float h) { check(altro_b200_problem_set_uniform_step(p_, h), "SetUniformStep"); }
  // SetCostFunction(std::make_shared<QuadraticCost>(Q, R, H, q, r, c), k) for k0 <= k < k1
  void SetCostFunction(int k0, int k1, const double* Q, const double* R, const double* H, const double* q,
                       const double* r, double c) {
    check(altro_b200_problem_set_cost(p_, k0, k1, Q, R, H, q, r, c), "SetCostFunction");
  }
  // SetConstraint(std::make_shared<GoalConstraint>(xf), k)
  void SetGoalConstraint(int k, const double* xf) { check(altro_b200_problem_add_goal(p_, k, xf), "SetConstraint"); }
  // SetConstraint(std::make_shared<ControlBound>(lb, ub), k)
  void SetControlBound(int k, const double* lb, const double* ub) {
    check(altro_b200_problem_add_control_bound(p_, k, lb, ub), "SetConstraint");
  }
  // SetConstraint(std::make_shared<CircleConstraint>(obstacles), k)
  void SetCircleConstraint(int k, const std::vector<double>& cx, const std::vector<double>& cy,
                           const std::vector<double>& r, int xi = 0, int yi = 1) {
    check(altro_b200_problem_add_circles(p_, k, static_cast<int>(cx.size()), cx.data(), cy.data(), r.data(), xi, yi),
          "SetConstraint");
  }
  void SetInitialState(const double* x0) { check(altro_b200_problem_set_initial_state(p_, x0), "SetInitialState"); }
  const altro_b200_problem* handle() const { return p_; }
  int n() const { return n_; }
  int m() const { return m_; }

 private:
  altro_b200_problem* p_ = nullptr;
  int n_, m_, N_;
};

// B independent AugmentedLagrangianiLQR<n,m> solves (altro/augmented_lagrangian/al_solver.hpp:28).
class BatchedAugmentedLagrangianiLQR {
 public:
  BatchedAugmentedLagrangianiLQR(const Problem& prob, int batch, int device = 0, bool use_constraints = true)
      : B_(batch), n_(prob.n()), m_(prob.m()), N_(prob.NumSegments()) {
    check(altro_b200_solver_create(prob.handle(), batch, use_constraints ? 1 : 0, device, &s_), "solver");
    altro_b200_default_options(&opts_);
  }
  ~BatchedAugmentedLagrangianiLQR() { altro_b200_solver_destroy(s_); }
  BatchedAugmentedLagrangianiLQR(const BatchedAugmentedLagrangianiLQR&) = delete;
  BatchedAugmentedLagrangianiLQR& operator=(const BatchedAugmentedLagrangianiLQR&) = delete;

  altro_b200_options& GetOptions() { return opts_; }
  // SetTrajectory(initial guess) + per-instance initial states: x0 [B][n], U0 [B][N][m] or the
  // nominal control of InitialTrajectory()
  void SetTrajectory(const double* x0, const double* U0, const double* u_nominal) {
    check(altro_b200_solver_set_inputs_host(s_, x0, U0, u_nominal, nullptr), "SetTrajectory");
  }
  void SetPenalty(double rho) { check(altro_b200_solver_set_penalty(s_, rho, nullptr), "SetPenalty"); }
  void Solve() {
    check(altro_b200_solver_set_options(s_, &opts_), "GetOptions");
    check(altro_b200_solve_al(s_, nullptr), "Solve");
  }
  // per-instance GetStatus(), GetStats().iterations_*, Cost(), GetMaxViolation()
  void GetResults(std::vector<double>* cost, std::vector<double>* viol, std::vector<int32_t>* status,
                  std::vector<int32_t>* iters) {
    if (cost) cost->resize(B_);
    if (viol) viol->resize(B_);
    if (status) status->resize(B_);
    if (iters) iters->resize(3 * static_cast<size_t>(B_));
    check(altro_b200_get_results_host(s_, cost ? cost->data() : nullptr, viol ? viol->data() : nullptr,
                                      status ? status->data() : nullptr, iters ? iters->data() : nullptr, nullptr),
          "GetResults");
  }
  void GetTrajectory(std::vector<double>* X, std::vector<double>* U) {
    if (X) X->resize(static_cast<size_t>(B_) * (N_ + 1) * n_);
    if (U) U->resize(static_cast<size_t>(B_) * N_ * m_);
    check(altro_b200_get_trajectory_host(s_, X ? X->data() : nullptr, U ? U->data() : nullptr, nullptr),
          "GetTrajectory");
  }
  altro_b200_solver* handle() { return s_; }

 private:
  altro_b200_solver* s_ = nullptr;
  altro_b200_options opts_;
  int B_, n_, m_, N_;
};

}  // namespace altro_b200
