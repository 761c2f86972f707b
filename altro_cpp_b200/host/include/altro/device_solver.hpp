// altro/device_solver.hpp — the one place where the host mirror touches the GPU library.
//
// `DeviceSolver` packs an altro::problem::Problem through the C ABI (include/altro_b200.h) and
// forwards every public method of iLQR<n,m> / AugmentedLagrangianiLQR<n,m> to the matching
// altro_b200_* entry point.  It holds no numerical code of the solve: if the library reports an error
// (no CUDA device, unsupported (n, m, model), a functor the device registry does not recognise —
// altro/device_registry.hpp) the call throws altro::DeviceError — there is no host fallback.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "altro/common/solver_options.hpp"
#include "altro/common/solver_stats.hpp"
#include "altro/common/trajectory.hpp"
#include "altro/device_registry.hpp"
#include "altro/problem/problem.hpp"
#include "altro_b200.h"

namespace altro {

struct DeviceError : std::runtime_error {
  int code;
  DeviceError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

namespace detail {

inline void Check(int rc, const char* what) {
  if (rc != ALTRO_B200_OK) throw DeviceError(rc, std::string(what) + ": " + altro_b200_last_error());
}

class DeviceSolver {
 public:
  // devices: more than one entry shards the batch over those GPUs (altro_b200_multi_*): whole solves only
  DeviceSolver(const problem::Problem& prob, int n, int m, bool use_constraints, int batch = 1, int device = 0,
               std::vector<int> devices = {})
      : prob_(prob.ConstrainedProblem()), n_(n), m_(m), N_(prob.NumSegments()), B_(batch), device_(device),
        use_constraints_(use_constraints), devices_(std::move(devices)) {
    if (!prob_.IsFullyDefined()) throw std::invalid_argument("Expected problem to be fully defined.");
    if (batch < 1) throw std::invalid_argument("batch must be positive");
    X_.assign(static_cast<size_t>(B_) * (N_ + 1) * n_, 0.0);
    U_.assign(static_cast<size_t>(B_) * N_ * m_, 0.0);
    x0_.assign(static_cast<size_t>(B_) * n_, 0.0);
  }
  ~DeviceSolver() {
    if (solver_) altro_b200_solver_destroy(solver_);
    if (multi_) altro_b200_multi_destroy(multi_);
  }
  bool Sharded() const { return devices_.size() > 1; }
  DeviceSolver(const DeviceSolver&) = delete;
  DeviceSolver& operator=(const DeviceSolver&) = delete;

  int n() const { return n_; }
  int m() const { return m_; }
  int NumSegments() const { return N_; }
  int Batch() const { return B_; }
  bool UsesConstraints() const { return use_constraints_; }
  SolverOptions& GetOptions() { return stats_.GetOptions(); }
  SolverStats& GetStats() { return stats_; }

  // ---- inputs ------------------------------------------------------------------------------
  // Times and steps live in the trajectory (knotpoint.hpp:180 there), the device problem needs them at
  // creation: the device solver is (re)built when a trajectory with another time grid arrives.
  void SetTimeGrid(const std::vector<float>& t, const std::vector<float>& h) {
    if (have_step_ && t == t_ && h == h_all_) return;
    if (solver_) {
      altro_b200_solver_destroy(solver_);
      solver_ = nullptr;
    }
    if (multi_) {
      altro_b200_multi_destroy(multi_);
      multi_ = nullptr;
    }
    t_ = t;
    h_all_ = h;
    have_step_ = true;
  }
  void SetStep(float h) {  // uniform grid, t_k = float(k) * h like Trajectory::SetUniformStep
    std::vector<float> t(static_cast<size_t>(N_) + 1), hs(static_cast<size_t>(N_) + 1, h);
    for (int k = 0; k < N_; ++k) t[k] = static_cast<float>(k) * h;
    t[N_] = static_cast<float>(h) * N_;
    hs[N_] = 0.0F;
    SetTimeGrid(t, hs);
  }
  void SetPenalty(double rho) {
    penalty_ = rho;
    have_penalty_ = true;
    if (solver_) Check(altro_b200_solver_set_penalty(solver_, rho, nullptr), "SetPenalty");
  }
  void SetPenaltyScaling(double phi) { penalty_scaling_ = phi; }
  // per-instance initial states (batched solvers); the single solvers read the problem's shared
  // initial state at every upload like the reference does (ilqr.hpp:455)
  void SetInitialStates(const std::vector<VectorXd>& x0) {
    if (static_cast<int>(x0.size()) != B_) throw std::invalid_argument("one initial state per instance");
    for (int b = 0; b < B_; ++b)
      for (int i = 0; i < n_; ++i) x0_[static_cast<size_t>(b) * n_ + i] = x0[b](i);
    explicit_x0_ = true;
  }

  // Z -> staging -> device: the states and controls of `Z` become those of every instance.
  template <class Traj>
  void Upload(const Traj& Z) {
    {  // per-knot times and steps, exactly as the trajectory stores them
      std::vector<float> t(static_cast<size_t>(N_) + 1), h(static_cast<size_t>(N_) + 1);
      for (int k = 0; k <= N_; ++k) {
        t[k] = static_cast<float>(Z.GetTime(k));
        h[k] = Z.GetStep(k);
      }
      SetTimeGrid(t, h);
    }
    Ensure();
    if (!explicit_x0_) {
      const VectorXd& x0 = prob_.GetInitialState();
      for (int b = 0; b < B_; ++b)
        for (int i = 0; i < n_; ++i) x0_[static_cast<size_t>(b) * n_ + i] = x0(i);
    }
    for (int b = 0; b < B_; ++b) {
      double* Xb = X_.data() + static_cast<size_t>(b) * (N_ + 1) * n_;
      double* Ub = U_.data() + static_cast<size_t>(b) * N_ * m_;
      for (int k = 0; k <= N_; ++k)
        for (int i = 0; i < n_; ++i) Xb[k * n_ + i] = Z.State(k)(i);
      for (int k = 0; k < N_; ++k)
        for (int j = 0; j < m_; ++j) Ub[k * m_ + j] = Z.Control(k)(j);
    }
    if (Sharded()) return;  // the sharded solve takes the host arrays directly
    Check(altro_b200_solver_set_inputs_host(solver_, x0_.data(), U_.data(), nullptr, nullptr), "SetTrajectory");
    Check(altro_b200_solver_set_states_host(solver_, X_.data(), nullptr), "SetTrajectory");
  }
  // device -> staging -> Z (instance b)
  template <class Traj>
  void Download(Traj* Z, int b = 0) {
    Fetch();
    CopyOut(Z, b);
  }
  void Fetch() {
    if (Sharded()) return;  // results arrived with the solve
    Need();
    Check(altro_b200_get_trajectory_host(solver_, X_.data(), U_.data(), nullptr), "GetTrajectory");
  }
  template <class Traj>
  void CopyOut(Traj* Z, int b) const {
    const double* Xb = X_.data() + static_cast<size_t>(b) * (N_ + 1) * n_;
    const double* Ub = U_.data() + static_cast<size_t>(b) * N_ * m_;
    for (int k = 0; k <= N_; ++k)
      for (int i = 0; i < n_; ++i) Z->State(k)(i) = Xb[k * n_ + i];
    for (int k = 0; k < N_; ++k)
      for (int j = 0; j < m_; ++j) Z->Control(k)(j) = Ub[k * m_ + j];
  }

  // ---- phases --------------------------------------------------------------------------------
  enum Phase {
    kRollout,
    kCost,
    kUpdateExpansions,
    kBackwardPass,
    kForwardPass,
    kUpdateConvergenceStatistics,
    kUpdateDuals,
    kUpdatePenalties,
    kAlInit,
    kSolveSetup,
    kSolveILQR,
    kSolveAL
  };
  // Duals edited on the host (GetALCost(k)->Get...Constraints()[i]->GetDuals() = ...) reach the device right
  // before the next device phase: the view hands out a buffer and registers it here.
  void PushDualsBeforeNextRun(int k, int row0, std::shared_ptr<VectorXd> values) {
    pending_duals_.push_back({k, row0, std::move(values)});
  }
  // An outer (dual-update) iteration the caller performed step by step (AugmentedLagrangianiLQR::
  // UpdateConvergenceStatistics): the device counts only the outer iterations of whole solves.
  void CountOuterIterationByHand() { ++outer_by_hand_; }
  // all multipliers of every instance to zero (AugmentedLagrangianiLQR::ResetDualVariables)
  void ZeroDuals() {
    Need();
    pending_duals_.clear();
    for (int k = 0; k <= N_; ++k) {
      const int rows = prob_.NumConstraints(k);
      if (rows == 0) continue;
      const std::vector<double> zeros(static_cast<size_t>(rows), 0.0);
      Check(altro_b200_solver_set_duals_host(solver_, k, zeros.data(), rows, nullptr), "ResetDualVariables");
    }
  }
  void FlushPendingDuals() {
    if (pending_duals_.empty() || Sharded() || !solver_) return;
    std::vector<PendingDuals> todo;
    todo.swap(pending_duals_);
    for (const PendingDuals& e : todo) {
      std::vector<double> all = Duals(e.k, 0);
      for (int i = 0; i < e.values->size(); ++i) all.at(static_cast<size_t>(e.row0 + i)) = (*e.values)(i);
      Check(altro_b200_solver_set_duals_host(solver_, e.k, all.data(), static_cast<int>(all.size()), nullptr), "GetDuals");
    }
  }
  void Run(Phase ph) {
    FlushPendingDuals();
    if (Sharded()) {
      if (ph != kSolveAL) throw DeviceError(ALTRO_B200_ERR_UNSUPPORTED, "a batch sharded over several GPUs supports whole solves only");
      if (!multi_) throw DeviceError(ALTRO_B200_ERR_STATE, "SetTrajectory must be called before this method");
      altro_b200_options d = MakeOptions();
      Check(altro_b200_multi_set_options(multi_, &d), "GetOptions");
      res_.cost.resize(B_);
      res_.viol.resize(B_);
      res_.status.resize(B_);
      res_.iters.resize(static_cast<size_t>(B_) * 3);
      std::vector<double> U0 = U_;  // U_ is overwritten with the solution
      Check(altro_b200_multi_solve_al_host(multi_, x0_.data(), U0.data(), nullptr, X_.data(), U_.data(), res_.cost.data(),
                                           res_.viol.data(), res_.status.data(), res_.iters.data()),
            "Solve");
      return;
    }
    Need();
    PushOptions();
    if (ph == kSolveAL || ph == kAlInit) outer_by_hand_ = 0;  // the device counts the outer iterations of a whole solve
    if (ph == kSolveAL || ph == kSolveILQR) take_initial_cost_ = true;  // a whole solve measures it itself
    int rc = ALTRO_B200_ERR_ARG;
    switch (ph) {
      case kRollout: rc = altro_b200_rollout(solver_, nullptr); break;
      case kCost: rc = altro_b200_cost(solver_, nullptr); break;
      case kUpdateExpansions: rc = altro_b200_update_expansions(solver_, nullptr); break;
      case kBackwardPass: rc = altro_b200_backward_pass(solver_, nullptr); break;
      case kForwardPass: rc = altro_b200_forward_pass(solver_, nullptr); break;
      case kUpdateConvergenceStatistics: rc = altro_b200_update_convergence_statistics(solver_, nullptr); break;
      case kUpdateDuals: rc = altro_b200_update_duals(solver_, nullptr); break;
      case kUpdatePenalties: rc = altro_b200_update_penalties(solver_, nullptr); break;
      case kAlInit: rc = altro_b200_al_init(solver_, nullptr); break;
      case kSolveSetup: rc = altro_b200_solve_setup(solver_, nullptr); break;
      case kSolveILQR: rc = altro_b200_solve_ilqr(solver_, nullptr); break;
      case kSolveAL: rc = altro_b200_solve_al(solver_, nullptr); break;
    }
    Check(rc, "device phase");
  }

  // ---- results -------------------------------------------------------------------------------
  struct Results {
    std::vector<double> cost, viol;
    std::vector<int32_t> status, ilqr_status, iters;
    std::vector<double> initial_cost, reg;
  };
  const Results& Pull() {
    if (Sharded()) return res_;
    Need();
    res_.cost.resize(B_);
    res_.viol.resize(B_);
    res_.status.resize(B_);
    res_.ilqr_status.resize(B_);
    res_.iters.resize(static_cast<size_t>(B_) * 3);
    res_.initial_cost.resize(B_);
    res_.reg.resize(B_);
    Check(altro_b200_get_results_host(solver_, res_.cost.data(), res_.viol.data(), res_.status.data(),
                                      res_.iters.data(), nullptr),
          "GetStats");
    Check(altro_b200_get_ilqr_status_host(solver_, res_.ilqr_status.data(), nullptr), "GetStatus");
    Check(altro_b200_get_scalars_host(solver_, res_.reg.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                      nullptr, res_.initial_cost.data(), nullptr),
          "GetStats");
    stats_.iterations_inner = res_.iters[0];
    stats_.iterations_outer = res_.iters[1] + outer_by_hand_;
    stats_.iterations_total = res_.iters[2];
    if (take_initial_cost_) stats_.initial_cost = res_.initial_cost[0];  // else the caller's (step-wise iterations)
    take_initial_cost_ = false;
    return res_;
  }
  const Results& Last() const { return res_; }
  // K [N][m*n] column-major and d [N][m] of instance b
  void Gains(int b, std::vector<double>* K, std::vector<double>* d) {
    Need();
    std::vector<double> Kall(static_cast<size_t>(B_) * N_ * m_ * n_), dall(static_cast<size_t>(B_) * N_ * m_);
    Check(altro_b200_get_gains_host(solver_, Kall.data(), dall.data(), nullptr), "GetFeedbackGain");
    K->assign(Kall.begin() + static_cast<size_t>(b) * N_ * m_ * n_,
              Kall.begin() + static_cast<size_t>(b + 1) * N_ * m_ * n_);
    d->assign(dall.begin() + static_cast<size_t>(b) * N_ * m_, dall.begin() + static_cast<size_t>(b + 1) * N_ * m_);
  }
  void CostToGo(int k, int b, MatrixXd* P, VectorXd* p) {
    Need();
    std::vector<double> Pall(static_cast<size_t>(B_) * n_ * n_), pall(static_cast<size_t>(B_) * n_);
    Check(altro_b200_get_ctg_host(solver_, k, Pall.data(), pall.data(), nullptr), "GetCostToGo");
    *P = MatrixXd(n_, n_);
    *p = VectorXd(n_);
    for (int j = 0; j < n_; ++j)
      for (int i = 0; i < n_; ++i) (*P)(i, j) = Pall[static_cast<size_t>(b) * n_ * n_ + j * n_ + i];
    for (int i = 0; i < n_; ++i) (*p)(i) = pall[static_cast<size_t>(b) * n_ + i];
  }
  // costs_(k), k = 0 .. N, of instance b as the last Cost() / UpdateExpansions() left them
  std::vector<double> Costs(int b) {
    Need();
    std::vector<double> all(static_cast<size_t>(B_) * (N_ + 1));
    Check(altro_b200_get_costs_host(solver_, all.data(), nullptr), "GetCosts");
    return std::vector<double>(all.begin() + static_cast<size_t>(b) * (N_ + 1),
                               all.begin() + static_cast<size_t>(b + 1) * (N_ + 1));
  }
  // stats.initial_cost of a caller who runs the inner iterations one by one (ilqr.hpp:292 there)
  void SetInitialCost(double J) {
    Need();
    Check(altro_b200_solver_set_initial_cost(solver_, J, nullptr), "GetStats");
  }
  // c(x_k, u_k) of instance b at knot k, ALCost row order (equalities, then inequalities)
  std::vector<double> ConstraintValues(int k, int b) {
    Need();
    int pmax = 0, pk = 0;
    Check(altro_b200_get_duals_host(solver_, k, nullptr, &pmax, nullptr), "GetConstraintInfo");
    std::vector<double> all(static_cast<size_t>(B_) * (pmax > 0 ? pmax : 1));
    Check(altro_b200_get_constraint_values_host(solver_, k, pmax > 0 ? all.data() : nullptr, &pk, nullptr),
          "GetConstraintInfo");
    return std::vector<double>(all.begin() + static_cast<size_t>(b) * pmax,
                               all.begin() + static_cast<size_t>(b) * pmax + pk);
  }
  std::vector<double> Duals(int k, int b) {
    Need();
    int pmax = 0, pk = 0;
    Check(altro_b200_get_duals_host(solver_, k, nullptr, &pmax, nullptr), "GetDuals");
    Check(altro_b200_get_constraint_values_host(solver_, k, nullptr, &pk, nullptr), "GetDuals");
    std::vector<double> all(static_cast<size_t>(B_) * (pmax > 0 ? pmax : 1));
    if (pmax > 0) Check(altro_b200_get_duals_host(solver_, k, all.data(), &pmax, nullptr), "GetDuals");
    return std::vector<double>(all.begin() + static_cast<size_t>(b) * pmax,
                               all.begin() + static_cast<size_t>(b) * pmax + pk);
  }
  double MaxPenalty(int b = 0) {
    Need();
    std::vector<double> pen(B_);
    Check(altro_b200_get_scalars_host(solver_, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                      pen.data(), nullptr, nullptr),
          "GetMaxPenalty");
    return pen[b];
  }
  void Expansion(int k, int b, MatrixXd* A, MatrixXd* Bm, MatrixXd* lxx, MatrixXd* lxu, MatrixXd* luu, VectorXd* lx,
                 VectorXd* lu) {
    Need();
    const size_t B = B_;
    std::vector<double> a(B * n_ * n_), bb(B * n_ * m_), xx(B * n_ * n_), xu(B * n_ * m_), uu(B * m_ * m_), x(B * n_),
        u(B * m_);
    Check(altro_b200_get_expansion_host(solver_, k, a.data(), bb.data(), xx.data(), xu.data(), uu.data(), x.data(),
                                        u.data(), nullptr),
          "GetCostExpansion");
    auto mat = [&](const std::vector<double>& v, int r, int c) {
      MatrixXd M(r, c);
      for (int j = 0; j < c; ++j)
        for (int i = 0; i < r; ++i) M(i, j) = v[static_cast<size_t>(b) * r * c + j * r + i];
      return M;
    };
    auto vec = [&](const std::vector<double>& v, int r) {
      VectorXd V(r);
      for (int i = 0; i < r; ++i) V(i) = v[static_cast<size_t>(b) * r + i];
      return V;
    };
    *A = mat(a, n_, n_);
    *Bm = mat(bb, n_, m_);
    *lxx = mat(xx, n_, n_);
    *lxu = mat(xu, n_, m_);
    *luu = mat(uu, m_, m_);
    *lx = vec(x, n_);
    *lu = vec(u, m_);
  }
  // step-wise getters that answer "not available yet" instead of throwing
  bool Ready() const { return solver_ != nullptr; }
  bool TryExpansion(int k, int b, MatrixXd* A, MatrixXd* Bm, MatrixXd* lxx, MatrixXd* lxu, MatrixXd* luu, VectorXd* lx,
                    VectorXd* lu) {
    try {
      Expansion(k, b, A, Bm, lxx, lxu, luu, lx, lu);
      return true;
    } catch (const DeviceError& e) {
      if (e.code == ALTRO_B200_ERR_STATE) return false;
      throw;
    }
  }
  bool TryCostToGo(int k, int b, MatrixXd* P, VectorXd* p) {
    try {
      CostToGo(k, b, P, p);
      return true;
    } catch (const DeviceError& e) {
      if (e.code == ALTRO_B200_ERR_STATE) return false;
      throw;
    }
  }
  // step-wise use (ForwardPass / UpdateConvergenceStatistics called one by one): the scalars the reference logs
  struct StepScalars { double alpha, z, dJ, grad; };
  StepScalars Scalars(int b = 0) {
    Need();
    std::vector<double> a(B_), z(B_), dj(B_), g(B_);
    Check(altro_b200_get_scalars_host(solver_, nullptr, nullptr, nullptr, a.data(), z.data(), dj.data(), g.data(), nullptr,
                                      nullptr, nullptr),
          "GetStats");
    return {a[b], z[b], dj[b], g[b]};
  }
  // Per-iteration SolverStats vectors (cost, alpha, improvement_ratio, gradient, cost_decrease, regularization,
  // violations, max_penalty) of a single-problem solver, rebuilt from the rows the device recorded: one row per
  // inner iteration plus the open slot NewIteration() leaves behind (solver_stats.cpp:54-66), into which the
  // AL solver logs the final violation and penalty (al_solver.hpp:361-362).  `al`: called by the AL solver.
  void PullHistory(bool al = true) {
    if (!history_ || Sharded() || !solver_) return;
    const int cap = history_rows_;
    std::vector<double> rows(static_cast<size_t>(cap) * ALTRO_B200_HISTORY_COLS);
    int nrows = 0;
    Check(altro_b200_get_history_host(solver_, 0, rows.data(), cap, &nrows, nullptr), "GetStats");
    std::vector<double> col[ALTRO_B200_HISTORY_COLS];
    for (int c = 0; c < ALTRO_B200_HISTORY_COLS; ++c) {
      col[c].resize(nrows + 1);
      for (int r = 0; r < nrows; ++r) col[c][r] = rows[static_cast<size_t>(r) * ALTRO_B200_HISTORY_COLS + c];
      col[c][nrows] = nrows > 0 ? col[c][nrows - 1] : 0.0;  // NewIteration(): the new slot starts as a copy
    }
    if (al) {
      col[6][nrows] = res_.viol.empty() ? 0.0 : res_.viol[0];
      col[7][nrows] = has_constraints_ ? MaxPenalty(0) : 0.0;
    }
    stats_.SetHistory(nrows + 1, col[0].data(), col[1].data(), col[2].data(), col[3].data(), col[4].data(),
                      col[5].data(), col[6].data(), col[7].data());
  }
  const problem::Problem& GetProblem() const { return prob_; }
  int64_t KernelLaunches() const { return solver_ ? altro_b200_kernel_launches(solver_) : 0; }

 private:
  void Need() const {
    if (!solver_) throw DeviceError(ALTRO_B200_ERR_STATE, "SetTrajectory must be called before this method");
  }
  void PushOptions() {
    altro_b200_options d = MakeOptions();
    Check(altro_b200_solver_set_options(solver_, &d), "GetOptions");
  }
  altro_b200_options MakeOptions() {
    const SolverOptions& o = stats_.GetOptions();
    altro_b200_options d;
    altro_b200_default_options(&d);
    d.max_iterations_total = o.max_iterations_total;
    d.max_iterations_outer = o.max_iterations_outer;
    d.max_iterations_inner = o.max_iterations_inner;
    d.bp_reg_fail_threshold = o.bp_reg_fail_threshold;
    d.check_forwardpass_bounds = o.check_forwardpass_bounds ? 1 : 0;
    d.line_search_max_iterations = o.line_search_max_iterations;
    d.reset_duals = o.reset_duals ? 1 : 0;
    d.cost_tolerance = o.cost_tolerance;
    d.gradient_tolerance = o.gradient_tolerance;
    d.bp_reg_increase_factor = o.bp_reg_increase_factor;
    d.bp_reg_initial = o.bp_reg_initial;
    d.bp_reg_max = o.bp_reg_max;
    d.bp_reg_min = o.bp_reg_min;
    d.state_max = o.state_max;
    d.control_max = o.control_max;
    d.line_search_lower_bound = o.line_search_lower_bound;
    d.line_search_upper_bound = o.line_search_upper_bound;
    d.line_search_decrease_factor = o.line_search_decrease_factor;
    d.constraint_tolerance = o.constraint_tolerance;
    d.maximum_penalty = o.maximum_penalty;
    d.initial_penalty = o.initial_penalty;
    if (penalty_scaling_ > 0) d.penalty_scaling = penalty_scaling_;
    return d;
  }

  // Problem -> altro_b200_problem -> altro_b200_solver
  void Ensure() {
    if (solver_ || multi_) return;
    if (!have_step_) throw DeviceError(ALTRO_B200_ERR_STATE, "the trajectory carries no time step");
    altro_b200_problem* p = nullptr;
    Check(altro_b200_problem_create(n_, m_, N_, &p), "Problem");
    struct Guard {
      altro_b200_problem* p;
      ~Guard() { altro_b200_problem_destroy(p); }
    } guard{p};

    device::ModelDesc model;
    std::string why;
    for (int k = 0; k < N_; ++k) {
      device::ModelDesc mk;
      if (k > 0 && prob_.GetDynamics(k) == prob_.GetDynamics(0)) continue;
      if (!device::DescribeDynamics(*prob_.GetDynamics(k), &mk, &why))
        throw DeviceError(ALTRO_B200_ERR_UNSUPPORTED, "dynamics at knot " + std::to_string(k) + ": " + why +
                                                          " (host callbacks cannot run on the GPU; see INTEGRATION.md)");
      if (k == 0) model = mk;
      if (mk.model != model.model || mk.params != model.params)
        throw DeviceError(ALTRO_B200_ERR_UNSUPPORTED, "all knot points must share one dynamics model");
    }
    Check(altro_b200_problem_set_model(p, model.model, model.params.data(), static_cast<int>(model.params.size())),
          "SetDynamics");
    Check(altro_b200_problem_set_steps(p, t_.data(), h_all_.data()), "SetUniformStep");
    for (int k = 0; k <= N_; ++k) {
      device::CostDesc c;
      if (!device::DescribeCost(*prob_.GetCostFunction(k), n_, m_, &c, &why))
        throw DeviceError(ALTRO_B200_ERR_UNSUPPORTED, "cost function at knot " + std::to_string(k) + ": " + why);
      Check(altro_b200_problem_set_cost(p, k, k + 1, c.Q.data(), c.R.data(), c.H.data(), c.q.data(), c.r.data(), c.c),
            "SetCostFunction");
    }
    has_constraints_ = false;
    if (use_constraints_) {
      for (int k = 0; k <= N_; ++k) {
        for (const auto& con : prob_.GetEqualityConstraints()[k]) AddConstraint(p, k, *con);
        for (const auto& con : prob_.GetInequalityConstraints()[k]) AddConstraint(p, k, *con);
        has_constraints_ = has_constraints_ || !prob_.GetEqualityConstraints()[k].empty() ||
                           !prob_.GetInequalityConstraints()[k].empty();
      }
    }
    Check(altro_b200_problem_set_initial_state(p, prob_.GetInitialState().data()), "SetInitialState");
    if (Sharded()) {
      Check(altro_b200_multi_create(p, B_, use_constraints_ ? 1 : 0, devices_.data(), static_cast<int>(devices_.size()), &multi_),
            "solver");
      return;
    }
    Check(altro_b200_solver_create(p, B_, use_constraints_ ? 1 : 0, device_, &solver_), "solver");
    if (B_ == 1) {  // the single-problem API keeps the reference's per-iteration SolverStats vectors
      history_rows_ = std::max(1000, stats_.GetOptions().max_iterations_total);
      const int rc = altro_b200_solver_enable_history(solver_, 1, history_rows_);
      if (rc != ALTRO_B200_ERR_UNSUPPORTED) Check(rc, "GetStats");  // (not recorded on the large-state path)
      history_ = rc == 0;
    }
    if (have_penalty_) Check(altro_b200_solver_set_penalty(solver_, penalty_, nullptr), "SetPenalty");
  }
  template <class Con>
  void AddConstraint(altro_b200_problem* p, int k, Con& con) {
    device::ConstraintDesc d;
    std::string why;
    if (!device::DescribeConstraint(con, n_, m_, &d, &why))
      throw DeviceError(ALTRO_B200_ERR_UNSUPPORTED, "knot " + std::to_string(k) + ": " + why);
    switch (d.kind) {
      case device::ConstraintDesc::kGoal:
        Check(altro_b200_problem_add_goal(p, k, d.a.data()), "SetConstraint");
        break;
      case device::ConstraintDesc::kControlBound:
        Check(altro_b200_problem_add_control_bound(p, k, d.a.data(), d.b.data()), "SetConstraint");
        break;
      case device::ConstraintDesc::kCircle:
        Check(altro_b200_problem_add_circles_r2(p, k, static_cast<int>(d.a.size()), d.a.data(), d.b.data(), d.c.data(),
                                                d.xi, d.yi),
              "SetConstraint");
        break;
      default:
        throw DeviceError(ALTRO_B200_ERR_UNSUPPORTED, "unknown constraint descriptor");
    }
  }

  problem::Problem prob_;  // shares the functors and the initial-state pointer with the caller
  int n_, m_, N_, B_, device_;
  struct PendingDuals {
    int k, row0;
    std::shared_ptr<VectorXd> values;
  };
  std::vector<PendingDuals> pending_duals_;
  bool use_constraints_;
  bool has_constraints_ = false, history_ = false;
  int history_rows_ = 0;
  altro_b200_solver* solver_ = nullptr;
  std::vector<int> devices_;
  altro_b200_multi* multi_ = nullptr;
  std::vector<float> t_, h_all_;
  bool have_step_ = false;
  int outer_by_hand_ = 0;
  bool take_initial_cost_ = false;
  double penalty_ = 0.0;
  bool have_penalty_ = false;
  double penalty_scaling_ = 0.0;
  bool explicit_x0_ = false;
  std::vector<double> X_, U_, x0_;
  SolverStats stats_;
  Results res_;
};

}  // namespace detail
}  // namespace altro
