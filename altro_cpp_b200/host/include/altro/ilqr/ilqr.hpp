// altro/ilqr/ilqr.hpp (B200 host mirror) — iLQR<n,m> with the reference's public methods
// (altro/ilqr/ilqr.hpp:47-730 there), every one of them a launch of the device path.
//
//   reference method                   device entry point (include/altro_b200.h)
//   Solve()                ilqr.hpp:284    altro_b200_solve_ilqr
//   Rollout()                      :453    altro_b200_rollout
//   Cost()                         :326    altro_b200_cost
//   UpdateExpansions()             :350    altro_b200_update_expansions
//   BackwardPass()                 :385    altro_b200_backward_pass
//   ForwardPass()                  :512    altro_b200_forward_pass
//   UpdateConvergenceStatistics()  :568    altro_b200_update_convergence_statistics
//   SolveSetup()                   :629    altro_b200_solve_setup
//
// The trajectory handed to SetTrajectory is the in/out buffer: it is uploaded when a solve or a
// rollout starts and overwritten with the device result when one ends.
#pragma once

#include <algorithm>
#include <cmath>
#include <functional>
#include <memory>
#include <utility>
#include <vector>

#include "altro/common/threadpool.hpp"
#include "altro/augmented_lagrangian/al_cost.hpp"
#include "altro/device_solver.hpp"
#include "altro/ilqr/knot_point_function_type.hpp"

namespace altro {
namespace ilqr {

template <int n = Eigen::Dynamic, int m = Eigen::Dynamic>
class iLQR {
 public:
  explicit iLQR(int N) : N_(N) {}
  // iLQR(prob): constraints are ignored unless the problem came out of BuildAugLagProblem
  explicit iLQR(const problem::Problem& prob, int device = 0) : N_(prob.NumSegments()), device_(device) {
    InitializeFromProblem(prob);
  }
  // the inner solver of an AugmentedLagrangianiLQR shares its device state
  explicit iLQR(std::shared_ptr<detail::DeviceSolver> core) : N_(core->NumSegments()), core_(std::move(core)) { MakeKnotPoints(); }
  // move-only like the reference (ilqr.hpp:56-71 there): two solvers never share device state by accident
  iLQR(const iLQR&) = delete;
  iLQR& operator=(const iLQR&) = delete;
  iLQR(iLQR&&) noexcept = default;
  iLQR& operator=(iLQR&&) noexcept = default;

  template <int n2 = n, int m2 = m>
  void CopyFromProblem(const problem::Problem& prob, int k_start, int k_stop) {
    ALTRO_ASSERT(prob.IsFullyDefined(), "Expected problem to be fully defined.");
    ALTRO_ASSERT(k_start == 0 && k_stop == N_ + 1, "The device solver takes the whole horizon at once.");
    ALTRO_UNUSED(k_start);
    ALTRO_UNUSED(k_stop);
    const int nn = prob.GetDynamics(0)->StateDimension(), mm = prob.GetDynamics(0)->ControlDimension();
    ALTRO_ASSERT(n == Eigen::Dynamic || n == nn, "Inconsistent state dimension.");
    ALTRO_ASSERT(m == Eigen::Dynamic || m == mm, "Inconsistent control dimension.");
    core_ = std::make_shared<detail::DeviceSolver>(prob, nn, mm, prob.IsAugmentedLagrangian(), 1, device_);
    MakeKnotPoints();
    if (Z_) core_->Upload(*Z_);
  }
  template <int n2 = n, int m2 = m>
  void InitializeFromProblem(const problem::Problem& prob) {
    ALTRO_ASSERT(prob.NumSegments() == N_, "Number of segments in problem should be equal to the number of segments in the solver.");
    CopyFromProblem<n2, m2>(prob, 0, N_ + 1);
  }

  std::shared_ptr<Trajectory<n, m>> GetTrajectory() { return Z_; }
  int NumSegments() const { return N_; }
  SolverStats& GetStats() { return Core().GetStats(); }
  SolverOptions& GetOptions() { return Core().GetOptions(); }
  // the device's verdict after Solve(); the verdict of IsDone() when the caller runs the iterations one by one
  SolverStatus GetStatus() {
    return by_hand_ ? status_by_hand_ : static_cast<SolverStatus>(Core().Pull().ilqr_status[0]);
  }
  // costs_(k) of the last Cost() / UpdateExpansions(), augmented-Lagrangian terms included (ilqr.hpp:163 there)
  VectorXd& GetCosts() {
    const std::vector<double> c = Core().Costs(0);
    costs_ = VectorXd::Zero(static_cast<int>(c.size()));
    for (int k = 0; k < costs_.size(); ++k) costs_(k) = c[static_cast<size_t>(k)];
    return costs_;
  }
  // the problem's initial state is shared, not copied (test/ilqr/ilqr_class_test.cpp:84-96 there)
  std::shared_ptr<VectorXd> GetInitialState() { return Core().GetProblem().GetInitialStatePointer(); }
  double GetRegularization() { return Core().Pull().reg[0]; }

  // Gains after a backward pass or a solve; expansions and cost-to-go after the step-wise
  // UpdateExpansions() / BackwardPass().  Refreshed from the device on every call.
  KnotPointFunctions<n, m>& GetKnotPointFunction(int k) {
    ALTRO_ASSERT(k >= 0 && k <= N_, "Invalid knot point index.");
    handed_out_.at(k) = 1;  // a reference the caller may keep: refreshed again after every device phase
    RefreshKnotPoint(k);
    return *knotpoints_.at(k);
  }

  // The batch axis replaces the thread pool: one launch covers every knot point.  These keep the
  // reference's bookkeeping (DefaultTaskAssignment, ilqr.hpp:740-751 there) so that programs tuning
  // nthreads / tasks_per_thread run unchanged; they have no effect on the device.
  std::vector<int>& GetTaskAssignment() {
    if (!custom_work_assignment_) DefaultTaskAssignment();
    return work_inds_;
  }
  void SetTaskAssignment(std::vector<int> inds) {
    work_inds_ = std::move(inds);
    custom_work_assignment_ = true;
  }
  size_t NumThreads() const { return static_cast<size_t>(nthreads_launched_); }
  int NumTasks() const { return static_cast<int>(work_inds_.size()) - 1; }

  // a zero trajectory of the right shape with the given step, installed as the in/out buffer
  std::shared_ptr<Trajectory<n, m>> MakeTrajectory(float dt) {
    auto Z = std::make_shared<Trajectory<n, m>>(Core().n(), Core().m(), N_);
    Z->SetUniformStep(dt);
    SetTrajectory(Z);
    return Z_;
  }
  void SetTrajectory(std::shared_ptr<Trajectory<n, m>> traj) {
    Z_ = std::move(traj);
    if (core_ && Z_) core_->Upload(*Z_);
  }
  void SetConstraintCallback(const std::function<double()>& max_violation) { max_violation_callback_ = max_violation; }

  void Solve() {
    Require();
    by_hand_ = false;
    Core().Upload(*Z_);
    Core().Run(detail::DeviceSolver::kSolveILQR);
    Core().Download(Z_.get());
    Core().Pull();
    Core().PullHistory(/*al=*/false);
    RefreshHandedOut();
  }
  void Rollout() {
    Require();
    Core().Upload(*Z_);
    Core().Run(detail::DeviceSolver::kRollout);
    Core().Download(Z_.get());
  }
  // The trajectory is shared with the caller (ilqr.hpp:226-235 there), who may have edited it in place: the phases
  // that read it as their input take the host copy to the device first.
  double Cost() {
    if (Z_) Core().Upload(*Z_);
    Core().Run(detail::DeviceSolver::kCost);
    return Core().Pull().cost[0];
  }
  double Cost(const Trajectory<n, m>& Z) {
    Core().Upload(Z);
    Core().Run(detail::DeviceSolver::kCost);
    return Core().Pull().cost[0];
  }
  void UpdateExpansions() {
    SyncThreadBookkeeping();
    if (Z_) Core().Upload(*Z_);
    Core().Run(detail::DeviceSolver::kUpdateExpansions);
    RefreshHandedOut();
  }
  // one launch covers every knot point; a block request runs the same launch (ilqr.hpp:670-677 there)
  void UpdateExpansionsBlock(int start, int stop) {
    ALTRO_UNUSED(start);
    ALTRO_UNUSED(stop);
    Core().Run(detail::DeviceSolver::kUpdateExpansions);
  }
  void BackwardPass() {
    Core().Run(detail::DeviceSolver::kBackwardPass);
    RefreshHandedOut();
  }
  void ForwardPass() {
    Require();
    Core().Run(detail::DeviceSolver::kForwardPass);
    Core().Download(Z_.get());
    const auto sc = Core().Scalars();  // stats_.Log("alpha" / "z"), ilqr.hpp:545-547 (carry-forward when the search failed)
    GetStats().Log("alpha", sc.alpha);
    GetStats().Log("z", sc.z);
  }
  // The first iteration's decrease is measured against stats.initial_cost, which the caller of the step-wise
  // methods assigns (`GetStats().initial_cost = Cost()`, as Solve() does there, ilqr.hpp:292): it goes to the device.
  void UpdateConvergenceStatistics() {
    Core().Pull();
    if (GetStats().iterations_inner == 0) Core().SetInitialCost(GetStats().initial_cost);
    Core().Run(detail::DeviceSolver::kUpdateConvergenceStatistics);
    Core().Pull();
    const auto sc = Core().Scalars();  // ilqr.hpp:578-584
    GetStats().Log("dJ", sc.dJ);
    GetStats().Log("grad", sc.grad);
    GetStats().NewIteration();
  }
  void SolveSetup() {
    SyncThreadBookkeeping();
    by_hand_ = false;
    Core().Run(detail::DeviceSolver::kSolveSetup);
    Core().Pull();
  }
  // ilqr.hpp:597-619 there, on the statistics of the last UpdateConvergenceStatistics()
  bool IsDone() {
    Core().Pull();
    const SolverStats& stats = GetStats();
    const SolverOptions& opts = GetOptions();
    const auto sc = Core().Scalars();
    const SolverStatus device = static_cast<SolverStatus>(Core().Last().ilqr_status[0]);
    by_hand_ = true;
    status_by_hand_ = device;
    if (sc.dJ < opts.cost_tolerance && sc.grad < opts.gradient_tolerance) status_by_hand_ = SolverStatus::kSolved;
    else if (stats.iterations_inner >= opts.max_iterations_inner) status_by_hand_ = SolverStatus::kMaxInnerIterations;
    else if (stats.iterations_total >= opts.max_iterations_total) status_by_hand_ = SolverStatus::kMaxIterations;
    return status_by_hand_ != SolverStatus::kUnsolved;
  }
  void WrapUp() {}
  // mean over the knot points of max_i |d_i| / (|u_i| + 1), from the device's gains and the current controls
  double NormalizedFeedforwardGain() {
    Require();
    std::vector<double> K, d;
    Core().Gains(0, &K, &d);
    const int mm = Core().m();
    double sum = 0.0;
    for (int k = 0; k < N_; ++k) {
      double worst = 0.0;
      for (int i = 0; i < mm; ++i)
        worst = std::max(worst, std::fabs(d[static_cast<size_t>(k) * mm + i]) / (std::fabs(Z_->Control(k)(i)) + 1));
      sum += worst;
    }
    return sum / N_;
  }
  std::shared_ptr<detail::DeviceSolver> CorePtr() const { return core_; }

 private:
  detail::DeviceSolver& Core() {
    if (!core_) throw DeviceError(ALTRO_B200_ERR_STATE, "the solver has no problem yet (InitializeFromProblem)");
    return *core_;
  }
  void Require() const {
    // a contract violation first of all (abort in debug builds, like the reference), an error in every build
    ALTRO_ASSERT(Z_ != nullptr, "Invalid trajectory pointer. May be uninitialized.");
    if (!Z_) throw DeviceError(ALTRO_B200_ERR_STATE, "Invalid trajectory pointer. May be uninitialized.");
  }
  void RefreshKnotPoint(int k) {
    KnotPointFunctions<n, m>& kpf = *knotpoints_.at(k);
    detail::DeviceSolver& c = Core();
    if (!c.Ready()) return;
    const int nn = c.n(), mm = c.m();
    if (k < N_) {
      std::vector<double> K, d;
      c.Gains(0, &K, &d);
      for (int j = 0; j < nn; ++j)
        for (int i = 0; i < mm; ++i) kpf.GetFeedbackGain()(i, j) = K[(static_cast<size_t>(k) * nn + j) * mm + i];
      for (int i = 0; i < mm; ++i) kpf.GetFeedforwardGain()(i) = d[static_cast<size_t>(k) * mm + i];
    }
    MatrixXd A, B;
    CostExpansion<n, m>& e = kpf.GetCostExpansion();
    if (c.TryExpansion(k, 0, &A, &B, &e.dxdx(), &e.dxdu(), &e.dudu(), &e.dx(), &e.du())) {
      MatrixXd& J = kpf.GetDynamicsExpansion().GetJacobian();
      J.topLeftCorner(nn, nn) = A;
      J.topRightCorner(nn, mm) = B;
    }
    c.TryCostToGo(k, 0, &kpf.GetCostToGoHessian(), &kpf.GetCostToGoGradient());
  }
  // references returned by GetKnotPointFunction stay current across device phases, like the reference's (which
  // point at the solver's own storage); knot points nobody asked for are not fetched
  void RefreshHandedOut() {
    for (int k = 0; k <= N_; ++k)
      if (handed_out_[k]) RefreshKnotPoint(k);
  }
  void MakeKnotPoints() {
    handed_out_.assign(static_cast<size_t>(N_) + 1, 0);
    knotpoints_.clear();
    const problem::Problem& prob = core_->GetProblem();
    // For an augmented-Lagrangian problem the cost object of a knot point is its ALCost (as in the reference, where
    // BuildAugLagProblem wraps every cost): GetCostFunPtr() then gives access to the duals and penalty of the knot.
    for (int k = 0; k <= N_; ++k) {
      std::shared_ptr<problem::CostFunction> cost = prob.GetCostFunction(k);
      if (core_->UsesConstraints()) cost = std::make_shared<augmented_lagrangian::ALCost<n, m>>(core_, k);
      knotpoints_.emplace_back(std::make_unique<KnotPointFunctions<n, m>>(prob.GetDynamics(k), cost));
    }
  }
  void SyncThreadBookkeeping() {
    nthreads_launched_ = core_ ? core_->GetOptions().NumThreads() : 1;
    if (nthreads_launched_ <= 1) nthreads_launched_ = 0;  // the reference launches no pool for one thread
    if (!custom_work_assignment_) DefaultTaskAssignment();
  }
  void DefaultTaskAssignment() {
    const SolverOptions& o = core_ ? core_->GetOptions() : SolverOptions();
    const int ntasks = std::max(1, o.NumThreads() * std::max(1, o.tasks_per_thread));
    const double step = (N_ + 1) / static_cast<double>(ntasks);
    work_inds_.clear();
    for (int i = 0; i <= ntasks; ++i) work_inds_.push_back(static_cast<int>(std::round(i * step)));
    work_inds_.back() = N_ + 1;
  }

  int N_ = 0;
  int device_ = 0;
  std::shared_ptr<detail::DeviceSolver> core_;
  std::shared_ptr<Trajectory<n, m>> Z_;
  std::vector<std::unique_ptr<KnotPointFunctions<n, m>>> knotpoints_;
  std::vector<char> handed_out_;  // knot points whose KnotPointFunctions reference was given to the caller
  VectorXd costs_;
  bool by_hand_ = false;  // the last verdict came from IsDone(), not from a whole-solve launch
  SolverStatus status_by_hand_ = SolverStatus::kUnsolved;
  std::vector<int> work_inds_ = {0, 1};
  bool custom_work_assignment_ = false;
  int nthreads_launched_ = 0;
  std::function<double()> max_violation_callback_;
};

}  // namespace ilqr
}  // namespace altro
