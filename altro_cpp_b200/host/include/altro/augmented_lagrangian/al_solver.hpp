// altro/augmented_lagrangian/al_solver.hpp (B200 host mirror) — AugmentedLagrangianiLQR<n,m>
// (reference: altro/augmented_lagrangian/al_solver.hpp:28-440) for one instance, and
// BatchedAugmentedLagrangianiLQR<n,m>, the form the device is built for: B instances of one
// problem that differ in initial state, solved by one set of kernel launches.
//
//   Solve()            al_solver.hpp:304   altro_b200_solve_al
//   UpdateDuals()                 :336     altro_b200_update_duals
//   UpdatePenalties()             :347     altro_b200_update_penalties
//   SetPenalty(rho)               :271     altro_b200_solver_set_penalty
//   MaxViolation()                :417     altro_b200_get_results_host (viol)
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <memory>
#include <utility>
#include <vector>

#include "altro/augmented_lagrangian/al_cost.hpp"
#include "altro/augmented_lagrangian/al_problem.hpp"
#include "altro/ilqr/ilqr.hpp"

namespace altro {
namespace augmented_lagrangian {

template <int n, int m>
class AugmentedLagrangianiLQR {
 public:
  // AugmentedLagrangianiLQR(N) + InitializeFromProblem(prob), al_solver.hpp:35-38 there
  explicit AugmentedLagrangianiLQR(int N) : ilqr_solver_(N) {}
  explicit AugmentedLagrangianiLQR(const problem::Problem& prob, int device = 0)
      : device_(device), ilqr_solver_(prob.NumSegments()) {
    InitializeFromProblem(prob);
  }
  void InitializeFromProblem(const problem::Problem& prob) {
    ALTRO_ASSERT(prob.IsFullyDefined(), "Expected problem to be fully defined.");
    std::shared_ptr<Trajectory<n, m>> Z = ilqr_solver_.GetTrajectory();
    core_ = std::make_shared<detail::DeviceSolver>(prob, prob.GetDynamics(0)->StateDimension(),
                                                   prob.GetDynamics(0)->ControlDimension(), true, 1, device_);
    ilqr_solver_ = ilqr::iLQR<n, m>(core_);
    costs_.clear();
    for (int k = 0; k <= prob.NumSegments(); ++k) costs_.emplace_back(std::make_shared<ALCost<n, m>>(core_, k));
    if (Z) ilqr_solver_.SetTrajectory(Z);
  }
  std::shared_ptr<ALCost<n, m>> GetALCost(int k) { return costs_.at(k); }

  void SetTrajectory(std::shared_ptr<Trajectory<n, m>> traj) { ilqr_solver_.SetTrajectory(std::move(traj)); }
  ilqr::iLQR<n, m>& GetiLQRSolver() { return ilqr_solver_; }
  SolverOptions& GetOptions() { return Initialized().GetOptions(); }
  SolverStats& GetStats() { return Initialized().GetStats(); }
  // the device's verdict after Solve(); the verdict of IsDone() when the caller runs the outer loop step by step
  SolverStatus GetStatus() {
    return by_hand_ ? status_by_hand_ : static_cast<SolverStatus>(Initialized().Pull().status[0]);
  }
  int NumSegments() const { return ilqr_solver_.NumSegments(); }

  void SetPenalty(double rho) { core_->SetPenalty(rho); }
  void SetPenaltyScaling(double phi) { core_->SetPenaltyScaling(phi); }

  void Solve() {
    auto Z = ilqr_solver_.GetTrajectory();
    ALTRO_ASSERT(Z != nullptr, "Invalid trajectory pointer. May be uninitialized.");
    if (!Z) throw DeviceError(ALTRO_B200_ERR_STATE, "Invalid trajectory pointer. May be uninitialized.");
    by_hand_ = false;
    core_->Upload(*Z);
    core_->Run(detail::DeviceSolver::kSolveAL);
    core_->Download(Z.get());
    core_->Pull();
    core_->PullHistory();
  }
  // what Solve() does first (al_solver.hpp:287-302 there): duals reset if reset_duals, penalties set to
  // initial_penalty if that is positive, statistics cleared.  Needs the trajectory (it sizes the device solver).
  void Init() {
    auto Z = ilqr_solver_.GetTrajectory();
    if (!Z) throw DeviceError(ALTRO_B200_ERR_STATE, "Invalid trajectory pointer. May be uninitialized.");
    by_hand_ = false;
    core_->Upload(*Z);
    core_->Run(detail::DeviceSolver::kAlInit);
    SolverStats& stats = GetStats();
    stats.Reset();
    stats.Log("iter_al", 0);
    stats.Log("viol", MaxViolation());
    stats.Log("pen", GetMaxPenalty());
  }
  // ---- the outer loop step by step (what Solve() runs in one launch; al_solver.hpp:304-334 there):
  //   Init(); loop { GetiLQRSolver().Solve(); UpdateDuals(); UpdateConvergenceStatistics(); if (IsDone()) break;
  //                  UpdatePenalties(); }
  void UpdateDuals() { core_->Run(detail::DeviceSolver::kUpdateDuals); }
  void UpdatePenalties() { core_->Run(detail::DeviceSolver::kUpdatePenalties); }
  void ResetDualVariables() { Initialized().ZeroDuals(); }
  void UpdateConvergenceStatistics() {
    Initialized().CountOuterIterationByHand();
    const double viol = GetMaxViolation();  // pulls the counters: iterations_outer now includes this iteration
    SolverStats& stats = GetStats();
    stats.Log("viol", viol);
    stats.Log("pen", GetMaxPenalty());
    stats.Log("iter_al", stats.iterations_outer);
  }
  // the termination tests of al_solver.hpp:368-400 there, on the statistics logged last
  bool IsDone() {
    SolverStats& stats = GetStats();
    const SolverOptions& opts = GetOptions();
    const SolverStatus inner = ilqr_solver_.GetStatus();
    by_hand_ = true;
    if (inner != SolverStatus::kSolved) return Verdict(inner);
    if (!stats.violations.empty() && stats.violations.back() < opts.constraint_tolerance) return Verdict(SolverStatus::kSolved);
    if (!stats.max_penalty.empty() && stats.max_penalty.back() > opts.maximum_penalty) return Verdict(SolverStatus::kMaxPenalty);
    if (stats.iterations_outer >= opts.max_iterations_outer) return Verdict(SolverStatus::kMaxOuterIterations);
    if (stats.iterations_total >= opts.max_iterations_total) return Verdict(SolverStatus::kMaxIterations);
    status_by_hand_ = SolverStatus::kUnsolved;
    return false;
  }
  double MaxViolation() {
    core_->Run(detail::DeviceSolver::kCost);
    return core_->Pull().viol[0];
  }
  // of another trajectory: it becomes the solver's current iterate on the device, as ilqr.Cost(Z) does there
  double MaxViolation(const Trajectory<n, m>& Z) {
    ilqr_solver_.Cost(Z);
    return core_->Pull().viol[0];
  }
  double GetMaxViolation() { return MaxViolation(); }
  double GetMaxPenalty() { return core_->MaxPenalty(0); }
  int NumConstraints(int k) const { return Initialized().GetProblem().NumConstraints(k); }
  int NumConstraints() const { return Initialized().GetProblem().NumConstraints(); }
  // dual variables of knot k, equalities then inequalities (GetALCost(k)->...->GetDuals() there)
  VectorXd GetDuals(int k) {
    const std::vector<double> lam = core_->Duals(k, 0);
    VectorXd out = VectorXd::Zero(static_cast<int>(lam.size()));
    for (int i = 0; i < out.size(); ++i) out(i) = lam[i];
    return out;
  }

  // al_solver.hpp:85-104 there: one entry per constraint and knot point with its violation
  // c - Pi_K(c) of the current trajectory, optionally sorted by its infinity norm
  std::vector<constraints::ConstraintInfo> GetConstraintInfo(bool should_sort = false) {
    const problem::Problem& prob = core_->GetProblem();
    std::vector<constraints::ConstraintInfo> coninfo;
    for (int k = 0; k <= NumSegments(); ++k) {
      if (prob.NumConstraints(k) == 0) continue;
      const std::vector<double> c = core_->ConstraintValues(k, 0);
      int row = 0;
      for (const auto& con : prob.GetEqualityConstraints()[k]) {
        constraints::ConstraintInfo info{con->GetLabel(), k, VectorXd::Zero(con->OutputDimension()), con->GetConstraintType()};
        for (int i = 0; i < info.violation.size(); ++i) info.violation(i) = c.at(row++);
        coninfo.push_back(info);
      }
      for (const auto& con : prob.GetInequalityConstraints()[k]) {
        constraints::ConstraintInfo info{con->GetLabel(), k, VectorXd::Zero(con->OutputDimension()), con->GetConstraintType()};
        for (int i = 0; i < info.violation.size(); ++i) {
          const double ci = c.at(row++);
          info.violation(i) = ci > 0.0 ? ci : 0.0;
        }
        coninfo.push_back(info);
      }
    }
    if (should_sort)
      std::stable_sort(coninfo.begin(), coninfo.end(),
                       [](const constraints::ConstraintInfo& a, const constraints::ConstraintInfo& b) {
                         return InfNorm(a.violation) > InfNorm(b.violation);
                       });
    return coninfo;
  }
  void PrintViolations(bool should_sort = false, int precision = 4) {
    const std::vector<constraints::ConstraintInfo> coninfo = GetConstraintInfo(should_sort);
    std::printf("Got %d constraints\n", static_cast<int>(coninfo.size()));
    for (const constraints::ConstraintInfo& info : coninfo) std::printf("%s\n", info.ToString(precision).c_str());
  }

 private:
  // a solver made with the (N) constructor has no device state until InitializeFromProblem
  detail::DeviceSolver& Initialized() const {
    ALTRO_ASSERT(core_ != nullptr, "Solver is empty: finish initializing the solver with a problem (InitializeFromProblem).");
    if (!core_) throw DeviceError(ALTRO_B200_ERR_STATE, "the solver has no problem yet (InitializeFromProblem)");
    return *core_;
  }
  bool Verdict(SolverStatus status) {
    status_by_hand_ = status;
    return true;
  }
  static double InfNorm(const VectorXd& v) {
    double r = 0.0;
    for (int i = 0; i < v.size(); ++i) r = std::fabs(v(i)) > r ? std::fabs(v(i)) : r;
    return r;
  }
  int device_ = 0;
  bool by_hand_ = false;  // the last verdict came from IsDone(), not from a whole-solve launch
  SolverStatus status_by_hand_ = SolverStatus::kUnsolved;
  std::shared_ptr<detail::DeviceSolver> core_;
  ilqr::iLQR<n, m> ilqr_solver_;
  std::vector<std::shared_ptr<ALCost<n, m>>> costs_;
};

// The batch axis replaces the reference's thread pool (SolverOptions::nthreads): instance b
// starts from initial state x0[b] and the controls of the trajectory given to SetTrajectory.
template <int n, int m>
class BatchedAugmentedLagrangianiLQR {
 public:
  BatchedAugmentedLagrangianiLQR(const problem::Problem& prob, int batch, int device = 0)
      : core_(std::make_shared<detail::DeviceSolver>(prob, prob.GetDynamics(0)->StateDimension(),
                                                     prob.GetDynamics(0)->ControlDimension(), true, batch, device)) {}
  // the batch cut into contiguous slices over several GPUs of the node (one solver and one host thread
  // per device, no collective; what SolverOptions::nthreads is to the reference, one level up)
  BatchedAugmentedLagrangianiLQR(const problem::Problem& prob, int batch, const std::vector<int>& devices)
      : core_(std::make_shared<detail::DeviceSolver>(prob, prob.GetDynamics(0)->StateDimension(),
                                                     prob.GetDynamics(0)->ControlDimension(), true, batch,
                                                     devices.empty() ? 0 : devices[0], devices)) {}

  int Batch() const { return core_->Batch(); }
  SolverOptions& GetOptions() { return core_->GetOptions(); }
  void SetPenalty(double rho) { core_->SetPenalty(rho); }
  void SetInitialStates(const std::vector<VectorXd>& x0) { core_->SetInitialStates(x0); }
  void SetTrajectory(std::shared_ptr<Trajectory<n, m>> nominal) {
    Z_ = std::move(nominal);
    core_->SetStep(Z_->GetStep(0));
  }
  void Solve() {
    if (!Z_) throw DeviceError(ALTRO_B200_ERR_STATE, "Invalid trajectory pointer. May be uninitialized.");
    core_->Upload(*Z_);
    core_->Run(detail::DeviceSolver::kSolveAL);
    core_->Fetch();
    core_->Pull();
  }
  SolverStatus GetStatus(int b) const { return static_cast<SolverStatus>(core_->Last().status.at(b)); }
  int GetIterations(int b) const { return core_->Last().iters.at(static_cast<size_t>(b) * 3 + 2); }
  int GetOuterIterations(int b) const { return core_->Last().iters.at(static_cast<size_t>(b) * 3 + 1); }
  double GetCost(int b) const { return core_->Last().cost.at(b); }
  double GetMaxViolation(int b) const { return core_->Last().viol.at(b); }
  Trajectory<n, m> GetTrajectory(int b) const {
    Trajectory<n, m> Z = *Z_;
    core_->CopyOut(&Z, b);
    return Z;
  }
  int64_t KernelLaunches() const { return core_->KernelLaunches(); }

 private:
  std::shared_ptr<detail::DeviceSolver> core_;
  std::shared_ptr<Trajectory<n, m>> Z_;
};

}  // namespace augmented_lagrangian
}  // namespace altro
