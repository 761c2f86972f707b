// altro/ilqr/ilqr.hpp (B200 host mirror) — iLQR<n,m> with the reference's public methods
// (altro/ilqr/ilqr.hpp:47-730), every one of them a launch of the device path.
//
//   reference method                   device entry point
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

#include <memory>
#include <utility>

#include "altro/device_solver.hpp"

namespace altro {
namespace ilqr {

// What GetKnotPointFunction(k) exposes to tests (knot_point_function_type.hpp:249-268 there).
struct KnotPointGains {
  MatrixXd K;  // feedback gain, m x n
  VectorXd d;  // feedforward gain, m
  const MatrixXd& GetFeedbackGain() const { return K; }
  const VectorXd& GetFeedforwardGain() const { return d; }
};

template <int n, int m>
class iLQR {
 public:
  // iLQR(prob): constraints are ignored unless the problem came out of BuildAugLagProblem
  explicit iLQR(const problem::Problem& prob, int device = 0)
      : core_(std::make_shared<detail::DeviceSolver>(prob, StateDim(prob), ControlDim(prob),
                                                     prob.IsAugmentedLagrangian(), 1, device)) {}
  // the inner solver of an AugmentedLagrangianiLQR shares its device state
  explicit iLQR(std::shared_ptr<detail::DeviceSolver> core) : core_(std::move(core)) {}

  void SetTrajectory(std::shared_ptr<Trajectory<n, m>> traj) {
    Z_ = std::move(traj);
    core_->Upload(*Z_);
  }
  std::shared_ptr<Trajectory<n, m>> GetTrajectory() const { return Z_; }
  int NumSegments() const { return core_->NumSegments(); }
  SolverOptions& GetOptions() { return core_->GetOptions(); }
  SolverStats& GetStats() { return core_->GetStats(); }
  SolverStatus GetStatus() { return static_cast<SolverStatus>(core_->Pull().ilqr_status[0]); }
  double GetRegularization() { return core_->Pull().reg[0]; }

  void Solve() {
    Require();
    core_->Upload(*Z_);
    core_->Run(detail::DeviceSolver::kSolveILQR);
    core_->Download(Z_.get());
    core_->Pull();
  }
  void Rollout() {
    Require();
    core_->Upload(*Z_);
    core_->Run(detail::DeviceSolver::kRollout);
    core_->Download(Z_.get());
  }
  double Cost() {
    core_->Run(detail::DeviceSolver::kCost);
    return core_->Pull().cost[0];
  }
  void UpdateExpansions() { core_->Run(detail::DeviceSolver::kUpdateExpansions); }
  void BackwardPass() { core_->Run(detail::DeviceSolver::kBackwardPass); }
  void ForwardPass() {
    Require();
    core_->Run(detail::DeviceSolver::kForwardPass);
    core_->Download(Z_.get());
  }
  void UpdateConvergenceStatistics() {
    core_->Run(detail::DeviceSolver::kUpdateConvergenceStatistics);
    core_->Pull();
  }
  void SolveSetup() { core_->Run(detail::DeviceSolver::kSolveSetup); }

  KnotPointGains GetKnotPointFunction(int k) {
    std::vector<double> K, d;
    core_->Gains(0, &K, &d);
    KnotPointGains g;
    g.K = MatrixXd(m_dim(), n_dim());
    g.d = VectorXd(m_dim());
    for (int j = 0; j < n_dim(); ++j)
      for (int i = 0; i < m_dim(); ++i) g.K(i, j) = K[(static_cast<size_t>(k) * n_dim() + j) * m_dim() + i];
    for (int i = 0; i < m_dim(); ++i) g.d(i) = d[static_cast<size_t>(k) * m_dim() + i];
    return g;
  }
  std::shared_ptr<detail::DeviceSolver> Core() const { return core_; }

 private:
  static int StateDim(const problem::Problem& prob) { return prob.GetDynamics(0)->StateDimension(); }
  static int ControlDim(const problem::Problem& prob) { return prob.GetDynamics(0)->ControlDimension(); }
  int n_dim() const { return core_->n(); }
  int m_dim() const { return core_->m(); }
  void Require() const {
    if (!Z_) throw DeviceError(ALTRO_B200_ERR_STATE, "Invalid trajectory pointer. May be uninitialized.");
  }

  std::shared_ptr<detail::DeviceSolver> core_;
  std::shared_ptr<Trajectory<n, m>> Z_;
};

}  // namespace ilqr
}  // namespace altro
