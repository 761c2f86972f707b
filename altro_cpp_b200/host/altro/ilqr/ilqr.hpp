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
#include <vector>

#include "altro/device_solver.hpp"

namespace altro {
namespace ilqr {

// Host copy of what GetKnotPointFunction(k) exposes (knot_point_function_type.hpp:243-268
// there): gains after a backward pass or a solve; expansions and cost-to-go after the step-wise
// UpdateExpansions() / BackwardPass().
struct CostExpansionView {
  MatrixXd xx, xu, uu;
  VectorXd x, u;
  const MatrixXd& dxdx() const { return xx; }
  const MatrixXd& dxdu() const { return xu; }
  const MatrixXd& dudu() const { return uu; }
  const VectorXd& dx() const { return x; }
  const VectorXd& du() const { return u; }
};
struct DynamicsExpansionView {
  MatrixXd A, B;
  const MatrixXd& GetA() const { return A; }
  const MatrixXd& GetB() const { return B; }
};
struct KnotPointView {
  MatrixXd K;  // feedback gain, m x n
  VectorXd d;  // feedforward gain, m
  MatrixXd P;  // cost-to-go Hessian
  VectorXd p;  // cost-to-go gradient
  CostExpansionView cost;
  DynamicsExpansionView dynamics;
  const MatrixXd& GetFeedbackGain() const { return K; }
  const VectorXd& GetFeedforwardGain() const { return d; }
  const MatrixXd& GetCostToGoHessian() const { return P; }
  const VectorXd& GetCostToGoGradient() const { return p; }
  const CostExpansionView& GetCostExpansion() const { return cost; }
  const DynamicsExpansionView& GetDynamicsExpansion() const { return dynamics; }
};
using KnotPointGains = KnotPointView;

template <int n, int m>
class iLQR {
 public:
  // iLQR(prob): constraints are ignored unless the problem came out of BuildAugLagProblem
  explicit iLQR(const problem::Problem& prob, int device = 0)
      : core_(std::make_shared<detail::DeviceSolver>(prob, StateDim(prob), ControlDim(prob),
                                                     prob.IsAugmentedLagrangian(), 1, device)) {}
  // iLQR(N) + InitializeFromProblem(prob), ilqr.hpp:50,97-133 there
  explicit iLQR(int N) : N_(N) {}
  void InitializeFromProblem(const problem::Problem& prob, int device = 0) {
    if (prob.NumSegments() != N_ && N_ >= 0) throw std::invalid_argument("Number of segments in problem inconsistent with solver.");
    core_ = std::make_shared<detail::DeviceSolver>(prob, StateDim(prob), ControlDim(prob),
                                                   prob.IsAugmentedLagrangian(), 1, device);
    if (Z_) core_->Upload(*Z_);
  }
  // the inner solver of an AugmentedLagrangianiLQR shares its device state
  explicit iLQR(std::shared_ptr<detail::DeviceSolver> core) : core_(std::move(core)) {}
  // move-only like the reference (ilqr.hpp:56-71): two solvers never share device state by accident
  iLQR(const iLQR&) = delete;
  iLQR& operator=(const iLQR&) = delete;
  iLQR(iLQR&&) = default;
  iLQR& operator=(iLQR&&) = default;

  void SetTrajectory(std::shared_ptr<Trajectory<n, m>> traj) {
    Z_ = std::move(traj);
    if (core_) core_->Upload(*Z_);
  }
  // a zero trajectory of the right shape with the given step, installed as the in/out buffer
  std::shared_ptr<Trajectory<n, m>> MakeTrajectory(float h) {
    auto Z = std::make_shared<Trajectory<n, m>>(core_->n(), core_->m(), core_->NumSegments());
    Z->SetUniformStep(h);
    SetTrajectory(Z);
    return Z;
  }
  // the problem's initial state is shared, not copied (test/ilqr/ilqr_class_test.cpp:84-96 there)
  std::shared_ptr<VectorXd> GetInitialState() const { return core_->GetProblem().GetInitialStatePointer(); }
  // the batch axis replaces the thread pool; these answer the questions perf/ sources ask
  int NumThreads() const { return 1; }
  int NumTasks() const { return 1; }
  std::vector<int> GetTaskAssignment() const { return {0, core_->NumSegments() + 1}; }
  void SetTaskAssignment(std::vector<int>) {}
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
  // one launch covers every knot point; a block request runs the same launch (ilqr.hpp:670-677)
  void UpdateExpansionsBlock(int /*start*/, int /*stop*/) { UpdateExpansions(); }
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

  // with_expansions: also fetch cost/dynamics expansion and cost-to-go (step-wise use only)
  KnotPointView GetKnotPointFunction(int k, bool with_expansions = false) {
    KnotPointView g;
    if (k < core_->NumSegments()) {
      std::vector<double> K, d;
      core_->Gains(0, &K, &d);
      g.K = MatrixXd(m_dim(), n_dim());
      g.d = VectorXd(m_dim());
      for (int j = 0; j < n_dim(); ++j)
        for (int i = 0; i < m_dim(); ++i) g.K(i, j) = K[(static_cast<size_t>(k) * n_dim() + j) * m_dim() + i];
      for (int i = 0; i < m_dim(); ++i) g.d(i) = d[static_cast<size_t>(k) * m_dim() + i];
    }
    if (with_expansions) {
      core_->Expansion(k, 0, &g.dynamics.A, &g.dynamics.B, &g.cost.xx, &g.cost.xu, &g.cost.uu, &g.cost.x, &g.cost.u);
      core_->CostToGo(k, 0, &g.P, &g.p);
    }
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

  int N_ = -1;
  std::shared_ptr<detail::DeviceSolver> core_;
  std::shared_ptr<Trajectory<n, m>> Z_;
};

}  // namespace ilqr
}  // namespace altro
