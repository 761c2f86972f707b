// Differential probe: the single-knot host arithmetic that is public API in the reference — ALCost::Evaluate /
// Gradient / Hessian with multipliers and penalties, KnotPointFunctions::CalcActionValueExpansion / CalcGains /
// CalcCostToGo — evaluated on fixed inputs and printed with 17 digits.  tests/test_host_arith_differential.py builds
// this file twice, once against this repo's host mirror and once against the reference's own headers and sources, and
// compares the two outputs.  Uses only names both trees have.
#include <cstdio>
#include <memory>

#include "altro/augmented_lagrangian/al_cost.hpp"
#include "altro/ilqr/ilqr.hpp"
#include "altro/problem/discretized_model.hpp"
#include "examples/problems/unicycle.hpp"
#include "examples/quadratic_cost.hpp"
#include "examples/triple_integrator.hpp"

namespace {

void Show(const char* name, const altro::MatrixXd& M) {
  std::printf("%s %dx%d\n", name, static_cast<int>(M.rows()), static_cast<int>(M.cols()));
  for (int i = 0; i < M.rows(); ++i) {
    for (int j = 0; j < M.cols(); ++j) std::printf(" %.17g", M(i, j));
    std::printf("\n");
  }
}
void Show(const char* name, double v) { std::printf("%s %.17g\n", name, v); }

// deterministic "random" numbers that do not depend on either tree's Random()
double Next() {
  static unsigned long long state = 88172645463325252ULL;
  state ^= state << 13;
  state ^= state >> 7;
  state ^= state << 17;
  return static_cast<double>(state % 2000001ULL) / 1000000.0 - 1.0;
}
altro::MatrixXd Fill(int rows, int cols) {
  altro::MatrixXd M = altro::MatrixXd::Zero(rows, cols);
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i) M(i, j) = Next();
  return M;
}

void ProbeALCost(int k) {
  altro::problems::UnicycleProblem def;
  def.SetScenario(altro::problems::UnicycleProblem::kThreeObstacles);
  altro::problem::Problem prob = def.MakeProblem();
  altro::augmented_lagrangian::ALCost<3, 2> alcost(prob, k);
  std::printf("knot %d: %d constraint rows\n", k, alcost.NumConstraints());
  for (auto& con : alcost.GetEqualityConstraints()) {
    con->SetPenalty(3.75);
    for (int i = 0; i < con->GetDuals().size(); ++i) con->GetDuals()(i) = Next();
  }
  for (auto& con : alcost.GetInequalityConstraints()) {
    con->SetPenalty(7.5);
    for (int i = 0; i < con->GetDuals().size(); ++i) con->GetDuals()(i) = -0.5 + 0.75 * Next();  // both signs
  }
  altro::VectorXd x(3), u(2);
  x << 0.7, 0.9, 0.4;   // inside the first obstacle's reach: active and inactive rows
  u << 2.1, -1.8;       // beyond the control bounds on both sides
  Show("value", alcost.Evaluate(x, u));
  altro::VectorXd dx = altro::VectorXd::Zero(3), du = altro::VectorXd::Zero(2);
  alcost.Gradient(x, u, dx, du);
  Show("dx", dx);
  Show("du", du);
  altro::MatrixXd dxdx = altro::MatrixXd::Zero(3, 3), dxdu = altro::MatrixXd::Zero(3, 2), dudu = altro::MatrixXd::Zero(2, 2);
  alcost.Hessian(x, u, dxdx, dxdu, dudu);
  Show("dxdx", dxdx);
  Show("dxdu", dxdu);
  Show("dudu", dudu);
  alcost.UpdateDuals();
  alcost.UpdatePenalties();
  Show("value after the dual and penalty updates", alcost.Evaluate(x, u));
  Show("max violation", alcost.MaxViolation());
  Show("max penalty", alcost.MaxPenalty());
}

void ProbeKnotPointFunctions() {
  constexpr int dof = 2, n = 3 * dof, m = dof;
  using Model = altro::problem::DiscretizedModel<altro::examples::TripleIntegrator>;
  std::shared_ptr<Model> model = std::make_shared<Model>(Model(altro::examples::TripleIntegrator(dof)));
  altro::MatrixXd Q = Fill(n, n), R = Fill(m, m);
  Q = Q.transpose() * Q;
  R = R.transpose() * R;
  for (int i = 0; i < m; ++i) R(i, i) += 1.0;
  std::shared_ptr<altro::examples::QuadraticCost> cost = std::make_shared<altro::examples::QuadraticCost>(
      Q, R, altro::MatrixXd::Zero(n, m), altro::VectorXd(Fill(n, 1)), altro::VectorXd(Fill(m, 1)), 0.25);
  altro::ilqr::KnotPointFunctions<n, m> kpf(model, cost);
  altro::VectorXd x = Fill(n, 1), u = Fill(m, 1);
  kpf.CalcCostExpansion(x, u);
  kpf.CalcDynamicsExpansion(x, u, 0.3F, 0.1F);
  altro::MatrixXd S = Fill(n, n);
  S = S.transpose() * S;
  altro::MatrixXd s = Fill(n, 1);
  kpf.CalcActionValueExpansion(S, s);
  Show("Qxx", kpf.GetActionValueExpansion().dxdx());
  Show("Qxu", kpf.GetActionValueExpansion().dxdu());
  Show("Quu", kpf.GetActionValueExpansion().dudu());
  Show("Qx", kpf.GetActionValueExpansion().dx());
  Show("Qu", kpf.GetActionValueExpansion().du());
  kpf.RegularizeActionValue(1e-3);
  std::printf("CalcGains -> %d\n", static_cast<int>(kpf.CalcGains()));
  Show("K", kpf.GetFeedbackGain());
  Show("d", kpf.GetFeedforwardGain());
  kpf.CalcCostToGo();
  Show("P", kpf.GetCostToGoHessian());
  Show("p", kpf.GetCostToGoGradient());
  Show("deltaV(1)", kpf.GetCostToGoDelta());
  Show("deltaV(0.25)", kpf.GetCostToGoDelta(0.25));
  kpf.CalcTerminalCostToGo();
  Show("terminal P", kpf.GetCostToGoHessian());
}

}  // namespace

int main() {
  ProbeALCost(1);
  ProbeALCost(100);
  ProbeKnotPointFunctions();
  return 0;
}
