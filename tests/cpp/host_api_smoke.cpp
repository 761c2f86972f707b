// Host C++ (the reference's language) against the C ABI through include/altro_b200.hpp:
// builds the reference's 3-obstacle unicycle problem (examples/problems/unicycle.cpp:27-83) and
// solves a small batch.  Without a GPU it must fail loudly with ALTRO_B200_ERR_CUDA.
#include <cmath>
#include <cstdio>
#include <vector>

#include "altro_b200.hpp"

int main() {
  const int n = 3, m = 2, N = 100, B = 64;
  try {
    altro_b200::Problem prob(n, m, N);
    prob.SetDynamics(ALTRO_B200_MODEL_UNICYCLE);
    const float tf = 5.0f;
    const float h = tf / N;
    prob.SetUniformStep(h);
    const double xf[3] = {3, 3, 0};
    auto diag = [](int d, double v) {
      std::vector<double> M(d * d, 0.0);
      for (int i = 0; i < d; ++i) M[i + i * d] = v;
      return M;
    };
    // QuadraticCost::LQRCost(Q, R, xf, uref = 0), examples/quadratic_cost.hpp:29-39
    auto lqr = [&](double qv, double rv, int k0, int k1) {
      std::vector<double> Q = diag(n, qv), R = diag(m, rv), H(n * m, 0.0), q(n), r(m, 0.0);
      double c = 0.0;
      for (int i = 0; i < n; ++i) {
        q[i] = -(qv * xf[i]);
        c += 0.5 * xf[i] * (qv * xf[i]);
      }
      prob.SetCostFunction(k0, k1, Q.data(), R.data(), H.data(), q.data(), r.data(), c);
    };
    lqr(1.0 * h, 0.5 * h, 0, N);
    lqr(10.0, 0.0, N, N + 1);
    const std::vector<double> cx = {0.75, 1.5, 2.25}, cr = {0.425, 0.425, 0.425};
    for (int k = 1; k < N; ++k) prob.SetCircleConstraint(k, cx, cx, cr);
    const double lb[2] = {0, -3}, ub[2] = {3, 3};
    for (int k = 0; k < N; ++k) prob.SetControlBound(k, lb, ub);
    prob.SetGoalConstraint(N, xf);
    const double x0n[3] = {0, 0, 0};
    prob.SetInitialState(x0n);

    altro_b200::BatchedAugmentedLagrangianiLQR solver(prob, B);
    std::vector<double> x0(B * n, 0.0);
    for (int b = 1; b < B; ++b) x0[b * n + 2] = 0.01 * b;  // small heading perturbations
    const double u0[2] = {0.01, 0.01};
    solver.SetTrajectory(x0.data(), nullptr, u0);
    solver.Solve();
    std::vector<double> cost, viol;
    std::vector<int32_t> status, iters;
    solver.GetResults(&cost, &viol, &status, &iters);
    std::printf("instance 0: status %d, iterations total/outer %d/%d, cost %.12f, viol %.3e\n", status[0],
                iters[2], iters[1], cost[0], viol[0]);
    // nominal instance: 50 iLQR iterations in 5 AL iterations, solved
    return (status[0] == ALTRO_B200_SOLVED && iters[2] == 50 && iters[1] == 5) ? 0 : 2;
  } catch (const altro_b200::Error& e) {
    std::printf("altro_b200 error %d: %s\n", e.code, e.what());
    return e.code == ALTRO_B200_ERR_CUDA ? 3 : 4;
  }
}
