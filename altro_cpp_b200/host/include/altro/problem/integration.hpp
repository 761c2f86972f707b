// altro/problem/integration.hpp (B200 host mirror) — explicit integrators over a user's
// ContinuousDynamics (altro/problem/integration.hpp:27,63,113 there).
//
// These are host-side evaluations of the user's plug-in, part of the public API
// (DiscretizedModel::Evaluate / Jacobian).  The solvers do not use them: rollouts and expansions run
// in csrc/device.cuh (rk4_step, rk4_jacobian) on the device.
#pragma once

#include <array>
#include <memory>
#include <vector>

#include "altro/common/state_control_sized.hpp"
#include "altro/eigentypes.hpp"
#include "altro/problem/dynamics.hpp"

namespace altro {
namespace problem {

template <int NStates, int NControls>
class ExplicitIntegrator : public StateControlSized<NStates, NControls> {
 protected:
  using DynamicsPtr = std::shared_ptr<ContinuousDynamics>;

 public:
  ExplicitIntegrator(int n, int m) : StateControlSized<NStates, NControls>(n, m) {}
  ExplicitIntegrator() : StateControlSized<NStates, NControls>() {}
  virtual ~ExplicitIntegrator() = default;
  virtual void Integrate(const DynamicsPtr& dynamics, const VectorXdRef& x, const VectorXdRef& u, float t, float h,
                         Eigen::Ref<VectorXd> xnext) = 0;
  virtual void Jacobian(const DynamicsPtr& dynamics, const VectorXdRef& x, const VectorXdRef& u, float t, float h,
                        Eigen::Ref<MatrixXd> jac) = 0;
};

class ExplicitEuler final : public ExplicitIntegrator<Eigen::Dynamic, Eigen::Dynamic> {
 public:
  ExplicitEuler(int n, int m) : ExplicitIntegrator<Eigen::Dynamic, Eigen::Dynamic>(n, m) {}
  void Integrate(const DynamicsPtr& dynamics, const VectorXdRef& x, const VectorXdRef& u, float t, float h,
                 Eigen::Ref<VectorXd> xnext) override {
    VectorXd xdot = VectorXd::Zero(x.size());
    dynamics->Evaluate(x, u, t, xdot);
    xnext = x + xdot * static_cast<double>(h);
  }
  void Jacobian(const DynamicsPtr& dynamics, const VectorXdRef& x, const VectorXdRef& u, float t, float h,
                Eigen::Ref<MatrixXd> jac) override {
    const int n = static_cast<int>(x.size()), m = static_cast<int>(u.size());
    dynamics->Jacobian(x, u, t, jac);
    jac = MatrixXd::Identity(n, n + m) + jac * static_cast<double>(h);
  }
};

// classic RK4 and its exact discrete Jacobian by the chain rule through the four stages
// (integration.hpp:124-167 there; float h promoted to double in every product; the
// stage Jacobians take 0.5 * t as their time like the reference does, Q2)
template <int NStates, int NControls>
class RungeKutta4 final : public ExplicitIntegrator<NStates, NControls> {
  using typename ExplicitIntegrator<NStates, NControls>::DynamicsPtr;

 public:
  RungeKutta4(int n, int m) : ExplicitIntegrator<NStates, NControls>(n, m) {}
  RungeKutta4() : ExplicitIntegrator<NStates, NControls>() {}

  void Integrate(const DynamicsPtr& dynamics, const VectorXdRef& x, const VectorXdRef& u, float t, float h,
                 Eigen::Ref<VectorXd> xnext) override {
    VectorXd k[4];
    Stages(dynamics, x, u, t, h, k, true);
    xnext = x + static_cast<double>(h) * (k[0] + 2 * k[1] + 2 * k[2] + k[3]) / 6;
  }
  void Jacobian(const DynamicsPtr& dynamics, const VectorXdRef& x, const VectorXdRef& u, float t, float h,
                Eigen::Ref<MatrixXd> jac) override {
    const int n = dynamics->StateDimension(), m = dynamics->ControlDimension();
    const double hd = static_cast<double>(h);
    VectorXd k[4];
    Stages(dynamics, x, u, t, h, k, false);
    const VectorXd xs[4] = {VectorXd(x), VectorXd(x + 0.5 * k[0] * hd), VectorXd(x + 0.5 * k[1] * hd),
                            VectorXd(x + k[2] * hd)};
    const float ts[4] = {t, 0.5F * t, 0.5F * t, t};
    const double w[4] = {0.0, 0.5, 0.5, 1.0};
    MatrixXd dA[4], dB[4];
    for (int s = 0; s < 4; ++s) {
      MatrixXd J = MatrixXd::Zero(n, n + m);
      dynamics->Jacobian(xs[s], u, ts[s], J);
      const MatrixXd A = J.topLeftCorner(n, n), B = J.topRightCorner(n, m);
      if (s == 0) {
        dA[0] = A * hd;
        dB[0] = B * hd;
      } else {
        dA[s] = A * (MatrixXd::Identity(n, n) + w[s] * dA[s - 1]) * hd;
        dB[s] = B * hd + w[s] * A * dB[s - 1] * hd;
      }
    }
    jac.topLeftCorner(n, n) = MatrixXd::Identity(n, n) + (dA[0] + 2 * dA[1] + 2 * dA[2] + dA[3]) / 6;
    jac.topRightCorner(n, m) = (dB[0] + 2 * dB[1] + 2 * dB[2] + dB[3]) / 6;
  }

 private:
  static void Stages(const DynamicsPtr& f, const VectorXdRef& x, const VectorXdRef& u, float t, float h, VectorXd* k,
                     bool all) {
    const double hd = static_cast<double>(h);
    for (int s = 0; s < 4; ++s) k[s] = VectorXd::Zero(x.size());
    f->Evaluate(x, u, t, k[0]);
    f->Evaluate(x + k[0] * 0.5 * hd, u, t + 0.5F * h, k[1]);
    f->Evaluate(x + k[1] * 0.5 * hd, u, t + 0.5F * h, k[2]);
    if (all) f->Evaluate(x + k[2] * hd, u, t + h, k[3]);
  }
};

}  // namespace problem
}  // namespace altro
