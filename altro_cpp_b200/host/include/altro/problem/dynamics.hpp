// altro/problem/dynamics.hpp (B200 host mirror) — ContinuousDynamics and DiscreteDynamics
// (altro/problem/dynamics.hpp:59,148 there): the time-aware virtuals plus the FunctionBase ones
// forwarding to them with the stored time (and step).
#pragma once

#include <iostream>

#include "altro/common/functionbase.hpp"
#include "altro/eigentypes.hpp"
#include "altro/utils/utils.hpp"

namespace altro {
namespace problem {

class ContinuousDynamics : public FunctionBase {
 public:
  using FunctionBase::Evaluate;
  using FunctionBase::Hessian;
  using FunctionBase::Jacobian;

  int OutputDimension() const override { return StateDimension(); }

  virtual void Evaluate(const VectorXdRef& x, const VectorXdRef& u, float t, Eigen::Ref<VectorXd> xdot) = 0;
  virtual void Jacobian(const VectorXdRef& x, const VectorXdRef& u, float t, Eigen::Ref<MatrixXd> jac) = 0;
  virtual void Hessian(const VectorXdRef& x, const VectorXdRef& u, float t, const VectorXdRef& b,
                       Eigen::Ref<MatrixXd> hess) = 0;

  VectorXd Evaluate(const VectorXdRef& x, const VectorXdRef& u, float t) {
    VectorXd xdot = VectorXd::Zero(x.size());
    Evaluate(x, u, t, xdot);
    return xdot;
  }
  VectorXd operator()(const VectorXdRef& x, const VectorXdRef& u, float t) { return Evaluate(x, u, t); }

  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> out) override { Evaluate(x, u, GetTime(), out); }
  void Jacobian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> jac) override { Jacobian(x, u, GetTime(), jac); }
  void Hessian(const VectorXdRef& x, const VectorXdRef& u, const VectorXdRef& b, Eigen::Ref<MatrixXd> hess) override {
    Hessian(x, u, GetTime(), b, hess);
  }

  float GetTime() const { return t_; }
  void SetTime(float t) { t_ = t; }

 protected:
  float t_ = 0.0F;
};

class DiscreteDynamics : public FunctionBase {
 public:
  using FunctionBase::Evaluate;
  using FunctionBase::Hessian;
  using FunctionBase::Jacobian;

  int OutputDimension() const override { return StateDimension(); }

  virtual void Evaluate(const VectorXdRef& x, const VectorXdRef& u, float t, float h, Eigen::Ref<VectorXd> xnext) = 0;
  virtual void Jacobian(const VectorXdRef& x, const VectorXdRef& u, float t, float h, Eigen::Ref<MatrixXd> jac) = 0;
  virtual void Hessian(const VectorXdRef& x, const VectorXdRef& u, float t, float h, const VectorXdRef& b,
                       Eigen::Ref<MatrixXd> hess) = 0;

  VectorXd Evaluate(const VectorXdRef& x, const VectorXdRef& u, float t, float h) {
    VectorXd xnext = VectorXd::Zero(x.size());
    Evaluate(x, u, t, h, xnext);
    return xnext;
  }
  VectorXd operator()(const VectorXdRef& x, const VectorXdRef& u, float t, float h) { return Evaluate(x, u, t, h); }

  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> out) override {
    Evaluate(x, u, GetTime(), GetStep(), out);
  }
  void Jacobian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> jac) override {
    Jacobian(x, u, GetTime(), GetStep(), jac);
  }
  void Hessian(const VectorXdRef& x, const VectorXdRef& u, const VectorXdRef& b, Eigen::Ref<MatrixXd> hess) override {
    Hessian(x, u, GetTime(), GetStep(), b, hess);
  }

  float GetTime() const { return t_; }
  void SetTime(float t) { t_ = t; }
  float GetStep() const { return h_; }
  void SetStep(float h) { h_ = h; }

 protected:
  float t_ = 0.0F;
  float h_ = 0.0F;
};

}  // namespace problem
}  // namespace altro
