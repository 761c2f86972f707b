// altro/common/functionbase.hpp (B200 host mirror) — the plug-in ABCs every dynamics, cost and
// constraint functor derives from (altro/common/functionbase.hpp:57,137 there): same virtuals, same
// Eigen::Ref argument types, column-major double.
//
// Where these run: a user's functor is ordinary host code.  The solvers never call it inside a
// solve — virtual calls cannot run on a GPU.  They call it a handful of times when a
// solver is built, to recognise the functor as one of the device-capable kinds and to read its
// parameters (altro/device_registry.hpp); everything numerical then happens on the device.
#pragma once

#include <type_traits>

#include "altro/eigentypes.hpp"
#include "altro/utils/derivative_checker.hpp"
#include "altro/utils/utils.hpp"

namespace altro {

class FunctionBase {
 public:
  virtual ~FunctionBase() = default;

  static constexpr int NStates = Eigen::Dynamic;
  static constexpr int NControls = Eigen::Dynamic;
  static constexpr int NOutputs = Eigen::Dynamic;

  virtual int StateDimension() const { return 0; }
  virtual int ControlDimension() const { return 0; }
  virtual int OutputDimension() const = 0;

  virtual void Evaluate(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> out) = 0;
  virtual void Jacobian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> jac) = 0;
  virtual void Hessian(const VectorXdRef& x, const VectorXdRef& u, const VectorXdRef& b, Eigen::Ref<MatrixXd> hess) {
    ALTRO_UNUSED(x);
    ALTRO_UNUSED(u);
    ALTRO_UNUSED(b);
    ALTRO_UNUSED(hess);
  }
  virtual bool HasHessian() const = 0;

  // finite-difference check of the user's Jacobian at (x, u)
  bool CheckJacobian(const VectorXdRef& x, const VectorXdRef& u, double eps = kDefaultTolerance, bool verbose = false) {
    const int n = static_cast<int>(x.size()), m = static_cast<int>(u.size()), p = OutputDimension();
    VectorXd z(n + m);
    for (int i = 0; i < n; ++i) z(i) = x(i);
    for (int j = 0; j < m; ++j) z(n + j) = u(j);
    auto f = [&](const VectorXd& zz, VectorXd& out) {
      VectorXd xx(n), uu(m);
      for (int i = 0; i < n; ++i) xx(i) = zz(i);
      for (int j = 0; j < m; ++j) uu(j) = zz(n + j);
      Evaluate(xx, uu, out);
    };
    const MatrixXd fd = utils::FiniteDiffJacobian(f, z, p);
    MatrixXd jac = MatrixXd::Zero(p, n + m);
    Jacobian(x, u, jac);
    const double err = (fd - jac).template lpNorm<Eigen::Infinity>();
    if (verbose) std::fprintf(stderr, "CheckJacobian: max |fd - jac| = %g\n", err);
    return err < eps;
  }
  bool CheckJacobian(double eps = kDefaultTolerance, bool verbose = false) {
    return CheckJacobian(VectorXd::Random(StateDimension()), VectorXd::Random(ControlDimension()), eps, verbose);
  }

 protected:
  static constexpr double kDefaultTolerance = 1e-4;
};

class ScalarFunction : public FunctionBase {
 public:
  static const int NOutputs = 1;
  int OutputDimension() const override { return 1; }

  virtual double Evaluate(const VectorXdRef& x, const VectorXdRef& u) = 0;
  virtual void Gradient(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> grad) = 0;
  virtual void Hessian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> hess) = 0;

  // the vector-valued interface expressed through the scalar one
  void Evaluate(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> out) override {
    ALTRO_ASSERT(out.size() == 1, "Output must be of size 1 for scalar functions");
    out(0) = Evaluate(x, u);
  }
  void Jacobian(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<MatrixXd> jac) override {
    ALTRO_ASSERT(jac.rows() == 1, "Jacobian of a scalar function must have a single row.");
    Eigen::Map<VectorXd> grad(jac.data(), jac.cols());  // a 1 x N row is N contiguous doubles
    Gradient(x, u, grad);
  }
  void Hessian(const VectorXdRef& x, const VectorXdRef& u, const VectorXdRef& b, Eigen::Ref<MatrixXd> hess) override {
    ALTRO_ASSERT(b.size() == 1 && b.isApproxToConstant(1), "The b vector for scalar Hessians must be a vector of a single 1.");
    ALTRO_UNUSED(b);
    Hessian(x, u, hess);
  }
  bool HasHessian() const override { return true; }
};

}  // namespace altro
