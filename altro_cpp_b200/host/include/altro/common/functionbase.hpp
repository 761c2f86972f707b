// altro/common/functionbase.hpp (B200 host mirror) — the plug-in ABCs every dynamics, cost and
// constraint functor derives from (altro/common/functionbase.hpp:57,137 there): same virtuals, same
// Eigen::Ref argument types, column-major double.
//
// Where these run: a user's functor is ordinary host code.  The solvers never call it inside a
// solve — virtual calls cannot run on a GPU.  They call it a handful of times when a
// solver is built, to recognise the functor as one of the device-capable kinds and to read its
// parameters (altro/device_registry.hpp); everything numerical then happens on the device.
#pragma once

#include <cstdio>
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

  // ---- finite-difference checks of the user's derivatives (host-side debugging aids; tolerance on the
  // Frobenius norm of the difference).  Without arguments: at a random point of the functor's own sizes.
  bool CheckJacobian(const VectorXdRef& x, const VectorXdRef& u, double eps = kDefaultTolerance, bool verbose = false) {
    const int n = static_cast<int>(x.size()), m = static_cast<int>(u.size()), p = OutputDimension();
    MatrixXd analytic = MatrixXd::Zero(p, n + m);
    Jacobian(x, u, analytic);
    auto stacked = [this, n, m, p](const VectorXd& z) -> VectorXd {
      VectorXd out = VectorXd::Zero(p);
      this->Evaluate(z.head(n), z.tail(m), out);
      return out;
    };
    const MatrixXd numeric = utils::FiniteDiffJacobian<Eigen::Dynamic, Eigen::Dynamic>(stacked, Stack(x, u));
    return Agree(numeric, analytic, eps, verbose, "Jacobian");
  }
  bool CheckJacobian(double eps = kDefaultTolerance, bool verbose = false) {
    return CheckJacobian(VectorXd::Random(StateDimension()), VectorXd::Random(ControlDimension()), eps, verbose);
  }
  // Hessian of b' f(x, u) with respect to (x, u)
  bool CheckHessian(const VectorXdRef& x, const VectorXdRef& u, const VectorXdRef& b, double eps = kDefaultTolerance,
                    bool verbose = false) {
    const int n = StateDimension(), m = ControlDimension(), p = OutputDimension();
    MatrixXd analytic = MatrixXd::Zero(n + m, n + m);
    Hessian(x, u, b, analytic);
    const VectorXd weights = b;
    auto weighted = [this, n, m, p, &weights](const VectorXd& z) -> double {
      VectorXd out = VectorXd::Zero(p);
      this->Evaluate(z.head(n), z.tail(m), out);
      return out.dot(weights);
    };
    const MatrixXd numeric = utils::FiniteDiffHessian<Eigen::Dynamic>(weighted, Stack(x, u));
    return Agree(numeric, analytic, eps, verbose, "Hessian");
  }
  bool CheckHessian(double eps = kDefaultTolerance, bool verbose = false) {
    const int p = OutputDimension();
    VectorXd b = (p == 1) ? VectorXd::Ones(p) : VectorXd::Random(p);
    return CheckHessian(VectorXd::Random(StateDimension()), VectorXd::Random(ControlDimension()), b, eps, verbose);
  }

 protected:
  static VectorXd Stack(const VectorXdRef& x, const VectorXdRef& u) {
    VectorXd z(x.size() + u.size());
    for (int i = 0; i < x.size(); ++i) z(i) = x(i);
    for (int j = 0; j < u.size(); ++j) z(x.size() + j) = u(j);
    return z;
  }
  template <class A, class B>
  static bool Agree(const A& numeric, const B& analytic, double eps, bool verbose, const char* what) {
    const double err = (numeric - analytic).norm();
    if (verbose) std::fprintf(stderr, "Check%s: |finite difference - analytic| = %g (tolerance %g)\n", what, err, eps);
    return err < eps;
  }

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

  bool CheckGradient(const VectorXdRef& x, const VectorXdRef& u, double eps = kDefaultTolerance, bool verbose = false) {
    const int n = static_cast<int>(x.size()), m = static_cast<int>(u.size());
    VectorXd analytic = VectorXd::Zero(n + m);
    Gradient(x, u, analytic);
    auto stacked = [this, n, m](const VectorXd& z) -> double { return this->Evaluate(z.head(n), z.tail(m)); };
    const VectorXd numeric = utils::FiniteDiffGradient<Eigen::Dynamic>(stacked, Stack(x, u));
    return Agree(numeric, analytic, eps, verbose, "Gradient");
  }
  bool CheckGradient(double eps = kDefaultTolerance, bool verbose = false) {
    return CheckGradient(VectorXd::Random(StateDimension()), VectorXd::Random(ControlDimension()), eps, verbose);
  }
};

}  // namespace altro
