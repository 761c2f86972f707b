// altro/utils/derivative_checker.hpp (B200 host mirror) — finite differences of host callables, the tool
// behind FunctionBase::CheckJacobian / CheckHessian and ScalarFunction::CheckGradient.  Same entry points
// and argument meaning as the reference's header of this name (FiniteDiffJacobian<nrows, ncols>(f, x, eps,
// central), FiniteDiffGradient<ncols>, FiniteDiffHessian<ncols>), so user code and the reference's own
// unit tests compile against it; `f` maps a vector to a vector (Jacobian) or to a scalar (gradient, Hessian).
//
// Host-only utilities: the device solve never differentiates numerically — models carry analytic
// Jacobians (csrc/device.cuh).
#pragma once

#include <cmath>

#include "altro/eigentypes.hpp"

namespace altro {
namespace utils {

namespace detail {
// One column of a difference quotient: (f(x + eps e_j) - f(x - eps e_j)) / 2 eps when `central`, else
// (f(x + eps e_j) - f0) / eps with f0 = f(x) computed once by the caller.
template <class Out, class Func, class Vec>
Out DifferenceAlong(const Func& f, const Vec& x, int j, double eps, bool central, const Out& f0) {
  Vec probe = x;
  probe(j) = x(j) + eps;
  Out ahead = f(probe);
  if (!central) return (ahead - f0) / eps;
  probe(j) = x(j) - eps;
  Out behind = f(probe);
  return (ahead - behind) / (2 * eps);
}
}  // namespace detail

// d f / d x of f : R^ncols -> R^nrows at x (sizes may be Eigen::Dynamic: they are then taken from x and f(x))
template <int nrows, int ncols, class Func>
Eigen::Matrix<double, nrows, ncols> FiniteDiffJacobian(const Func& f, const Eigen::Ref<const Eigen::Matrix<double, ncols, 1>>& x,
                                                       const double eps = 1e-6, const bool central = false) {
  using In = Eigen::Matrix<double, ncols, 1>;
  using Out = Eigen::Matrix<double, nrows, 1>;
  const In x0 = x;
  const Out f0 = f(x0);
  const int inputs = static_cast<int>(x0.size()), outputs = static_cast<int>(f0.size());
  Eigen::Matrix<double, nrows, ncols> jac = Eigen::Matrix<double, nrows, ncols>::Zero(outputs, inputs);
  for (int j = 0; j < inputs; ++j) {
    const Out column = detail::DifferenceAlong<Out>(f, x0, j, eps, central, f0);
    for (int i = 0; i < outputs; ++i) jac(i, j) = column(i);
  }
  return jac;
}

template <class Func>
Eigen::MatrixXd FiniteDiffJacobian(const Func& f, const VectorXdRef& x, const double eps = 1e-6, const bool central = false) {
  return FiniteDiffJacobian<Eigen::Dynamic, Eigen::Dynamic, Func>(f, x, eps, central);
}

// gradient of a scalar function f : R^ncols -> R
template <int ncols, class Func>
Eigen::Matrix<double, ncols, 1> FiniteDiffGradient(const Func& f, const Eigen::Matrix<double, ncols, 1>& x,
                                                   const double eps = 1e-6, const bool central = false) {
  using In = Eigen::Matrix<double, ncols, 1>;
  const int inputs = static_cast<int>(x.size());
  const double f0 = f(x);
  In grad = In::Zero(inputs);
  for (int j = 0; j < inputs; ++j) grad(j) = detail::DifferenceAlong<double>(f, x, j, eps, central, f0);
  return grad;
}

// Function objects kept for source compatibility with code written against the reference's header: a scalar
// function wrapped as a one-row vector function, and "the finite-difference gradient of f" as a callable.
template <class Func>
struct ScalarToVec {
  using Vector1d = Eigen::Matrix<double, 1, 1>;
  Func f;
  Vector1d operator()(const VectorXd& x) const { return Vector1d::Constant(f(x)); }
};
template <int nrows, class Func, class T>
struct FiniteDiffGradientFunc {
  using GradVec = Eigen::Matrix<T, nrows, 1>;
  Func f;
  double eps;
  bool central;
  GradVec operator()(const GradVec& x) const { return FiniteDiffGradient<nrows, Func>(f, x, eps, central); }
};

// Hessian of a scalar function: the Jacobian of its finite-difference gradient (same step for both levels)
template <int ncols, class Func>
Eigen::Matrix<double, ncols, ncols> FiniteDiffHessian(const Func& f, const Eigen::Matrix<double, ncols, 1>& x,
                                                      const double eps = 1e-4, const bool central = true) {
  using In = Eigen::Matrix<double, ncols, 1>;
  auto gradient = [&f, eps, central](const In& at) -> In { return FiniteDiffGradient<ncols, Func>(f, at, eps, central); };
  return FiniteDiffJacobian<ncols, ncols>(gradient, x, eps, central);
}

}  // namespace utils
}  // namespace altro
