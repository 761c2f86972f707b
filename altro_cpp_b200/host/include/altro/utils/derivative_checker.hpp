// altro/utils/derivative_checker.hpp (B200 host mirror) — central finite differences of a host
// callable, the tool behind FunctionBase::CheckJacobian (altro/utils/derivative_checker.hpp there).
#pragma once

#include <cmath>

#include "altro/eigentypes.hpp"

namespace altro {
namespace utils {

// Jacobian of f : R^nin -> R^nout at z (f(z, out) fills out), step eps, central differences
template <class Func>
MatrixXd FiniteDiffJacobian(const Func& f, const VectorXd& z, int nout, double eps = 1e-6) {
  const int nin = static_cast<int>(z.size());
  MatrixXd J = MatrixXd::Zero(nout, nin);
  VectorXd zp = z, zm = z, fp = VectorXd::Zero(nout), fm = VectorXd::Zero(nout);
  for (int j = 0; j < nin; ++j) {
    zp(j) = z(j) + eps;
    zm(j) = z(j) - eps;
    f(zp, fp);
    f(zm, fm);
    for (int i = 0; i < nout; ++i) J(i, j) = (fp(i) - fm(i)) / (2 * eps);
    zp(j) = z(j);
    zm(j) = z(j);
  }
  return J;
}

}  // namespace utils
}  // namespace altro
