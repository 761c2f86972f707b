// examples/quadratic_cost.hpp (B200 host mirror) — l(x,u) = 1/2 x'Qx + x'Hu + 1/2 u'Ru + q'x + r'u + c
// (reference: examples/quadratic_cost.hpp:12-75, quadratic_cost.cpp:8-28).  Value, gradient and
// Hessian are evaluated on the device (device.cuh quad_eval, quad_gradient, knot_expansion).
#pragma once

#include <stdexcept>

#include "altro/problem/costfunction.hpp"

namespace altro {
namespace examples {

class QuadraticCost : public problem::CostFunction {
 public:
  QuadraticCost(const MatrixXd& Q, const MatrixXd& R, const MatrixXd& H, const VectorXd& q, const VectorXd& r,
                double c = 0, bool terminal = false)
      : n_(q.size()), m_(r.size()), Q_(Q), R_(R), H_(H), q_(q), r_(r), c_(c), terminal_(terminal) {
    if (Q_.rows() != n_ || Q_.cols() != n_) throw std::invalid_argument("Q has the wrong size");
    if (R_.rows() != m_ || R_.cols() != m_) throw std::invalid_argument("R has the wrong size");
    if (H_.rows() != n_ || H_.cols() != m_) throw std::invalid_argument("H has the wrong size");
  }

  // quadratic_cost.hpp:29-39
  static QuadraticCost LQRCost(const MatrixXd& Q, const MatrixXd& R, const VectorXd& xref, const VectorXd& uref,
                               bool terminal = false) {
    const int n = Q.rows(), m = R.rows();
    MatrixXd H = MatrixXd::Zero(n, m);
    VectorXd q = -(Q * xref);
    VectorXd r = -(R * uref);
    const double c = 0.5 * dot(xref, Q * xref) + 0.5 * dot(uref, R * uref);
    return QuadraticCost(Q, R, H, q, r, c, terminal);
  }

  int StateDimension() const override { return n_; }
  int ControlDimension() const override { return m_; }
  const MatrixXd& GetQ() const { return Q_; }
  const MatrixXd& GetR() const { return R_; }
  const MatrixXd& GetH() const { return H_; }
  const VectorXd& Getq() const { return q_; }
  const VectorXd& Getr() const { return r_; }
  double GetConstant() const { return c_; }

  bool Describe(device::CostDesc* d) const override {
    d->Q.assign(Q_.data(), Q_.data() + n_ * n_);
    d->R.assign(R_.data(), R_.data() + m_ * m_);
    d->H.assign(H_.data(), H_.data() + n_ * m_);
    d->q.assign(q_.data(), q_.data() + n_);
    d->r.assign(r_.data(), r_.data() + m_);
    d->c = c_;
    return true;
  }

 private:
  int n_, m_;
  MatrixXd Q_, R_, H_;
  VectorXd q_, r_;
  double c_;
  bool terminal_;
};

}  // namespace examples
}  // namespace altro
