// examples/quadratic_cost.hpp (shipped with the B200 host mirror) —
// l(x,u) = 1/2 x'Qx + x'Hu + 1/2 u'Ru + q'x + r'u + c (examples/quadratic_cost.hpp:12-75 there).
// In a solve, value, gradient and Hessian are evaluated on the device (csrc/device.cuh quad_eval,
// quad_gradient, knot_expansion); the virtuals below are the host plug-in interface.
#pragma once

#include <stdexcept>

#include "altro/device_registry.hpp"
#include "altro/problem/costfunction.hpp"

namespace altro {
namespace examples {

class QuadraticCost : public problem::CostFunction, public device::Describable {
 public:
  QuadraticCost(const MatrixXd& Q, const MatrixXd& R, const MatrixXd& H, const VectorXd& q, const VectorXd& r,
                double c = 0, bool terminal = false)
      : n_(static_cast<int>(q.size())), m_(static_cast<int>(r.size())), Q_(Q), R_(R), H_(H), q_(q), r_(r), c_(c),
        terminal_(terminal) {
    if (Q_.rows() != n_ || Q_.cols() != n_) throw std::invalid_argument("Q has the wrong size");
    if (R_.rows() != m_ || R_.cols() != m_) throw std::invalid_argument("R has the wrong size");
    if (H_.rows() != n_ || H_.cols() != m_) throw std::invalid_argument("H has the wrong size");
    if (!Q_.isApprox(Q_.transpose()) || !R_.isApprox(R_.transpose())) throw std::invalid_argument("Q and R must be symmetric");
    if (!terminal_ && Eigen::LLT<MatrixXd>(R_).info() != Eigen::Success) throw std::invalid_argument("R must be positive definite");
  }

  // tracking cost 1/2 (x - xref)'Q(x - xref) + 1/2 (u - uref)'R(u - uref) in expanded form
  static QuadraticCost LQRCost(const MatrixXd& Q, const MatrixXd& R, const VectorXd& xref, const VectorXd& uref,
                               bool terminal = false) {
    const VectorXd Qx = Q * xref, Ru = R * uref;
    return QuadraticCost(Q, R, MatrixXd::Zero(Q.rows(), R.rows()), -Qx, -Ru, 0.5 * xref.dot(Qx) + 0.5 * uref.dot(Ru),
                         terminal);
  }

  int StateDimension() const override { return n_; }
  int ControlDimension() const override { return m_; }
  double Evaluate(const VectorXdRef& x, const VectorXdRef& u) override {
    return 0.5 * x.dot(Q_ * x) + x.dot(H_ * u) + 0.5 * u.dot(R_ * u) + q_.dot(x) + r_.dot(u) + c_;
  }
  void Gradient(const VectorXdRef& x, const VectorXdRef& u, Eigen::Ref<VectorXd> dx, Eigen::Ref<VectorXd> du) override {
    dx = Q_ * x + q_ + H_ * u;
    du = R_ * u + r_ + H_.transpose() * x;
  }
  void Hessian(const VectorXdRef&, const VectorXdRef&, Eigen::Ref<MatrixXd> dxdx, Eigen::Ref<MatrixXd> dxdu,
               Eigen::Ref<MatrixXd> dudu) override {
    dxdx = Q_;
    dxdu = H_;
    dudu = R_;
  }
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
