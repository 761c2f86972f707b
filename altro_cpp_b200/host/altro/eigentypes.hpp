// altro/eigentypes.hpp (B200 host mirror)
//
// The reference's public API is typed with Eigen (altro/eigentypes.hpp:8-27 there).  Eigen is not
// available in this image, so this mirror ships the few dense types the problem-definition code
// of examples/ and perf/ needs: dynamically sized column vectors and column-major matrices with
// the Eigen spellings used on that path (operator(), <<-comma initialisation, setZero,
// setConstant, Zero/Constant/Identity, diagonal().setConstant, scalar products).  No numerical
// work of the solver happens in these types: they only carry problem data to the device.
// With a real Eigen installed, INTEGRATION.md shows the shim that uses it instead.
#pragma once

#include <cstddef>
#include <initializer_list>
#include <stdexcept>
#include <vector>

namespace altro {

class VectorXd {
 public:
  VectorXd() = default;
  explicit VectorXd(int n) : d_(n, 0.0) {}
  VectorXd(std::initializer_list<double> v) : d_(v) {}
  static VectorXd Zero(int n) { return VectorXd(n); }
  static VectorXd Constant(int n, double v) {
    VectorXd r(n);
    r.setConstant(v);
    return r;
  }
  int size() const { return static_cast<int>(d_.size()); }
  int rows() const { return size(); }
  double& operator()(int i) { return d_.at(i); }
  double operator()(int i) const { return d_.at(i); }
  double& operator[](int i) { return d_.at(i); }
  double operator[](int i) const { return d_.at(i); }
  void setZero() { setConstant(0.0); }
  void setZero(int n) { d_.assign(n, 0.0); }
  void setConstant(double v) {
    for (double& x : d_) x = v;
  }
  const double* data() const { return d_.data(); }
  double* data() { return d_.data(); }
  VectorXd& operator*=(double s) {
    for (double& x : d_) x *= s;
    return *this;
  }
  VectorXd operator-() const {
    VectorXd r(*this);
    r *= -1.0;
    return r;
  }
  // `v << a, b, c;`
  struct CommaInit {
    VectorXd* v;
    int i;
    CommaInit& operator,(double x) {
      (*v)(i++) = x;
      return *this;
    }
  };
  CommaInit operator<<(double x) {
    (*this)(0) = x;
    return CommaInit{this, 1};
  }

 private:
  std::vector<double> d_;
};

class MatrixXd {
 public:
  MatrixXd() = default;
  MatrixXd(int r, int c) : r_(r), c_(c), d_(static_cast<size_t>(r) * c, 0.0) {}
  static MatrixXd Zero(int r, int c) { return MatrixXd(r, c); }
  static MatrixXd Identity(int r, int c) {
    MatrixXd m(r, c);
    for (int i = 0; i < r && i < c; ++i) m(i, i) = 1.0;
    return m;
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  double& operator()(int i, int j) { return d_.at(static_cast<size_t>(i) + static_cast<size_t>(j) * r_); }
  double operator()(int i, int j) const { return d_.at(static_cast<size_t>(i) + static_cast<size_t>(j) * r_); }
  const double* data() const { return d_.data(); }  // column-major, like Eigen
  void setZero() {
    for (double& x : d_) x = 0.0;
  }
  MatrixXd operator*(double s) const {
    MatrixXd r(*this);
    for (double& x : r.d_) x *= s;
    return r;
  }
  VectorXd operator*(const VectorXd& v) const {
    if (v.size() != c_) throw std::invalid_argument("MatrixXd * VectorXd: size mismatch");
    VectorXd r(r_);
    for (int i = 0; i < r_; ++i) {
      double acc = 0.0;
      for (int j = 0; j < c_; ++j) acc += (*this)(i, j) * v(j);
      r(i) = acc;
    }
    return r;
  }
  // `Q.diagonal().setConstant(v)`
  struct DiagonalProxy {
    MatrixXd* m;
    void setConstant(double v) {
      for (int i = 0; i < m->rows() && i < m->cols(); ++i) (*m)(i, i) = v;
    }
  };
  DiagonalProxy diagonal() { return DiagonalProxy{this}; }

 private:
  int r_ = 0, c_ = 0;
  std::vector<double> d_;
};

inline double dot(const VectorXd& a, const VectorXd& b) {
  double s = 0.0;
  for (int i = 0; i < a.size(); ++i) s += a(i) * b(i);
  return s;
}

using VectorXdRef = const VectorXd&;

}  // namespace altro
