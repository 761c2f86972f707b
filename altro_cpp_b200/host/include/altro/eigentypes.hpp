// altro/eigentypes.hpp (B200 host mirror) — the Eigen aliases the reference's API is spelled in
// (altro/eigentypes.hpp:8-27 there).  <eigen3/Eigen/Dense> resolves to a real Eigen when one is on
// the include path, else to the small stand-in shipped next to these headers.
#pragma once

#include <eigen3/Eigen/Dense>

namespace altro {

template <int n, class T = double>
using VectorN = Eigen::Matrix<T, n, 1>;
template <int n>
using VectorNd = Eigen::Matrix<double, n, 1>;
template <int n, int m>
using MatrixNxMd = Eigen::Matrix<double, n, m>;
using VectorXdRef = Eigen::Ref<const Eigen::VectorXd>;
template <int n, int m>
using RowMajorNxMd = Eigen::Matrix<double, n, m, Eigen::RowMajor>;
using RowMajorXd = RowMajorNxMd<Eigen::Dynamic, Eigen::Dynamic>;
using VectorXd = Eigen::VectorXd;
using VectorXf = Eigen::VectorXf;
using MatrixXd = Eigen::MatrixXd;
using MatrixXf = Eigen::MatrixXf;

}  // namespace altro
