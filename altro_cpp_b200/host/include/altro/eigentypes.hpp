// altro/eigentypes.hpp (B200 host mirror) — the matrix / vector spellings the reference's API is written
// in (altro/eigentypes.hpp:8-27 there).  <eigen3/Eigen/Dense> resolves to a real Eigen when one is on the
// include path, else to the small stand-in shipped next to these headers; everything that crosses the
// C ABI (include/altro_b200.h) is `.data()` of these column-major types.
#pragma once

#include <eigen3/Eigen/Dense>

namespace altro {

// run-time sized, what the virtual interfaces (FunctionBase, CostFunction, Constraint) exchange
using MatrixXd = Eigen::MatrixXd;
using VectorXd = Eigen::VectorXd;
using MatrixXf = Eigen::MatrixXf;
using VectorXf = Eigen::VectorXf;
using VectorXdRef = Eigen::Ref<const Eigen::VectorXd>;  // read-only view: a vector, a segment, a Map

// compile-time sized, what the solver templates <n, m> hold per knot point
template <int n, int m>
using MatrixNxMd = Eigen::Matrix<double, n, m>;
template <int n>
using VectorNd = Eigen::Matrix<double, n, 1>;
template <int n, class T = double>
using VectorN = Eigen::Matrix<T, n, 1>;

// row-major variants (Jacobians handed to row-wise consumers)
template <int n, int m>
using RowMajorNxMd = Eigen::Matrix<double, n, m, Eigen::RowMajor>;
using RowMajorXd = RowMajorNxMd<Eigen::Dynamic, Eigen::Dynamic>;

}  // namespace altro

// Programs that print matrices with fmt ("{}" of a VectorXd) relied on fmt < 9 picking up operator<< by itself;
// newer fmt wants the formatter named.  Only when the program has included fmt before this header.
#if defined(FMT_VERSION) && FMT_VERSION >= 90000 && defined(FMT_OSTREAM_H_)
template <class T, int R, int C, int Opt>
struct fmt::formatter<Eigen::Matrix<T, R, C, Opt>> : fmt::ostream_formatter {};
#endif
