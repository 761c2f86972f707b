// oracle/ref_shim/compat.hpp — force-included (-include) in front of every translation unit of the reference when
// oracle/build_ref.py compiles it where it lies (/root/reference).  Test infrastructure, not part of the product.
//
// The reference was written against fmt < 9, which formatted any type with an operator<< on its own; the fmt that
// ships with this image wants the formatter named.  Nothing of the reference is changed: these are the
// specialisations its `fmt::print("{}", vector)` calls need.
#pragma once

#include <fmt/format.h>
#include <fmt/ostream.h>

#include <eigen3/Eigen/Dense>

template <class T, int R, int C, int Opt>
struct fmt::formatter<Eigen::Matrix<T, R, C, Opt>> : fmt::ostream_formatter {};
template <class Plain>
struct fmt::formatter<Eigen::Ref<Plain>> : fmt::ostream_formatter {};
template <class D>
struct fmt::formatter<Eigen::WithFormat<D>> : fmt::ostream_formatter {};
