// altro/device_registry.hpp — how a user-level functor becomes something the GPU can evaluate.
//
// The reference's extension mechanism is subclassing with virtual Evaluate / Jacobian calls; a
// virtual call cannot run on the device.  The device library has a closed set of
// functor kinds (include/altro_b200.h): RK4-discretised registered continuous models, quadratic
// costs, and goal / control-bound / circle constraints.  A functor joins in one of two ways:
//
//  1. it says what it is: it also derives from device::Describable and fills a small POD
//     (examples shipped with this mirror do that);
//  2. it is RECOGNISED: the solver evaluates the functor's own virtuals at a handful of probe points
//     when it is built, reads the parameters off the answers (a quadratic cost is its Hessian,
//     its gradient at 0 and its value at 0; x - xf at x = 0 gives -xf; ...), and then CHECKS the
//     recovered description against the functor at random points.  The reference's unmodified
//     example classes (examples/quadratic_cost.hpp, basic_constraints.hpp, obstacle_constraints.hpp,
//     unicycle.hpp, triple_integrator.hpp) have no accessors for their parameters and take this road.
//
// A functor that neither describes itself nor passes recognition makes solver construction throw
// altro::DeviceError(ALTRO_B200_ERR_UNSUPPORTED) — there is no host fallback for the solve.
#pragma once

#include <cmath>
#include <cstdlib>
#include <limits>
#include <memory>
#include <string>
#include <typeinfo>
#include <vector>

#include "altro/constraints/constraint.hpp"
#include "altro/problem/costfunction.hpp"
#include "altro/problem/discretized_model.hpp"
#include "altro/problem/dynamics.hpp"
#include "altro_b200.h"

namespace altro {
namespace device {

struct ModelDesc {
  int model = -1;  // altro_b200_model
  std::vector<double> params;
};
struct CostDesc {
  std::vector<double> Q, R, H, q, r;  // column-major
  double c = 0.0;
};
struct ConstraintDesc {
  enum Kind { kNone = -1, kGoal = 0, kControlBound = 1, kCircle = 2 } kind = kNone;
  std::vector<double> a, b, c;  // goal: a = xf | bound: a = lb, b = ub (+-inf = absent) | circle: cx, cy, r^2
  int xi = 0, yi = 1;
};

// way 1: self-description
class Describable {
 public:
  virtual ~Describable() = default;
  virtual bool Describe(ModelDesc*) const { return false; }
  virtual bool Describe(CostDesc*) const { return false; }
  virtual bool Describe(ConstraintDesc*) const { return false; }
};

namespace detail {

inline VectorXd Probe(int n, unsigned seed) {  // deterministic, irrational-looking probe points
  VectorXd v = VectorXd::Zero(n);
  unsigned long long s = 0x9E3779B97F4A7C15ull * (seed + 1);
  for (int i = 0; i < n; ++i) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    v(i) = static_cast<double>(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
  }
  return v;
}
inline bool Close(double a, double b, double tol) { return std::fabs(a - b) <= tol * (1.0 + std::fabs(a) + std::fabs(b)); }

}  // namespace detail

// ---- dynamics ----------------------------------------------------------------------------------
// A registered continuous model is recognised by its behaviour: xdot and its Jacobian at probe points
// must equal the registered model's.  (No parameters to recover for the unicycle / triple integrator.)
inline bool RecogniseContinuous(problem::ContinuousDynamics& f, ModelDesc* out) {
  const int n = f.StateDimension(), m = f.ControlDimension();
  bool unicycle = (n == 3 && m == 2), triple = (m >= 1 && n == 3 * m);
  for (unsigned s = 0; s < 3 && (unicycle || triple); ++s) {
    const VectorXd x = detail::Probe(n, 10 + s), u = detail::Probe(m, 20 + s);
    VectorXd xd = VectorXd::Zero(n);
    f.Evaluate(x, u, 0.0F, xd);
    if (unicycle) {  // [v cos(theta), v sin(theta), omega]
      unicycle = detail::Close(xd(0), u(0) * std::cos(x(2)), 1e-14) && detail::Close(xd(1), u(0) * std::sin(x(2)), 1e-14) &&
                 xd(2) == u(1);
    }
    if (triple) {  // chain of integrators driven by the jerk
      for (int i = 0; i < m; ++i) triple = triple && xd(i) == x(i + m) && xd(i + m) == x(i + 2 * m) && xd(i + 2 * m) == u(i);
    }
  }
  if (unicycle) {
    out->model = ALTRO_B200_MODEL_UNICYCLE;
    return true;
  }
  if (triple) {
    out->model = ALTRO_B200_MODEL_TRIPLE_INTEGRATOR;
    return true;
  }
  return false;
}

inline bool DescribeDynamics(problem::DiscreteDynamics& dyn, ModelDesc* out, std::string* why) {
  if (const auto* d = dynamic_cast<const Describable*>(&dyn))
    if (d->Describe(out)) return true;
  const auto* disc = dynamic_cast<const problem::DiscretizedModelBase*>(&dyn);
  if (disc == nullptr) {
    *why = std::string("dynamics of type ") + typeid(dyn).name() + " is not a DiscretizedModel of a registered continuous model";
    return false;
  }
  if (!disc->IsRungeKutta4()) {
    *why = "only RungeKutta4-discretised models run on the device";
    return false;
  }
  std::shared_ptr<problem::ContinuousDynamics> f = disc->GetContinuousModel();
  if (const auto* d = dynamic_cast<const Describable*>(f.get()))
    if (d->Describe(out)) return true;
  if (RecogniseContinuous(*f, out)) return true;
  *why = std::string("continuous model of type ") + typeid(*f).name() + " is not in the device registry";
  return false;
}

// ---- cost --------------------------------------------------------------------------------------
// Any cost that IS a quadratic form is device-capable: Q, R, H from the Hessian, q, r from the
// gradient at the origin, c from the value at the origin; then l(x,u) and its gradient must match the
// recovered form at two probe points.
inline bool DescribeCost(problem::CostFunction& cost, int n, int m, CostDesc* out, std::string* why) {
  if (const auto* d = dynamic_cast<const Describable*>(&cost))
    if (d->Describe(out)) return true;
  const VectorXd x0 = VectorXd::Zero(n), u0 = VectorXd::Zero(m);
  MatrixXd Q = MatrixXd::Zero(n, n), H = MatrixXd::Zero(n, m), R = MatrixXd::Zero(m, m);
  VectorXd q = VectorXd::Zero(n), r = VectorXd::Zero(m);
  cost.Hessian(x0, u0, Q, H, R);
  cost.Gradient(x0, u0, q, r);
  const double c = cost.Evaluate(x0, u0);
  for (unsigned s = 0; s < 2; ++s) {
    const VectorXd x = detail::Probe(n, 30 + s), u = detail::Probe(m, 40 + s);
    MatrixXd Q2 = MatrixXd::Zero(n, n), H2 = MatrixXd::Zero(n, m), R2 = MatrixXd::Zero(m, m);
    cost.Hessian(x, u, Q2, H2, R2);
    const double want = 0.5 * x.dot(Q * x) + x.dot(H * u) + 0.5 * u.dot(R * u) + q.dot(x) + r.dot(u) + c;
    if (!(Q2 - Q).isApproxToConstant(0.0, 0.0) || !(R2 - R).isApproxToConstant(0.0, 0.0) ||
        !(H2 - H).isApproxToConstant(0.0, 0.0) || !detail::Close(cost.Evaluate(x, u), want, 1e-12)) {
      *why = std::string("cost function of type ") + typeid(cost).name() + " is not a quadratic form";
      return false;
    }
  }
  out->Q.assign(Q.data(), Q.data() + n * n);
  out->R.assign(R.data(), R.data() + m * m);
  out->H.assign(H.data(), H.data() + n * m);
  out->q.assign(q.data(), q.data() + n);
  out->r.assign(r.data(), r.data() + m);
  out->c = c;
  return true;
}

// ---- constraints ---------------------------------------------------------------------------------
template <class ConType>
bool DescribeConstraint(constraints::Constraint<ConType>& con, int n, int m, ConstraintDesc* out, std::string* why) {
  if (const auto* d = dynamic_cast<const Describable*>(&con))
    if (d->Describe(out)) return true;
  const int p = con.OutputDimension();
  const bool equality = std::is_same<ConType, constraints::Equality>::value;
  const VectorXd x0 = VectorXd::Zero(n), u0 = VectorXd::Zero(m);
  VectorXd c0 = VectorXd::Zero(p);
  MatrixXd J0 = MatrixXd::Zero(p, n + m);
  con.Evaluate(x0, u0, c0);
  con.Jacobian(x0, u0, J0);
  const VectorXd x1 = detail::Probe(n, 50), u1 = detail::Probe(m, 60);
  VectorXd c1 = VectorXd::Zero(p);
  MatrixXd J1 = MatrixXd::Zero(p, n + m);
  con.Evaluate(x1, u1, c1);
  con.Jacobian(x1, u1, J1);
  auto fail = [&]() {
    *why = "constraint '" + con.GetLabel() + "' (" + typeid(con).name() + ") is not a goal, control-bound or circle constraint";
    return false;
  };
  const bool affine = (J1 - J0).isApproxToConstant(0.0, 0.0);
  if (affine && equality && p == n) {  // c = x - xf ?
    ConstraintDesc d;
    d.kind = ConstraintDesc::kGoal;
    for (int i = 0; i < n; ++i) d.a.push_back(-c0(i));  // 0 - xf is exact
    for (int i = 0; i < n; ++i)
      if (c1(i) != x1(i) - d.a[i]) return fail();
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < n + m; ++j)
        if (J0(i, j) != ((i == j) ? 1.0 : 0.0)) return fail();
    *out = d;
    return true;
  }
  if (affine && !equality) {  // rows lb_j - u_j (lower, first) then u_j - ub_j (upper) ?
    const double inf = std::numeric_limits<double>::infinity();
    ConstraintDesc d;
    d.kind = ConstraintDesc::kControlBound;
    d.a.assign(m, -inf);
    d.b.assign(m, +inf);
    bool upper_seen = false;
    int last_lo = -1, last_up = -1;
    for (int i = 0; i < p; ++i) {
      int col = -1;
      for (int j = 0; j < n + m; ++j) {
        if (J0(i, j) == 0.0) continue;
        if (col >= 0 || j < n || std::fabs(J0(i, j)) != 1.0) return fail();
        col = j - n;
      }
      if (col < 0) return fail();
      if (J0(i, n + col) < 0) {  // lower bound row: c = lb - u
        if (upper_seen || col <= last_lo) return fail();
        d.a[col] = c0(i);
        last_lo = col;
      } else {  // upper bound row: c = u - ub
        if (col <= last_up) return fail();
        upper_seen = true;
        d.b[col] = -c0(i);
        last_up = col;
      }
    }
    VectorXd want = VectorXd::Zero(p);
    int row = 0;
    for (int j = 0; j < m; ++j)
      if (std::fabs(d.a[j]) < std::numeric_limits<double>::max()) want(row++) = d.a[j] - u1(j);
    for (int j = 0; j < m; ++j)
      if (std::fabs(d.b[j]) < std::numeric_limits<double>::max()) want(row++) = u1(j) - d.b[j];
    if (row != p) return fail();
    for (int i = 0; i < p; ++i)
      if (c1(i) != want(i)) return fail();
    *out = d;
    return true;
  }
  if (!affine && !equality && p >= 1 && n >= 2) {  // circles: c_i = r_i^2 - (px - cx_i)^2 - (py - cy_i)^2 ?
    // the reference writes d c_i / d(px, py) into columns 0 and 1 whatever the position indices are
    // (examples/obstacle_constraints.hpp:117-118): centres from the Jacobian at the origin
    ConstraintDesc d;
    d.kind = ConstraintDesc::kCircle;
    for (int i = 0; i < p; ++i) {
      d.a.push_back(J0(i, 0) / 2);
      d.b.push_back(J0(i, 1) / 2);
      for (int j = 2; j < n + m; ++j)
        if (J0(i, j) != 0.0) return fail();
    }
    // which states are the position: the value must not change when any other state or control does
    std::vector<int> pos;
    for (int j = 0; j < n; ++j) {
      VectorXd xe = VectorXd::Zero(n), ce = VectorXd::Zero(p);
      xe(j) = 0.5;
      con.Evaluate(xe, u0, ce);
      if (!(ce - c0).isApproxToConstant(0.0, 0.0)) pos.push_back(j);
    }
    if (pos.size() != 2) return fail();
    // order of the two indices: moving state xi moves Jacobian column 0
    MatrixXd Je = MatrixXd::Zero(p, n + m);
    VectorXd xe = VectorXd::Zero(n);
    xe(pos[0]) = 0.5;
    con.Jacobian(xe, u0, Je);
    const bool first_is_x = Je(0, 0) != J0(0, 0);
    d.xi = first_is_x ? pos[0] : pos[1];
    d.yi = first_is_x ? pos[1] : pos[0];
    // radius^2: the value at the centre (both squared distances are exactly zero there)
    for (int i = 0; i < p; ++i) {
      VectorXd xc = VectorXd::Zero(n), cc = VectorXd::Zero(p);
      xc(d.xi) = d.a[i];
      xc(d.yi) = d.b[i];
      con.Evaluate(xc, u0, cc);
      d.c.push_back(cc(i));
    }
    for (int i = 0; i < p; ++i) {
      const double dx = x1(d.xi) - d.a[i], dy = x1(d.yi) - d.b[i];
      if (!detail::Close(c1(i), -(dx * dx + dy * dy - d.c[i]), 1e-14)) return fail();
      if (J1(i, 0) != 2 * (d.a[i] - x1(d.xi)) || J1(i, 1) != 2 * (d.b[i] - x1(d.yi))) return fail();
    }
    *out = d;
    return true;
  }
  return fail();
}

}  // namespace device
}  // namespace altro
