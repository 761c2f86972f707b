// oracle/altro_oracle.hpp
//
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU restatement (plain C++17, no Eigen) of the AL-iLQR hot path of
// optimusride/altro-cpp @ d5e8cfe.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / `--impl reference` legs may build, load or call
// anything in oracle/.  The product path (altro_cpp_b200/) never includes this.
//
// Parity status: PINNED against the reference's own known-answer tests
// (tests/test_oracle_golden.py lists every value with its reference file:line)
// and against the reference's own solver code: oracle/build_ref.py compiles it
// where it lies on the host mirror's Eigen stand-in (Eigen 3.3 itself is absent
// from the image) into oracle/_ref, and
// tests/test_oracle_vs_reference_build.py finds this restatement bit-identical
// to it on every instance it solves.
//
// Every function cites the reference file:line it follows.  Arithmetic is done
// in the reference's order of operations: `t`,`h` are
// float and promoted to double in expressions, chained products associate
// left-to-right with temporaries, inner products accumulate k = 0..K-1.
// Matrices are column-major like Eigen: M(i,j) = M[i + j*rows].
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace altro_oracle {

// ---------------------------------------------------------------------------
// Enumerations shared with the C ABI (oracle/altro_oracle.h) and mirrored by
// include/altro_b200.h.
// ---------------------------------------------------------------------------
enum ModelKind : int {
  kModelUnicycle = 0,          // examples/unicycle.cpp:12-33
  kModelTripleIntegrator = 1,  // examples/triple_integrator.cpp:9-33
  kModelCartpole = 2,          // not in the reference (BASELINE config C4)
  kModelLinear = 3,            // not in the reference (BASELINE config C5); discrete LTI
};

enum ConstraintKind : int {
  kConGoal = 0,          // examples/basic_constraints.hpp:15-40   (Equality)
  kConControlBound = 1,  // examples/basic_constraints.hpp:42-150  (Inequality)
  kConCircle = 2,        // examples/obstacle_constraints.hpp:69-126 (Inequality)
};

// altro/common/solver_stats.hpp:20-31
enum SolverStatus : int {
  kSolved = 0,
  kUnsolved = 1,
  kStateLimit = 2,
  kControlLimit = 3,
  kCostIncrease = 4,
  kMaxIterations = 5,
  kMaxOuterIterations = 6,
  kMaxInnerIterations = 7,
  kMaxPenalty = 8,
  kBackwardPassRegularizationFailed = 9,
};

// altro/common/solver_options.hpp:19-65 (numeric fields only)
struct Options {
  int max_iterations_total = 300;
  int max_iterations_outer = 30;
  int max_iterations_inner = 100;
  double cost_tolerance = 1e-4;
  double gradient_tolerance = 1e-2;
  double bp_reg_increase_factor = 1.6;
  double bp_reg_initial = 0.0;
  double bp_reg_max = 1e8;
  double bp_reg_min = 1e-8;
  int bp_reg_fail_threshold = 100;
  int check_forwardpass_bounds = 1;
  double state_max = 1e8;
  double control_max = 1e8;
  int line_search_max_iterations = 20;
  double line_search_lower_bound = 1e-8;
  double line_search_upper_bound = 10.0;
  double line_search_decrease_factor = 2;
  double constraint_tolerance = 1e-4;
  double maximum_penalty = 1e8;
  double initial_penalty = 1.0;
  int reset_duals = 1;
  double penalty_scaling = 10.0;  // constraint_values.hpp:30 kDefaultPenaltyScaling
};

// ---------------------------------------------------------------------------
// Problem description (the data behind altro/problem/problem.hpp:65 for the
// closed set of example functors).
// ---------------------------------------------------------------------------
struct QuadCost {  // examples/quadratic_cost.hpp:12-75
  std::vector<double> Q, R, H, q, r;  // n*n, m*m, n*m (col-major), n, m
  double c = 0.0;
};

struct ConstraintDef {
  int kind = 0;
  bool equality = false;
  int p = 0;  // OutputDimension()
  // goal
  std::vector<double> xf;
  // control bound: finite index lists, basic_constraints.hpp:136-143
  std::vector<double> lb, ub;
  std::vector<int> idx_lb, idx_ub;
  // circles: obstacle_constraints.hpp:69-126
  std::vector<double> cx, cy, cr;
  int xi = 0, yi = 1;
};

struct Problem {
  int n = 0, m = 0, N = 0;
  int model = kModelUnicycle;
  std::vector<double> model_params;
  std::vector<float> h;  // N+1 (h[N] = 0), altro/common/trajectory.hpp:122-130
  std::vector<float> t;  // N+1
  std::vector<std::shared_ptr<QuadCost>> cost;        // N+1
  std::vector<std::vector<ConstraintDef>> eq, ineq;   // N+1 each, insertion order
  std::vector<double> x0;

  Problem(int n_, int m_, int N_)
      : n(n_), m(m_), N(N_), h(N_ + 1, 0.f), t(N_ + 1, 0.f), cost(N_ + 1), eq(N_ + 1),
        ineq(N_ + 1), x0(n_, 0.0) {}

  // altro/common/trajectory.hpp:122-130
  void SetUniformStep(float hs) {
    for (int k = 0; k < N; ++k) {
      h[k] = hs;
      t[k] = static_cast<float>(k) * hs;
    }
    h[N] = 0.0f;
    t[N] = static_cast<float>(hs) * N;
  }
};

// ---------------------------------------------------------------------------
// Continuous models (value + Jacobian). jac is n x (n+m) column-major and is
// only written where the reference writes it.
// ---------------------------------------------------------------------------
inline void ModelEvaluate(const Problem& P, const double* x, const double* u, float /*t*/,
                          double* xdot) {
  switch (P.model) {
    case kModelUnicycle: {  // examples/unicycle.cpp:12-21
      double theta = x[2];
      double v = u[0];
      double omega = u[1];
      xdot[0] = v * std::cos(theta);
      xdot[1] = v * std::sin(theta);
      xdot[2] = omega;
      break;
    }
    case kModelTripleIntegrator: {  // examples/triple_integrator.cpp:9-19
      const int dof = P.m;
      for (int i = 0; i < dof; ++i) {
        xdot[i] = x[i + dof];
        xdot[i + dof] = x[i + 2 * dof];
        xdot[i + 2 * dof] = u[i];
      }
      break;
    }
    case kModelCartpole: {
      // Frictionless cart-pole, state (x, theta, xd, thetad), theta = 0 hanging
      // down.  Definition owned by this repo; the device code
      // in altro_cpp_b200/csrc/models.cuh implements the same closed form.
      const double mc = P.model_params[0], mp = P.model_params[1], l = P.model_params[2],
                   g = P.model_params[3];
      const double th = x[1], xd = x[2], thd = x[3];
      const double s = std::sin(th), c = std::cos(th);
      const double den = mc + mp * s * s;
      const double F = u[0];
      const double xdd = (F + mp * s * (l * thd * thd + g * c)) / den;
      const double thdd = (-F * c - mp * l * thd * thd * c * s - (mc + mp) * g * s) / (l * den);
      xdot[0] = xd;
      xdot[1] = thd;
      xdot[2] = xdd;
      xdot[3] = thdd;
      break;
    }
    default:
      break;
  }
}

inline void ModelJacobian(const Problem& P, const double* x, const double* u, float /*t*/,
                          double* jac) {
  const int n = P.n;
  switch (P.model) {
    case kModelUnicycle: {  // examples/unicycle.cpp:23-33 (writes 5 entries only)
      double theta = x[2];
      double v = u[0];
      jac[0 + 2 * n] = -v * std::sin(theta);
      jac[0 + 3 * n] = std::cos(theta);
      jac[1 + 2 * n] = v * std::cos(theta);
      jac[1 + 3 * n] = std::sin(theta);
      jac[2 + 4 * n] = 1;
      break;
    }
    case kModelTripleIntegrator: {  // examples/triple_integrator.cpp:21-33
      const int dof = P.m;
      for (int i = 0; i < n * (n + P.m); ++i) jac[i] = 0.0;
      for (int i = 0; i < dof; ++i) {
        jac[i + (i + dof) * n] = 1;
        jac[(i + dof) + (i + 2 * dof) * n] = 1;
        jac[(i + 2 * dof) + (i + 3 * dof) * n] = 1;
      }
      break;
    }
    case kModelCartpole: {
      const double mc = P.model_params[0], mp = P.model_params[1], l = P.model_params[2],
                   g = P.model_params[3];
      const double th = x[1], thd = x[3];
      const double s = std::sin(th), c = std::cos(th);
      const double den = mc + mp * s * s;
      const double F = u[0];
      const double numx = F + mp * s * (l * thd * thd + g * c);
      const double numt = -F * c - mp * l * thd * thd * c * s - (mc + mp) * g * s;
      const double dden = 2.0 * mp * s * c;  // d den / d theta
      const double dnumx = mp * c * (l * thd * thd + g * c) - mp * s * g * s;
      const double dnumt = F * s - mp * l * thd * thd * (c * c - s * s) - (mc + mp) * g * c;
      for (int i = 0; i < n * (n + P.m); ++i) jac[i] = 0.0;
      jac[0 + 2 * n] = 1.0;
      jac[1 + 3 * n] = 1.0;
      jac[2 + 1 * n] = (dnumx * den - numx * dden) / (den * den);
      jac[2 + 3 * n] = (2.0 * mp * s * l * thd) / den;
      jac[2 + 4 * n] = 1.0 / den;
      jac[3 + 1 * n] = (dnumt * den - numt * dden) / (l * den * den);
      jac[3 + 3 * n] = (-2.0 * mp * l * thd * c * s) / (l * den);
      jac[3 + 4 * n] = -c / (l * den);
      break;
    }
    default:
      break;
  }
}

// ---------------------------------------------------------------------------
// Discrete dynamics.  RK4 value: altro/problem/integration.hpp:124-131.
// ---------------------------------------------------------------------------
template <int n, int m>
inline void DiscreteEvaluate(const Problem& P, const double* x, const double* u, float t, float h,
                             double* xnext) {
  if (P.model == kModelLinear) {  // x+ = A x + B u, params = [A (n*n col-major), B (n*m)]
    const double* A = P.model_params.data();
    const double* B = A + n * n;
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc += A[i + j * n] * x[j];
      double accb = 0.0;
      for (int j = 0; j < m; ++j) accb += B[i + j * n] * u[j];
      xnext[i] = acc + accb;
    }
    return;
  }
  double k1[n], k2[n], k3[n], k4[n], xt[n];
  ModelEvaluate(P, x, u, t, k1);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * 0.5 * h;
  ModelEvaluate(P, xt, u, static_cast<float>(t + 0.5 * h), k2);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k2[i] * 0.5 * h;
  ModelEvaluate(P, xt, u, static_cast<float>(t + 0.5 * h), k3);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k3[i] * h;
  ModelEvaluate(P, xt, u, t + h, k4);
  for (int i = 0; i < n; ++i) {
    xnext[i] = x[i] + h * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) / 6;
  }
}

// n x n times n x c product, sequential accumulation (Eigen coefficient-based
// lazy product order for small fixed sizes).
template <int r, int k, int c>
inline void MatMul(const double* A, const double* B, double* C) {
  for (int j = 0; j < c; ++j) {
    for (int i = 0; i < r; ++i) {
      double acc = A[i] * B[j * k];
      for (int l = 1; l < k; ++l) acc += A[i + l * r] * B[l + j * k];
      C[i + j * r] = acc;
    }
  }
}
// C = A^T * B with A (k x r), B (k x c)
template <int r, int k, int c>
inline void MatTMul(const double* A, const double* B, double* C) {
  for (int j = 0; j < c; ++j) {
    for (int i = 0; i < r; ++i) {
      double acc = A[i * k] * B[j * k];
      for (int l = 1; l < k; ++l) acc += A[l + i * k] * B[l + j * k];
      C[i + j * r] = acc;
    }
  }
}

// RK4 Jacobian: altro/problem/integration.hpp:132-167.  jac (n x (n+m)) must be
// zero on entry (knot_point_function_type.hpp:125).
template <int n, int m>
inline void DiscreteJacobian(const Problem& P, const double* x, const double* u, float t, float h,
                             double* jac) {
  if (P.model == kModelLinear) {
    std::memcpy(jac, P.model_params.data(), sizeof(double) * n * (n + m));
    return;
  }
  double k1[n], k2[n], k3[n], xt[n];
  double A[4][n * n], B[4][n * m], dA[4][n * n], dB[4][n * m];
  ModelEvaluate(P, x, u, t, k1);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * 0.5 * h;
  ModelEvaluate(P, xt, u, static_cast<float>(t + 0.5 * h), k2);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k2[i] * 0.5 * h;
  ModelEvaluate(P, xt, u, static_cast<float>(t + 0.5 * h), k3);

  auto split = [&](int s) {
    std::memcpy(A[s], jac, sizeof(double) * n * n);
    std::memcpy(B[s], jac + n * n, sizeof(double) * n * m);
  };
  ModelJacobian(P, x, u, t, jac);
  split(0);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + 0.5 * k1[i] * h;
  ModelJacobian(P, xt, u, static_cast<float>(0.5 * t), jac);  // Q2: 0.5*t
  split(1);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + 0.5 * k2[i] * h;
  ModelJacobian(P, xt, u, static_cast<float>(0.5 * t), jac);
  split(2);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k3[i] * h;
  ModelJacobian(P, xt, u, t, jac);
  split(3);

  double T[n * n], T2[n * n], TB[n * m];
  // dA0 = A0*h
  for (int i = 0; i < n * n; ++i) dA[0][i] = A[0][i] * h;
  // dA1 = A1*(I + 0.5*dA0)*h ; dA2 = A2*(I + 0.5*dA1)*h ; dA3 = A3*(I + dA2)*h
  for (int s = 1; s < 4; ++s) {
    const double f = (s == 3) ? 1.0 : 0.5;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i)
        T[i + j * n] = (i == j ? 1.0 : 0.0) + (s == 3 ? dA[s - 1][i + j * n] : f * dA[s - 1][i + j * n]);
    MatMul<n, n, n>(A[s], T, T2);
    for (int i = 0; i < n * n; ++i) dA[s][i] = T2[i] * h;
  }
  // dB0 = B0*h ; dB1 = B1*h + 0.5*A1*dB0*h ; dB2 = B2*h + 0.5*A2*dB1*h ; dB3 = B3*h + A3*dB2*h
  for (int i = 0; i < n * m; ++i) dB[0][i] = B[0][i] * h;
  for (int s = 1; s < 4; ++s) {
    if (s < 3) {
      for (int i = 0; i < n * n; ++i) T[i] = 0.5 * A[s][i];
      MatMul<n, n, m>(T, dB[s - 1], TB);
    } else {
      MatMul<n, n, m>(A[s], dB[s - 1], TB);
    }
    for (int i = 0; i < n * m; ++i) dB[s][i] = B[s][i] * h + TB[i] * h;
  }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      const int q = i + j * n;
      jac[q] = (i == j ? 1.0 : 0.0) + (dA[0][q] + 2 * dA[1][q] + 2 * dA[2][q] + dA[3][q]) / 6;
    }
  for (int q = 0; q < n * m; ++q) {
    jac[n * n + q] = (dB[0][q] + 2 * dB[1][q] + 2 * dB[2][q] + dB[3][q]) / 6;
  }
}

// ---------------------------------------------------------------------------
// QuadraticCost: examples/quadratic_cost.cpp:8-28
// ---------------------------------------------------------------------------
template <int n, int m>
inline double QuadEvaluate(const QuadCost& C, const double* x, const double* u) {
  double Qx[n], Hu[n], Ru[m];
  for (int i = 0; i < n; ++i) {
    double a = C.Q[i] * x[0];
    for (int j = 1; j < n; ++j) a += C.Q[i + j * n] * x[j];
    Qx[i] = a;
    double b = C.H[i] * u[0];
    for (int j = 1; j < m; ++j) b += C.H[i + j * n] * u[j];
    Hu[i] = b;
  }
  for (int i = 0; i < m; ++i) {
    double a = C.R[i] * u[0];
    for (int j = 1; j < m; ++j) a += C.R[i + j * m] * u[j];
    Ru[i] = a;
  }
  auto dot = [](const double* a, const double* b, int len) {
    double s = a[0] * b[0];
    for (int i = 1; i < len; ++i) s += a[i] * b[i];
    return s;
  };
  return 0.5 * dot(x, Qx, n) + dot(x, Hu, n) + 0.5 * dot(u, Ru, m) + dot(C.q.data(), x, n) +
         dot(C.r.data(), u, m) + C.c;
}

template <int n, int m>
inline void QuadGradient(const QuadCost& C, const double* x, const double* u, double* dx,
                         double* du) {
  // dx = Q*x + q + H*u ; du = R*u + r + H^T*x
  for (int i = 0; i < n; ++i) {
    double a = C.Q[i] * x[0];
    for (int j = 1; j < n; ++j) a += C.Q[i + j * n] * x[j];
    double b = C.H[i] * u[0];
    for (int j = 1; j < m; ++j) b += C.H[i + j * n] * u[j];
    dx[i] = a + C.q[i] + b;
  }
  for (int i = 0; i < m; ++i) {
    double a = C.R[i] * u[0];
    for (int j = 1; j < m; ++j) a += C.R[i + j * m] * u[j];
    double b = C.H[i * n] * x[0];
    for (int j = 1; j < n; ++j) b += C.H[j + i * n] * x[j];
    du[i] = a + C.r[i] + b;
  }
}

// ---------------------------------------------------------------------------
// Constraint leaf functions.
// ---------------------------------------------------------------------------
inline void ConEvaluate(const ConstraintDef& D, int n, const double* x, const double* u,
                        double* c) {
  (void)n;
  switch (D.kind) {
    case kConGoal:  // basic_constraints.hpp:27-30
      for (int i = 0; i < D.p; ++i) c[i] = x[i] - D.xf[i];
      break;
    case kConControlBound: {  // basic_constraints.hpp:98-111
      const int nl = static_cast<int>(D.idx_lb.size());
      for (int i = 0; i < nl; ++i) c[i] = D.lb[D.idx_lb[i]] - u[D.idx_lb[i]];
      for (size_t i = 0; i < D.idx_ub.size(); ++i) c[i + nl] = u[D.idx_ub[i]] - D.ub[D.idx_ub[i]];
      break;
    }
    case kConCircle: {  // obstacle_constraints.hpp:98-106, Circle::Distance2 :39-42
      const double px = x[D.xi], py = x[D.yi];
      for (int i = 0; i < D.p; ++i) {
        const double dx = px - D.cx[i], dy = py - D.cy[i];
        c[i] = -(dx * dx + dy * dy - D.cr[i] * D.cr[i]);
      }
      break;
    }
  }
}

// jac is p x (n+m) column-major, persistent per constraint (only the entries the
// reference writes are written; see each case).
inline void ConJacobian(const ConstraintDef& D, int n, int m, const double* x, const double* u,
                        double* jac) {
  (void)u;
  const int p = D.p;
  switch (D.kind) {
    case kConGoal:  // basic_constraints.hpp:31-36: jac.setIdentity() on p x (n+m)
      for (int j = 0; j < n + m; ++j)
        for (int i = 0; i < p; ++i) jac[i + j * p] = (i == j) ? 1.0 : 0.0;
      break;
    case kConControlBound: {  // basic_constraints.hpp:113-129
      for (int q = 0; q < p * (n + m); ++q) jac[q] = 0.0;
      const int nl = static_cast<int>(D.idx_lb.size());
      for (int i = 0; i < nl; ++i) jac[i + (n + D.idx_lb[i]) * p] = -1;
      for (size_t i = 0; i < D.idx_ub.size(); ++i) jac[(i + nl) + (n + D.idx_ub[i]) * p] = 1;
      break;
    }
    case kConCircle: {  // obstacle_constraints.hpp:108-120 (writes columns 0 and 1!)
      const double px = x[D.xi], py = x[D.yi];
      for (int i = 0; i < p; ++i) {
        jac[i + 0 * p] = 2 * (D.cx[i] - px);
        jac[i + 1 * p] = 2 * (D.cy[i] - py);
      }
      break;
    }
  }
}

// ---------------------------------------------------------------------------
// ConstraintValues<n,m,ConType>: altro/constraints/constraint_values.hpp:24-273
// Cones: altro/constraints/constraint.hpp:28-122.  The dual cone of Equality is
// the identity, of Inequality the negative orthant (self-dual).
// ---------------------------------------------------------------------------
struct ConstraintValues {
  ConstraintDef def;
  int n, m, p;
  std::vector<double> c, lambda, penalty, jac, lambda_proj, c_proj, proj_jac, jac_proj, arg;
  double penalty_scaling = 10.0;

  ConstraintValues(int n_, int m_, const ConstraintDef& d)
      : def(d), n(n_), m(m_), p(d.p), c(p, 0.0), lambda(p, 0.0), penalty(p, 1.0),
        jac(p * (n_ + m_), 0.0), lambda_proj(p, 0.0), c_proj(p, 0.0), proj_jac(p * p, 0.0),
        jac_proj(p * (n_ + m_), 0.0), arg(p, 0.0) {}

  void DualProjection(const double* v, double* out) const {
    if (def.equality) {  // IdentityCone::Projection constraint.hpp:70-73
      for (int i = 0; i < p; ++i) out[i] = v[i];
    } else {  // NegativeOrthant::Projection constraint.hpp:103-108
      for (int i = 0; i < p; ++i) out[i] = std::min(0.0, v[i]);
    }
  }
  void DualProjJacobian(const double* v) {
    if (def.equality) {  // IdentityCone::Jacobian constraint.hpp:74-78
      for (int j = 0; j < p; ++j)
        for (int i = 0; i < p; ++i) proj_jac[i + j * p] = (i == j) ? 1.0 : 0.0;
    } else {  // NegativeOrthant::Jacobian constraint.hpp:109-114 (diagonal only, Q12)
      for (int i = 0; i < p; ++i) proj_jac[i + i * p] = v[i] > 0 ? 0 : 1;
    }
  }

  // constraint_values.hpp:111-119
  double AugLag(const double* x, const double* u) {
    const double rho = penalty[0];
    ConEvaluate(def, n, x, u, c.data());
    for (int i = 0; i < p; ++i) arg[i] = lambda[i] - rho * c[i];
    DualProjection(arg.data(), lambda_proj.data());
    double a = 0.0, b = 0.0;
    for (int i = 0; i < p; ++i) a += lambda_proj[i] * lambda_proj[i];
    for (int i = 0; i < p; ++i) b += lambda[i] * lambda[i];
    double J = a - b;
    J = J / (2 * rho);
    return J;
  }

  void CalcJacProj(const double* x, const double* u) {
    const double rho = penalty[0];
    ConEvaluate(def, n, x, u, c.data());
    ConJacobian(def, n, m, x, u, jac.data());
    for (int i = 0; i < p; ++i) arg[i] = lambda[i] - rho * c[i];
    DualProjection(arg.data(), lambda_proj.data());
    DualProjJacobian(arg.data());
    // jac_proj = proj_jac * jac  (p x p times p x (n+m)), dense
    for (int j = 0; j < n + m; ++j)
      for (int i = 0; i < p; ++i) {
        double acc = proj_jac[i] * jac[j * p];
        for (int l = 1; l < p; ++l) acc += proj_jac[i + l * p] * jac[l + j * p];
        jac_proj[i + j * p] = acc;
      }
  }

  // constraint_values.hpp:131-143
  void AugLagGradient(const double* x, const double* u, double* dx, double* du) {
    CalcJacProj(x, u);
    for (int a = 0; a < n; ++a) {
      double acc = (-jac_proj[a * p]) * lambda_proj[0];
      for (int i = 1; i < p; ++i) acc += (-jac_proj[i + a * p]) * lambda_proj[i];
      dx[a] = acc;
    }
    for (int a = 0; a < m; ++a) {
      double acc = (-jac_proj[(n + a) * p]) * lambda_proj[0];
      for (int i = 1; i < p; ++i) acc += (-jac_proj[i + (n + a) * p]) * lambda_proj[i];
      du[a] = acc;
    }
  }

  // constraint_values.hpp:156-177 (Gauss-Newton)
  void AugLagHessian(const double* x, const double* u, double* dxdx, double* dxdu, double* dudu) {
    const double rho = penalty[0];
    CalcJacProj(x, u);
    auto term = [&](int ca, int cb) {
      double acc = (rho * jac_proj[ca * p]) * jac_proj[cb * p];
      for (int i = 1; i < p; ++i) acc += (rho * jac_proj[i + ca * p]) * jac_proj[i + cb * p];
      return acc;
    };
    for (int b = 0; b < n; ++b)
      for (int a = 0; a < n; ++a) dxdx[a + b * n] = term(a, b);
    for (int b = 0; b < m; ++b)
      for (int a = 0; a < n; ++a) dxdu[a + b * n] = term(a, n + b);
    for (int b = 0; b < m; ++b)
      for (int a = 0; a < m; ++a) dudu[a + b * m] = term(n + a, n + b);
  }

  // constraint_values.hpp:192-194
  void UpdateDuals() {
    for (int i = 0; i < p; ++i) arg[i] = lambda[i] - penalty[i] * c[i];
    DualProjection(arg.data(), lambda.data());
  }
  // constraint_values.hpp:202-207
  void UpdatePenalties() {
    for (int i = 0; i < p; ++i) penalty[i] *= penalty_scaling;
  }
  // constraint_values.hpp:216-221 with cone projection constraint.hpp:33-37,103-108
  double MaxViolation() {
    double v = 0.0;
    for (int i = 0; i < p; ++i) {
      const double proj = def.equality ? 0.0 : std::min(0.0, c[i]);
      c_proj[i] = c[i] - proj;
      v = std::max(v, std::fabs(c_proj[i]));
    }
    return v;
  }
  double MaxPenalty() const {
    double v = penalty[0];
    for (int i = 1; i < p; ++i) v = std::max(v, penalty[i]);
    return v;
  }
};

// ---------------------------------------------------------------------------
// ALCost<n,m>: altro/augmented_lagrangian/al_cost.hpp:35-438
// ---------------------------------------------------------------------------
template <int n, int m>
struct ALCost {
  std::shared_ptr<QuadCost> costfun;
  std::vector<ConstraintValues> eq, ineq;
  bool plain = false;  // true: behaves as the bare QuadraticCost (no ALCost wrapper)

  double Evaluate(const double* x, const double* u) {  // al_cost.hpp:264-274
    double J = QuadEvaluate<n, m>(*costfun, x, u);
    for (auto& cv : eq) J += cv.AugLag(x, u);
    for (auto& cv : ineq) J += cv.AugLag(x, u);
    return J;
  }
  void Gradient(const double* x, const double* u, double* dx, double* du) {  // :276-290
    double dxt[n], dut[m];
    QuadGradient<n, m>(*costfun, x, u, dx, du);
    for (auto* vec : {&eq, &ineq})
      for (auto& cv : *vec) {
        cv.AugLagGradient(x, u, dxt, dut);
        for (int i = 0; i < n; ++i) dx[i] += dxt[i];
        for (int i = 0; i < m; ++i) du[i] += dut[i];
      }
  }
  void Hessian(const double* x, const double* u, double* dxdx, double* dxdu,
               double* dudu) {  // :292-308
    double a[n * n], b[n * m], c[m * m];
    std::memcpy(dxdx, costfun->Q.data(), sizeof(double) * n * n);  // quadratic_cost.cpp:20-28
    std::memcpy(dudu, costfun->R.data(), sizeof(double) * m * m);
    std::memcpy(dxdu, costfun->H.data(), sizeof(double) * n * m);
    for (auto* vec : {&eq, &ineq})
      for (auto& cv : *vec) {
        cv.AugLagHessian(x, u, a, b, c);
        for (int i = 0; i < n * n; ++i) dxdx[i] += a[i];
        for (int i = 0; i < n * m; ++i) dxdu[i] += b[i];
        for (int i = 0; i < m * m; ++i) dudu[i] += c[i];
      }
  }
  void UpdateDuals() {  // :314-321
    for (auto& cv : eq) cv.UpdateDuals();
    for (auto& cv : ineq) cv.UpdateDuals();
  }
  void UpdatePenalties() {  // :327-334
    for (auto& cv : eq) cv.UpdatePenalties();
    for (auto& cv : ineq) cv.UpdatePenalties();
  }
  double MaxViolation() {  // :343-352
    double ve = 0.0, vi = 0.0;
    for (auto& cv : eq) ve = std::max(ve, std::fabs(cv.MaxViolation()));
    for (auto& cv : ineq) vi = std::max(vi, std::fabs(cv.MaxViolation()));
    return std::max(ve, vi);
  }
  double MaxPenalty() const {  // :360-369
    double v = 0.0;
    for (auto& cv : eq) v = std::max(v, cv.MaxPenalty());
    for (auto& cv : ineq) v = std::max(v, cv.MaxPenalty());
    return v;
  }
  void SetPenalty(double rho) {
    for (auto* vec : {&eq, &ineq})
      for (auto& cv : *vec) std::fill(cv.penalty.begin(), cv.penalty.end(), rho);
  }
  void SetPenaltyScaling(double phi) {
    for (auto* vec : {&eq, &ineq})
      for (auto& cv : *vec) cv.penalty_scaling = phi;
  }
  void ResetDualVariables() {  // :371-378
    for (auto* vec : {&eq, &ineq})
      for (auto& cv : *vec) std::fill(cv.lambda.begin(), cv.lambda.end(), 0.0);
  }
};

// ---------------------------------------------------------------------------
// KnotPointFunctions<n,m>: altro/ilqr/knot_point_function_type.hpp:37-299
// ---------------------------------------------------------------------------
template <int n, int m>
struct KnotPointFunctions {
  // cost expansion (cost_expansion.hpp:27)
  double lxx[n * n], lxu[n * m], luu[m * m], lx[n], lu[m];
  // dynamics expansion (dynamics_expansion.hpp:18): [A | B], n x (n+m)
  double jac[n * (n + m)];
  // action-value expansion and its regularised copy
  double Qxx[n * n], Qxu[n * m], Quu[m * m], Qx[n], Qu[m];
  double Quu_reg[m * m];
  double K[m * n], d[m];
  double P[n * n], p[n];
  double ctg_delta[2];

  KnotPointFunctions() { std::memset(this, 0, sizeof(*this)); }

  const double* A() const { return jac; }
  const double* B() const { return jac + n * n; }

  void CalcTerminalCostToGo() {  // :135-138
    std::memcpy(P, lxx, sizeof(P));
    std::memcpy(p, lx, sizeof(p));
  }

  // :149-164 (Q20: (A^T P) A association)
  void CalcActionValueExpansion(const double* Pn, const double* pn) {
    double AtP[n * n], BtP[m * n], T1[n * n], T2[n * m], T3[m * m], v[n], w[m];
    MatTMul<n, n, n>(A(), Pn, AtP);
    MatMul<n, n, n>(AtP, A(), T1);
    for (int i = 0; i < n * n; ++i) Qxx[i] = lxx[i] + T1[i];
    MatMul<n, n, m>(AtP, B(), T2);
    for (int i = 0; i < n * m; ++i) Qxu[i] = lxu[i] + T2[i];
    MatTMul<m, n, n>(B(), Pn, BtP);
    MatMul<m, n, m>(BtP, B(), T3);
    for (int i = 0; i < m * m; ++i) Quu[i] = luu[i] + T3[i];
    MatTMul<n, n, 1>(A(), pn, v);
    for (int i = 0; i < n; ++i) Qx[i] = lx[i] + v[i];
    MatTMul<m, n, 1>(B(), pn, w);
    for (int i = 0; i < m; ++i) Qu[i] = lu[i] + w[i];
  }

  // :175-186 (control-only regularisation)
  void RegularizeActionValue(double rho) {
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) Quu_reg[i + j * m] = Quu[i + j * m] + (i == j ? 1.0 : 0.0) * rho;
  }

  // :197-211.  Eigen 3.3 LLT<Lower> unblocked + two triangular solves (Q19).
  // Returns true on success.
  bool CalcGains() {
    double L[m * m];
    std::memcpy(L, Quu_reg, sizeof(L));
    for (int k = 0; k < m; ++k) {
      double x = L[k + k * m];
      if (k > 0) {
        double sq = 0.0;
        for (int j = 0; j < k; ++j) sq += L[k + j * m] * L[k + j * m];
        x -= sq;
      }
      if (x <= 0.0) return false;
      x = std::sqrt(x);
      L[k + k * m] = x;
      for (int i = k + 1; i < m; ++i) {
        double a = L[i + k * m];
        if (k > 0) {
          double dot = 0.0;
          for (int j = 0; j < k; ++j) dot += L[i + j * m] * L[k + j * m];
          a -= dot;
        }
        L[i + k * m] = a / x;
      }
    }
    auto solve = [&](double* b) {  // in place, L L^T y = b
      for (int i = 0; i < m; ++i) {
        double s = b[i];
        for (int j = 0; j < i; ++j) s -= L[i + j * m] * b[j];
        b[i] = s / L[i + i * m];
      }
      for (int i = m - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < m; ++j) s -= L[j + i * m] * b[j];
        b[i] = s / L[i + i * m];
      }
    };
    // K = -(Quu_reg \ Qxu^T)  (m x n); d = -(Quu_reg \ Qu).  The regularised
    // copy of Qxu/Qu equals the unregularised one (:178).
    for (int c = 0; c < n; ++c) {
      double b[m];
      for (int i = 0; i < m; ++i) b[i] = Qxu[c + i * n];
      solve(b);
      for (int i = 0; i < m; ++i) K[i + c * m] = b[i] * -1;
    }
    double b[m];
    for (int i = 0; i < m; ++i) b[i] = Qu[i];
    solve(b);
    for (int i = 0; i < m; ++i) d[i] = b[i] * -1;
    return true;
  }

  // :220-230 (Q3: unregularised Q)
  void CalcCostToGo() {
    double KtQuu[n * m], v1[n], v2[n], v3[n], T1[n * n], T2[n * n], T3[n * n], Quud[m];
    MatTMul<n, m, m>(K, Quu, KtQuu);   // K^T Quu  (n x m)
    MatMul<n, m, 1>(KtQuu, d, v1);     // (K^T Quu) d
    MatTMul<n, m, 1>(K, Qu, v2);       // K^T Qu
    MatMul<n, m, 1>(Qxu, d, v3);       // Qxu d
    for (int i = 0; i < n; ++i) p[i] = Qx[i] + v1[i] + v2[i] + v3[i];
    MatMul<n, m, n>(KtQuu, K, T1);     // (K^T Quu) K
    // K^T Qxu^T : (n x m)(m x n) ; Qxu^T(i,j) = Qxu(j,i)
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double acc = K[i * m] * Qxu[j];
        for (int l = 1; l < m; ++l) acc += K[l + i * m] * Qxu[j + l * n];
        T2[i + j * n] = acc;
      }
    MatMul<n, m, n>(Qxu, K, T3);       // Qxu K
    for (int i = 0; i < n * n; ++i) P[i] = Qxx[i] + T1[i] + T2[i] + T3[i];
    double a = d[0] * Qu[0];
    for (int i = 1; i < m; ++i) a += d[i] * Qu[i];
    ctg_delta[0] = a;
    MatMul<m, m, 1>(Quu, d, Quud);
    double b = d[0] * Quud[0];
    for (int i = 1; i < m; ++i) b += d[i] * Quud[i];
    ctg_delta[1] = 0.5 * b;
  }
};

// ---------------------------------------------------------------------------
// SolverStats carry-forward semantics (altro/common/solver_stats.cpp:31-66,
// solver_stats.hpp:193-204): Log() writes the last slot, NewIteration() appends
// a slot pre-filled with the previous value.
// ---------------------------------------------------------------------------
struct Stats {
  double initial_cost = 0.0;
  int iterations_inner = 0, iterations_outer = 0, iterations_total = 0;
  std::vector<double> cost, alpha, z, gradient, cost_decrease, regularization, violations,
      max_penalty;
  int len = 0;
  std::vector<double>* all[8] = {&cost, &alpha, &z, &gradient, &cost_decrease, &regularization,
                                 &violations, &max_penalty};
  Stats() = default;
  Stats(const Stats&) = delete;
  void Reset() {
    initial_cost = 0.0;
    iterations_inner = iterations_outer = iterations_total = 0;
    len = 0;
    for (auto* v : all) v->clear();
  }
  void NewIteration() {
    len++;
    for (auto* v : all) {
      v->resize(len);
      v->back() = (len > 1) ? (*v)[len - 2] : 0.0;
    }
  }
  void Log(std::vector<double>& v, double value) {
    if (len == 0) NewIteration();
    v.back() = value;
  }
  void Touch() {  // Log() of an integer field ("iters", "iter_al") still registers slot 0
    if (len == 0) NewIteration();
  }
};

// ---------------------------------------------------------------------------
// iLQR<n,m>: altro/ilqr/ilqr.hpp:47-813
// ---------------------------------------------------------------------------
template <int n, int m>
struct iLQR {
  const Problem* prob;
  int N;
  Options opts;
  Stats stats;
  std::vector<ALCost<n, m>> costfun;  // per knot; `plain` ignores constraints
  std::vector<KnotPointFunctions<n, m>> kp;
  std::vector<double> X, U, Xbar, Ubar;  // (N+1)*n, (N+1)*m  (u_N exists and is zero, Q14)
  std::vector<double> x0;
  std::vector<double> costs, grad;
  int status = kUnsolved;
  double rho = 0.0, drho = 0.0;
  double deltaV[2] = {0.0, 0.0};
  // profiling counters (not in the reference; used by the bench harness)
  long n_backward = 0, n_rollout_cl = 0, n_cost = 0, n_expansions = 0;

  iLQR(const Problem& P, bool use_constraints)
      : prob(&P), N(P.N), costfun(P.N + 1), kp(P.N + 1), X((P.N + 1) * n, 0.0),
        U((P.N + 1) * m, 0.0), Xbar((P.N + 1) * n, 0.0), Ubar((P.N + 1) * m, 0.0), x0(P.x0),
        costs(P.N + 1, 0.0), grad(P.N, 0.0) {
    for (int k = 0; k <= N; ++k) {  // al_problem.hpp:30-51, al_cost.hpp:53-61
      costfun[k].costfun = P.cost[k];
      costfun[k].plain = !use_constraints;
      if (use_constraints) {
        for (auto& d : P.eq[k]) costfun[k].eq.emplace_back(n, m, d);
        for (auto& d : P.ineq[k]) costfun[k].ineq.emplace_back(n, m, d);
      }
    }
    ResetInternalVariables();
  }

  double* x(int k) { return &X[k * n]; }
  double* u(int k) { return &U[k * m]; }
  double* xb(int k) { return &Xbar[k * n]; }
  double* ub(int k) { return &Ubar[k * m]; }

  void ResetInternalVariables() {  // ilqr.hpp:680-690
    status = kUnsolved;
    std::fill(costs.begin(), costs.end(), 0.0);
    std::fill(grad.begin(), grad.end(), 0.0);
    deltaV[0] = deltaV[1] = 0.0;
    rho = opts.bp_reg_initial;
    drho = 0.0;
  }

  void SolveSetup() {  // ilqr.hpp:629-645
    stats.iterations_inner = 0;
    ResetInternalVariables();
  }

  void Rollout() {  // ilqr.hpp:453-459
    for (int i = 0; i < n; ++i) x(0)[i] = x0[i];
    for (int k = 0; k < N; ++k)
      DiscreteEvaluate<n, m>(*prob, x(k), u(k), prob->t[k], prob->h[k], x(k + 1));
  }

  double CostOf(std::vector<double>& Xs, std::vector<double>& Us) {  // ilqr.hpp:326-334,758-763
    n_cost++;
    for (int k = 0; k <= N; ++k) costs[k] = costfun[k].Evaluate(&Xs[k * n], &Us[k * m]);
    double s = 0.0;
    for (int k = 0; k <= N; ++k) s += costs[k];
    return s;
  }
  double Cost() { return CostOf(X, U); }

  void UpdateExpansions() {  // ilqr.hpp:350-366, :670-677
    n_expansions++;
    for (int k = 0; k <= N; ++k) {
      KnotPointFunctions<n, m>& f = kp[k];
      // CalcCostExpansion knot_point_function_type.hpp:107-111
      costfun[k].Gradient(x(k), u(k), f.lx, f.lu);
      costfun[k].Hessian(x(k), u(k), f.lxx, f.lxu, f.luu);
      // CalcDynamicsExpansion :121-128 (knot N holds IdentityDynamics, problem.hpp:161-164)
      std::memset(f.jac, 0, sizeof(f.jac));
      if (k < N) {
        DiscreteJacobian<n, m>(*prob, x(k), u(k), prob->t[k], prob->h[k], f.jac);
      } else {  // IdentityDynamics::Jacobian, problem.hpp:40-43: jac.setIdentity() on n x (n+m)
        for (int i = 0; i < n; ++i) f.jac[i + i * n] = 1.0;
      }
      costs[k] = costfun[k].Evaluate(x(k), u(k));
    }
  }

  void IncreaseRegularization() {  // ilqr.hpp:770-775
    drho = std::max(drho * opts.bp_reg_increase_factor, opts.bp_reg_increase_factor);
    rho = std::max(rho * drho, opts.bp_reg_min);
    rho = std::min(rho, opts.bp_reg_max);
  }
  void DecreaseRegularization() {  // ilqr.hpp:781-786
    drho = std::min(drho / opts.bp_reg_increase_factor, 1 / opts.bp_reg_increase_factor);
    rho = std::max(rho * drho, opts.bp_reg_min);
    rho = std::min(rho, opts.bp_reg_max);
  }

  void BackwardPass() {  // ilqr.hpp:385-445
    n_backward++;
    kp[N].CalcTerminalCostToGo();
    const double* Sxx = kp[N].P;
    const double* Sx = kp[N].p;
    int max_reg_count = 0;
    deltaV[0] = 0.0;
    deltaV[1] = 0.0;
    bool repeat = true;
    while (repeat) {
      for (int k = N - 1; k >= 0; --k) {
        kp[k].CalcActionValueExpansion(Sxx, Sx);
        kp[k].RegularizeActionValue(rho);
        const bool ok = kp[k].CalcGains();
        if (!ok) {
          IncreaseRegularization();
          Sxx = kp[N].P;
          Sx = kp[N].p;
          if (rho >= opts.bp_reg_max) max_reg_count++;
          if (max_reg_count >= opts.bp_reg_fail_threshold) {
            status = kBackwardPassRegularizationFailed;
            repeat = false;
          }
          break;
        }
        kp[k].CalcCostToGo();
        deltaV[0] += kp[k].ctg_delta[0];
        deltaV[1] += kp[k].ctg_delta[1];
        Sxx = kp[k].P;
        Sx = kp[k].p;
        if (k == 0) repeat = false;
      }
    }
    stats.Log(stats.regularization, rho);
    DecreaseRegularization();
  }

  bool RolloutClosedLoop(double alpha) {  // ilqr.hpp:468-499
    n_rollout_cl++;
    for (int i = 0; i < n; ++i) xb(0)[i] = x0[i];
    for (int k = 0; k < N; ++k) {
      const double* K = kp[k].K;
      const double* d = kp[k].d;
      double dx[n];
      for (int i = 0; i < n; ++i) dx[i] = xb(k)[i] - x(k)[i];
      for (int i = 0; i < m; ++i) {
        double acc = K[i] * dx[0];
        for (int j = 1; j < n; ++j) acc += K[i + j * m] * dx[j];
        ub(k)[i] = u(k)[i] + acc + d[i] * alpha;
      }
      DiscreteEvaluate<n, m>(*prob, xb(k), ub(k), prob->t[k], prob->h[k], xb(k + 1));
      if (opts.check_forwardpass_bounds) {
        double sx = 0.0, su = 0.0;
        for (int i = 0; i < n; ++i) sx += xb(k + 1)[i] * xb(k + 1)[i];
        for (int i = 0; i < m; ++i) su += ub(k)[i] * ub(k)[i];
        if (std::sqrt(sx) > opts.state_max) {
          status = kStateLimit;
          return false;
        }
        if (std::sqrt(su) > opts.control_max) {
          status = kControlLimit;
          return false;
        }
      }
    }
    status = kUnsolved;
    return true;
  }

  void ForwardPass() {  // ilqr.hpp:512-558
    double J0 = 0.0;
    for (int k = 0; k <= N; ++k) J0 += costs[k];  // Q7
    double alpha = 1.0;
    double z = -1.0;
    bool success = false;
    double J = J0;
    for (int it = 0; it < opts.line_search_max_iterations; ++it) {
      if (RolloutClosedLoop(alpha)) {
        J = CostOf(Xbar, Ubar);
        const double expected = -alpha * (deltaV[0] + alpha * deltaV[1]);
        if (expected > 0.0) {
          z = (J0 - J) / expected;
        } else {
          z = -1.0;
        }
        if (opts.line_search_lower_bound <= z && z <= opts.line_search_upper_bound && J < J0) {
          success = true;
          stats.Log(stats.cost, J);
          stats.Log(stats.alpha, alpha);
          stats.Log(stats.z, z);
          break;
        }
      }
      alpha /= opts.line_search_decrease_factor;
    }
    if (success) {
      // (*Z_) = (*Zbar_): copies x, u for every knot (u_N of Zbar stays zero)
      X = Xbar;
      U = Ubar;
    } else {
      IncreaseRegularization();
      J = J0;
    }
    if (J > J0) status = kCostIncrease;
  }

  double NormalizedFeedforwardGain() {  // ilqr.hpp:662-668
    for (int k = 0; k < N; ++k) {
      double g = std::fabs(kp[k].d[0]) / (std::fabs(u(k)[0]) + 1);
      for (int i = 1; i < m; ++i) g = std::max(g, std::fabs(kp[k].d[i]) / (std::fabs(u(k)[i]) + 1));
      grad[k] = g;
    }
    double s = 0.0;
    for (int k = 0; k < N; ++k) s += grad[k];
    return s / static_cast<double>(N);
  }

  double MaxViolationStored() {  // al_solver.hpp:417-422 (reads stored c_, Q8)
    double v = 0.0;
    for (int k = 0; k <= N; ++k) v = std::max(v, std::fabs(costfun[k].MaxViolation()));
    return v;
  }

  void UpdateConvergenceStatistics() {  // ilqr.hpp:568-587
    const double dgrad = NormalizedFeedforwardGain();
    double dJ = 0.0;
    if (stats.iterations_inner == 0) {
      dJ = stats.initial_cost - stats.cost.back();
    } else {
      dJ = stats.cost[stats.cost.size() - 2] - stats.cost.back();
    }
    stats.iterations_inner++;
    stats.iterations_total++;
    stats.Log(stats.cost_decrease, dJ);
    stats.Log(stats.violations, MaxViolationStored());
    stats.Log(stats.gradient, dgrad);
    stats.NewIteration();
  }

  bool IsDone() {  // ilqr.hpp:597-619
    const bool cost_decrease = stats.cost_decrease.back() < opts.cost_tolerance;
    const bool gradient = stats.gradient.back() < opts.gradient_tolerance;
    bool is_done = false;
    if (cost_decrease && gradient) {
      status = kSolved;
      is_done = true;
    } else if (stats.iterations_inner >= opts.max_iterations_inner) {
      status = kMaxInnerIterations;
      is_done = true;
    } else if (stats.iterations_total >= opts.max_iterations_total) {
      status = kMaxIterations;
      is_done = true;
    } else if (status != kUnsolved) {
      is_done = true;
    }
    return is_done;
  }

  void Solve() {  // ilqr.hpp:284-316
    SolveSetup();
    Rollout();
    stats.initial_cost = Cost();
    for (int iter = 0; iter < opts.max_iterations_inner; ++iter) {
      UpdateExpansions();
      BackwardPass();
      ForwardPass();
      UpdateConvergenceStatistics();
      if (IsDone()) break;
    }
  }
};

// ---------------------------------------------------------------------------
// AugmentedLagrangianiLQR<n,m>: altro/augmented_lagrangian/al_solver.hpp:28-441
// ---------------------------------------------------------------------------
template <int n, int m>
struct ALSolver {
  iLQR<n, m> ilqr;
  int status = kUnsolved;

  explicit ALSolver(const Problem& P, bool use_constraints = true) : ilqr(P, use_constraints) {}

  Options& opts() { return ilqr.opts; }
  Stats& stats() { return ilqr.stats; }

  void SetPenalty(double rho) {  // :271-276
    for (auto& c : ilqr.costfun) c.SetPenalty(rho);
  }
  void SetPenaltyScaling(double phi) {  // :278-284
    for (auto& c : ilqr.costfun) c.SetPenaltyScaling(phi);
  }
  double GetMaxViolation() { return ilqr.MaxViolationStored(); }  // :417-422
  double MaxViolation() {  // :404-408
    ilqr.Cost();
    return GetMaxViolation();
  }
  double GetMaxPenalty() {  // :425-432
    double v = 0.0;
    for (auto& c : ilqr.costfun) v = std::max(v, c.MaxPenalty());
    return v;
  }
  void UpdateDuals() {  // :336-345
    for (auto& c : ilqr.costfun) c.UpdateDuals();
  }
  void UpdatePenalties() {  // :347-355
    for (auto& c : ilqr.costfun) c.UpdatePenalties();
  }

  void Init() {  // :287-302
    if (opts().reset_duals)
      for (auto& c : ilqr.costfun) c.ResetDualVariables();
    if (opts().initial_penalty > 0) SetPenalty(opts().initial_penalty);
    SetPenaltyScaling(opts().penalty_scaling);
    stats().Reset();
    stats().Touch();
    stats().Log(stats().violations, MaxViolation());
    stats().Log(stats().max_penalty, GetMaxPenalty());
  }

  bool IsDone() {  // :368-401 (Q16)
    Stats& s = stats();
    const bool sat = s.violations.back() < opts().constraint_tolerance;
    const bool maxpen = s.max_penalty.back() > opts().maximum_penalty;
    const bool maxouter = s.iterations_outer >= opts().max_iterations_outer;
    const bool maxtotal = s.iterations_total >= opts().max_iterations_total;
    if (ilqr.status != kSolved) {
      status = ilqr.status;
      return true;
    }
    if (sat) {
      status = kSolved;
      return true;
    }
    if (maxpen) {
      status = kMaxPenalty;
      return true;
    }
    if (maxouter) {
      status = kMaxOuterIterations;
      return true;
    }
    if (maxtotal) {
      status = kMaxIterations;
      return true;
    }
    return false;
  }

  void Solve() {  // :304-334
    Init();
    for (int it = 0; it < opts().max_iterations_outer; ++it) {
      ilqr.Solve();
      UpdateDuals();
      // UpdateConvergenceStatistics :357-365
      stats().iterations_outer++;
      stats().Log(stats().violations, GetMaxViolation());
      stats().Log(stats().max_penalty, GetMaxPenalty());
      if (IsDone()) break;
      UpdatePenalties();
    }
  }
};

}  // namespace altro_oracle
