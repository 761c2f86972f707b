// device.cuh — per-knot device functions of the AL-iLQR hot path, one problem instance per
// thread, everything in registers, sizes known at compile time.
//
// These are NOT translations of the reference's Eigen code: the constraint terms are evaluated
// structure-aware (a control bound touches one diagonal entry, a circle a 2x2 block) and no
// dense p x (n+m) Jacobian is ever formed.  They are written so that every floating-point
// sum keeps the association and term order of the reference — skipping
// exact-zero terms does not change a sum — which keeps the fp64 results within ~1e-12 of the
// CPU oracle and the discrete decisions (line search, Cholesky success) identical.
#pragma once

#include "common.cuh"

namespace altro_b200 {

#define ALTRO_UNROLL _Pragma("unroll")

// x / y given ry = RN(1/y): one multiply and two FMAs instead of the ~12-instruction division
// sequence.  With a correctly rounded reciprocal the residual correction returns the correctly
// rounded quotient (Markstein), i.e. the same bits as `x / y`, so the reference's divisions
// (`/ 6` in RK4, the triangular solves of LLT, `/ (2 rho)`) keep their results.
__device__ __forceinline__ double div_by(double x, double y, double ry) {
  const double q = x * ry;
  const double r = fma(-q, y, x);
  return fma(r, ry, q);
}

// C (r x c) = A (r x k) * B (k x c), column-major, inner index ascending.
template <int r, int k, int c>
__device__ __forceinline__ void matmul(const double* A, const double* B, double* C) {
  ALTRO_UNROLL
  for (int j = 0; j < c; ++j) {
    ALTRO_UNROLL
    for (int i = 0; i < r; ++i) {
      double acc = A[i] * B[j * k];
      ALTRO_UNROLL
      for (int l = 1; l < k; ++l) acc += A[i + l * r] * B[l + j * k];
      C[i + j * r] = acc;
    }
  }
}
// C (r x c) = A^T * B with A (k x r), B (k x c)
template <int r, int k, int c>
__device__ __forceinline__ void mattmul(const double* A, const double* B, double* C) {
  ALTRO_UNROLL
  for (int j = 0; j < c; ++j) {
    ALTRO_UNROLL
    for (int i = 0; i < r; ++i) {
      double acc = A[i * k] * B[j * k];
      ALTRO_UNROLL
      for (int l = 1; l < k; ++l) acc += A[l + i * k] * B[l + j * k];
      C[i + j * r] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Dynamics functors.  THE MODEL CONCEPT every kernel template is written against (and what a
// plug-in supplied at run time has to provide — csrc/modules.inl, plugins/*.cuh):
//
//   struct Model {
//     static constexpr int n, m;                       // state and control dimensions
//     static constexpr bool kDiscrete;                 // false: continuous model, discretised by RK4 here
//     static constexpr bool kStage3RepeatsStage2;      // true only if RK4 stages 2 and 3 provably coincide
//     // xdot = f(x, u); P = the problem's model parameters (altro_b200_problem_set_model)
//     static __device__ void eval(const double* P, const double* x, const double* u, double* xd);
//     // dense Jacobians, column-major: A = df/dx (n x n), B = df/du (n x m); every entry written
//     static __device__ void jac(const double* P, const double* x, const double* u, double* A, double* B);
//     // optional: both at once when they share work
//     static __device__ void eval_jac(const double* P, const double* x, const double* u, double* xd, double* A, double* B);
//   };
//
// This is the device-side counterpart of the reference's ContinuousDynamics::Evaluate / Jacobian
// (altro/problem/dynamics.hpp:59-99 there).  Everything is a compile-time-sized register array.
// ------------------------------------------------------------------------------------------
struct Unicycle {  // examples/unicycle.cpp:12-33
  static constexpr int n = 3, m = 2;
  static constexpr bool kDiscrete = false;
  // xdot depends on the state only through x[2] and xdot[2] = u[1] does not depend on the state:
  // RK4 stages 2 and 3 are evaluated at bitwise the same x[2], so k3 == k2 (and the stage
  // Jacobians coincide) — one sin/cos pair less per step with identical results
  static constexpr bool kStage3RepeatsStage2 = true;
  // value and Jacobian at the same point share one sin/cos evaluation (same results)
  static __device__ __forceinline__ void eval_jac(const double*, const double* x, const double* u,
                                                  double* xd, double* A, double* B) {
    double s, c;
    sincos(x[2], &s, &c);
    xd[0] = u[0] * c;
    xd[1] = u[0] * s;
    xd[2] = u[1];
    ALTRO_UNROLL
    for (int i = 0; i < 9; ++i) A[i] = 0.0;
    ALTRO_UNROLL
    for (int i = 0; i < 6; ++i) B[i] = 0.0;
    A[0 + 2 * 3] = -u[0] * s;
    A[1 + 2 * 3] = u[0] * c;
    B[0 + 0 * 3] = c;
    B[1 + 0 * 3] = s;
    B[2 + 1 * 3] = 1.0;
  }
  static __device__ __forceinline__ void eval(const double*, const double* x, const double* u,
                                              double* xd) {
    double s, c;
    sincos(x[2], &s, &c);
    xd[0] = u[0] * c;
    xd[1] = u[0] * s;
    xd[2] = u[1];
  }
  static __device__ __forceinline__ void jac(const double*, const double* x, const double* u,
                                             double* A, double* B) {
    double s, c;
    sincos(x[2], &s, &c);
    ALTRO_UNROLL
    for (int i = 0; i < 9; ++i) A[i] = 0.0;
    ALTRO_UNROLL
    for (int i = 0; i < 6; ++i) B[i] = 0.0;
    A[0 + 2 * 3] = -u[0] * s;
    A[1 + 2 * 3] = u[0] * c;
    B[0 + 0 * 3] = c;
    B[1 + 0 * 3] = s;
    B[2 + 1 * 3] = 1.0;
  }
};

template <int dof>
struct TripleIntegrator {  // examples/triple_integrator.cpp:9-33
  static constexpr int n = 3 * dof, m = dof;
  static constexpr bool kDiscrete = false;
  static constexpr bool kStage3RepeatsStage2 = false;
  static __device__ __forceinline__ void eval(const double*, const double* x, const double* u,
                                              double* xd) {
    ALTRO_UNROLL
    for (int i = 0; i < dof; ++i) {
      xd[i] = x[i + dof];
      xd[i + dof] = x[i + 2 * dof];
      xd[i + 2 * dof] = u[i];
    }
  }
  static __device__ __forceinline__ void jac(const double*, const double*, const double*,
                                             double* A, double* B) {
    ALTRO_UNROLL
    for (int i = 0; i < n * n; ++i) A[i] = 0.0;
    ALTRO_UNROLL
    for (int i = 0; i < n * m; ++i) B[i] = 0.0;
    ALTRO_UNROLL
    for (int i = 0; i < dof; ++i) {
      A[i + (i + dof) * n] = 1.0;
      A[(i + dof) + (i + 2 * dof) * n] = 1.0;
      B[(i + 2 * dof) + i * n] = 1.0;
    }
  }
};

// value + Jacobian of a continuous model at one point; models may provide a fused eval_jac
template <class M, class = void>
struct EvalJac {
  static __device__ __forceinline__ void run(const double* P, const double* x, const double* u, double* xd,
                                             double* A, double* B) {
    M::eval(P, x, u, xd);
    M::jac(P, x, u, A, B);
  }
};
template <class M>
struct EvalJac<M, decltype((void)&M::eval_jac)> {
  static __device__ __forceinline__ void run(const double* P, const double* x, const double* u, double* xd,
                                             double* A, double* B) {
    M::eval_jac(P, x, u, xd, A, B);
  }
};

// RungeKutta4::Integrate, altro/problem/integration.hpp:124-131.  h is float, promoted (Q1).
template <class M>
__device__ __forceinline__ void rk4_step(const double* P, const double* x, const double* u,
                                         float hf, double* xn) {
  constexpr int n = M::n;
  const double h = static_cast<double>(hf);
  double k1[n], k2[n], k3[n], k4[n], xt[n];
  M::eval(P, x, u, k1);
  ALTRO_UNROLL
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * 0.5 * h;
  M::eval(P, xt, u, k2);
  if (M::kStage3RepeatsStage2) {
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) k3[i] = k2[i];
  } else {
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) xt[i] = x[i] + k2[i] * 0.5 * h;
    M::eval(P, xt, u, k3);
  }
  ALTRO_UNROLL
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k3[i] * h;
  M::eval(P, xt, u, k4);
  const double r6 = 1.0 / 6.0;
  ALTRO_UNROLL
  for (int i = 0; i < n; ++i) xn[i] = x[i] + div_by(h * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]), 6.0, r6);
}

// RungeKutta4::Jacobian, altro/problem/integration.hpp:132-167.
template <class M>
__device__ __forceinline__ void rk4_jacobian(const double* P, const double* x, const double* u,
                                             float hf, double* A, double* B) {
  constexpr int n = M::n, m = M::m;
  const double h = static_cast<double>(hf);
  double k1[n], k2[n], k3[n], xt[n];
  double As[n * n], Bs[n * m], dA[n * n], dB[n * m], T[n * n], T2[n * n], TB[n * m];
  // stage 0
  EvalJac<M>::run(P, x, u, k1, As, Bs);
  ALTRO_UNROLL
  for (int i = 0; i < n * n; ++i) { dA[i] = As[i] * h; A[i] = dA[i]; }
  ALTRO_UNROLL
  for (int i = 0; i < n * m; ++i) { dB[i] = Bs[i] * h; B[i] = dB[i]; }
  // stage 1: x + 0.5*k1*h
  ALTRO_UNROLL
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * 0.5 * h;
  EvalJac<M>::run(P, xt, u, k2, As, Bs);
  ALTRO_UNROLL
  for (int j = 0; j < n; ++j)
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) T[i + j * n] = (i == j ? 1.0 : 0.0) + 0.5 * dA[i + j * n];
  matmul<n, n, n>(As, T, T2);
  ALTRO_UNROLL
  for (int i = 0; i < n * n; ++i) T[i] = 0.5 * As[i];
  matmul<n, n, m>(T, dB, TB);
  ALTRO_UNROLL
  for (int i = 0; i < n * n; ++i) { dA[i] = T2[i] * h; A[i] = A[i] + 2 * dA[i]; }
  ALTRO_UNROLL
  for (int i = 0; i < n * m; ++i) { dB[i] = Bs[i] * h + TB[i] * h; B[i] = B[i] + 2 * dB[i]; }
  // stage 2: x + 0.5*k2*h
  if (M::kStage3RepeatsStage2) {  // same point as stage 1: As, Bs stay, k3 = k2
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) k3[i] = k2[i];
  } else {
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) xt[i] = x[i] + k2[i] * 0.5 * h;
    EvalJac<M>::run(P, xt, u, k3, As, Bs);
  }
  ALTRO_UNROLL
  for (int j = 0; j < n; ++j)
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) T[i + j * n] = (i == j ? 1.0 : 0.0) + 0.5 * dA[i + j * n];
  matmul<n, n, n>(As, T, T2);
  ALTRO_UNROLL
  for (int i = 0; i < n * n; ++i) T[i] = 0.5 * As[i];
  matmul<n, n, m>(T, dB, TB);
  ALTRO_UNROLL
  for (int i = 0; i < n * n; ++i) { dA[i] = T2[i] * h; A[i] = A[i] + 2 * dA[i]; }
  ALTRO_UNROLL
  for (int i = 0; i < n * m; ++i) { dB[i] = Bs[i] * h + TB[i] * h; B[i] = B[i] + 2 * dB[i]; }
  // stage 3: x + k3*h
  ALTRO_UNROLL
  for (int i = 0; i < n; ++i) xt[i] = x[i] + k3[i] * h;
  M::jac(P, xt, u, As, Bs);
  ALTRO_UNROLL
  for (int j = 0; j < n; ++j)
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) T[i + j * n] = (i == j ? 1.0 : 0.0) + dA[i + j * n];
  matmul<n, n, n>(As, T, T2);
  matmul<n, n, m>(As, dB, TB);
  ALTRO_UNROLL
  for (int j = 0; j < n; ++j)
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) {
      const int q = i + j * n;
      A[q] = (i == j ? 1.0 : 0.0) + div_by(A[q] + T2[q] * h, 6.0, 1.0 / 6.0);
    }
  ALTRO_UNROLL
  for (int i = 0; i < n * m; ++i) B[i] = div_by(B[i] + (Bs[i] * h + TB[i] * h), 6.0, 1.0 / 6.0);
}

// ------------------------------------------------------------------------------------------
// QuadraticCost (examples/quadratic_cost.cpp:8-28).  C = [Q (n*n), R (m*m), H (n*m), q, r, c].
// ------------------------------------------------------------------------------------------
// shape of the cost record (CostShape), stored as a double in the slot after c
template <int n, int m>
__device__ __forceinline__ int quad_shape(const double* C) {
  return static_cast<int>(C[n * n + m * m + n * m + n + m + 1]);
}

template <int n, int m>
__device__ __forceinline__ double quad_eval(const double* C, const double* x, const double* u) {
  const double *Q = C, *R = C + n * n, *H = R + m * m, *q = H + n * m, *r = q + n;
  if (quad_shape<n, m>(C) == kCostDiagonal) {
    // the dense sums below with their exact-zero products left out: Qx_i = Q_ii x_i, Hu = 0, Ru_i = R_ii u_i
    double xQx = x[0] * (Q[0] * x[0]), qx = q[0] * x[0];
    ALTRO_UNROLL
    for (int i = 1; i < n; ++i) {
      xQx += x[i] * (Q[i + i * n] * x[i]);
      qx += q[i] * x[i];
    }
    double uRu = u[0] * (R[0] * u[0]), ru = r[0] * u[0];
    ALTRO_UNROLL
    for (int i = 1; i < m; ++i) {
      uRu += u[i] * (R[i + i * m] * u[i]);
      ru += r[i] * u[i];
    }
    return 0.5 * xQx + 0.5 * uRu + qx + ru + r[m];
  }
  double Qx[n], Hu[n], Ru[m];
  ALTRO_UNROLL
  for (int i = 0; i < n; ++i) {
    double a = Q[i] * x[0];
    ALTRO_UNROLL
    for (int j = 1; j < n; ++j) a += Q[i + j * n] * x[j];
    Qx[i] = a;
    double b = H[i] * u[0];
    ALTRO_UNROLL
    for (int j = 1; j < m; ++j) b += H[i + j * n] * u[j];
    Hu[i] = b;
  }
  ALTRO_UNROLL
  for (int i = 0; i < m; ++i) {
    double a = R[i] * u[0];
    ALTRO_UNROLL
    for (int j = 1; j < m; ++j) a += R[i + j * m] * u[j];
    Ru[i] = a;
  }
  double xQx = x[0] * Qx[0], xHu = x[0] * Hu[0], qx = q[0] * x[0];
  ALTRO_UNROLL
  for (int i = 1; i < n; ++i) {
    xQx += x[i] * Qx[i];
    xHu += x[i] * Hu[i];
    qx += q[i] * x[i];
  }
  double uRu = u[0] * Ru[0], ru = r[0] * u[0];
  ALTRO_UNROLL
  for (int i = 1; i < m; ++i) {
    uRu += u[i] * Ru[i];
    ru += r[i] * u[i];
  }
  return 0.5 * xQx + xHu + 0.5 * uRu + qx + ru + r[m];
}

template <int n, int m>
__device__ __forceinline__ void quad_gradient(const double* C, const double* x, const double* u,
                                              double* dx, double* du) {
  const double *Q = C, *R = C + n * n, *H = R + m * m, *q = H + n * m, *r = q + n;
  if (quad_shape<n, m>(C) == kCostDiagonal) {
    // product and sum rounded separately, as in the dense path (where the product is the tail of
    // a sum of exact zeros and cannot fuse with the addition of q_i)
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) dx[i] = __dadd_rn(__dmul_rn(Q[i + i * n], x[i]), q[i]);
    ALTRO_UNROLL
    for (int i = 0; i < m; ++i) du[i] = __dadd_rn(__dmul_rn(R[i + i * m], u[i]), r[i]);
    return;
  }
  ALTRO_UNROLL
  for (int i = 0; i < n; ++i) {
    double a = Q[i] * x[0];
    ALTRO_UNROLL
    for (int j = 1; j < n; ++j) a += Q[i + j * n] * x[j];
    double b = H[i] * u[0];
    ALTRO_UNROLL
    for (int j = 1; j < m; ++j) b += H[i + j * n] * u[j];
    dx[i] = a + q[i] + b;
  }
  ALTRO_UNROLL
  for (int i = 0; i < m; ++i) {
    double a = R[i] * u[0];
    ALTRO_UNROLL
    for (int j = 1; j < m; ++j) a += R[i + j * m] * u[j];
    double b = H[i * n] * x[0];
    ALTRO_UNROLL
    for (int j = 1; j < n; ++j) b += H[j + i * n] * x[j];
    du[i] = a + r[i] + b;
  }
}

// ------------------------------------------------------------------------------------------
// Augmented-Lagrangian terms of one knot.  lam points at the lane's dual column of the knot
// (row r at lam[r*32]); rho is the scalar penalty (constraint_values.hpp:112, Q9).
//
// Reference: ConstraintValues::AugLag/AugLagGradient/AugLagHessian
// (altro/constraints/constraint_values.hpp:111-177), cones constraint.hpp:28-122, summed per
// constraint by ALCost (altro/augmented_lagrangian/al_cost.hpp:264-308).
// ------------------------------------------------------------------------------------------

// dynamic-index read of a small register array without local memory
template <int len>
__device__ __forceinline__ double pick(const double* v, int j) {
  double r = v[0];
  ALTRO_UNROLL
  for (int i = 1; i < len; ++i) r = (j == i) ? v[i] : r;
  return r;
}
template <int len>
__device__ __forceinline__ void add_at(double* v, int j, double val) {
  ALTRO_UNROLL
  for (int i = 0; i < len; ++i) v[i] = (j == i) ? (v[i] + val) : v[i];
}

template <int n, int m>
__device__ __forceinline__ double con_row_fast(const ConBlock& b, int i, const double* x,
                                               const double* u) {
  switch (b.kind) {
    case kGoal:
      return pick<n>(x, i) - b.a[i];
    case kControlBound: {
      const double uj = pick<m>(u, b.idx[i]);
      return (i < b.nl) ? (b.a[i] - uj) : (uj - b.a[i]);
    }
    default: {
      const double dx = pick<n>(x, b.xi) - b.a[i], dy = pick<n>(x, b.yi) - b.b[i];
      return -(dx * dx + dy * dy - b.c[i]);
    }
  }
}

// Pi_{K*} for the negative orthant: min(0, x) (constraint.hpp:98-104).  Same value as
// fmin(0.0, x) for every x including NaN, in 3 instructions instead of 8.
__device__ __forceinline__ double neg_part(double x) { return x < 0.0 ? x : 0.0; }

// Rows of one constraint block, four at a time.  The per-row control flow of an interpreter
// (load kind -> branch -> load operands -> compute) serialises on shared-memory latency; here
// the kind is resolved once per block and a chunk's operand loads and constraint values are
// independent, so they overlap.  f(c, ic, ok) receives the values c[q] of rows ic[q] = r0 + q;
// rows past the block's end have ok[q] = false and repeat the last row (to be masked by f).
constexpr int kChunk = 4;

// scalar fields of a constraint block, unpacked from ConSet::hdr (one 16-byte load)
struct BlockHdr {
  int kind, p, row0, nl, xi, yi, shape;
  bool eq;
  __device__ __forceinline__ explicit BlockHdr(const int4& h)
      : kind(h.x & 0xff), p(h.x >> 16), row0(h.y), nl(h.z), xi(h.w & 0xff), yi(h.w >> 8),
        shape((h.x >> 12) & 0xf), eq(((h.x >> 8) & 1) != 0) {}
};

// penalty and the reciprocal used by `/ (2 rho)` (constraint_values.hpp:118), computed once per sweep
struct AlPen {
  double rho, two_rho, r_two_rho;
  __device__ __forceinline__ explicit AlPen(double r) : rho(r), two_rho(2 * r), r_two_rho(1.0 / (2 * r)) {}
};

template <int n, int m, class F>
__device__ __forceinline__ void for_row_chunks(const BlockHdr& hd, const ConBlock& b, const double* x,
                                               const double* u, F&& f) {
  const int p = hd.p;
  if (hd.shape == kShapeBoundsFull) {
    // every control has a finite lower and upper bound: rows j < m are lb_j - u_j, rows m + j are
    // u_j - ub_j; everything but the bound values is known at compile time
    ALTRO_UNROLL
    for (int r0 = 0; r0 < 2 * m; r0 += kChunk) {
      double c[kChunk];
      int ic[kChunk];
      bool ok[kChunk];
      ALTRO_UNROLL
      for (int q = 0; q < kChunk; ++q) {
        const int r = r0 + q;
        ok[q] = r < 2 * m;
        ic[q] = ok[q] ? r : 2 * m - 1;
        const double uj = u[ic[q] % m];
        c[q] = (ic[q] < m) ? (b.a[ic[q]] - uj) : (uj - b.a[ic[q]]);
      }
      f(c, ic, ok);
    }
  } else if (hd.shape == kShapeCirclesOneChunk) {
    const double px = pick<n>(x, hd.xi), py = pick<n>(x, hd.yi);
    double c[kChunk];
    int ic[kChunk];
    bool ok[kChunk];
    ALTRO_UNROLL
    for (int q = 0; q < kChunk; ++q) {
      ok[q] = q < p;
      ic[q] = ok[q] ? q : p - 1;
      const double dx = px - b.a[q], dy = py - b.b[q];  // rows past p: zero-filled, masked by f
      c[q] = -(dx * dx + dy * dy - b.c[q]);
    }
    f(c, ic, ok);
  } else if (hd.kind == kCircle) {
    const double px = pick<n>(x, hd.xi), py = pick<n>(x, hd.yi);
    for (int r0 = 0; r0 < p; r0 += kChunk) {
      double c[kChunk];
      int ic[kChunk];
      bool ok[kChunk];
      ALTRO_UNROLL
      for (int q = 0; q < kChunk; ++q) {
        ok[q] = r0 + q < p;
        ic[q] = ok[q] ? r0 + q : p - 1;
        const double dx = px - b.a[ic[q]], dy = py - b.b[ic[q]];
        c[q] = -(dx * dx + dy * dy - b.c[ic[q]]);  // obstacle_constraints.hpp:107-110
      }
      f(c, ic, ok);
    }
  } else if (hd.kind == kControlBound) {
    const int nl = hd.nl;
    for (int r0 = 0; r0 < p; r0 += kChunk) {
      double c[kChunk];
      int ic[kChunk];
      bool ok[kChunk];
      ALTRO_UNROLL
      for (int q = 0; q < kChunk; ++q) {
        ok[q] = r0 + q < p;
        ic[q] = ok[q] ? r0 + q : p - 1;
        const double uj = pick<m>(u, b.idx[ic[q]]);
        c[q] = (ic[q] < nl) ? (b.a[ic[q]] - uj) : (uj - b.a[ic[q]]);  // basic_constraints.hpp:110-118
      }
      f(c, ic, ok);
    }
  } else {  // kGoal: c = x - xf (basic_constraints.hpp:31)
    for (int r0 = 0; r0 < p; r0 += kChunk) {
      double c[kChunk];
      int ic[kChunk];
      bool ok[kChunk];
      ALTRO_UNROLL
      for (int q = 0; q < kChunk; ++q) {
        ok[q] = r0 + q < p;
        ic[q] = ok[q] ? r0 + q : p - 1;
        c[q] = pick<n>(x, ic[q]) - b.a[ic[q]];
      }
      f(c, ic, ok);
    }
  }
}

// Adds the AL value of every constraint of a knot to J, one constraint at a time
// (al_cost.hpp:266-272); also returns the max violation |c - Pi_K(c)|_inf
// (constraint_values.hpp:216-221).
template <int n, int m, int LS>
__device__ __forceinline__ double al_value(const ConSet& cs, const double* x, const double* u,
                                           const double* lam, const AlPen& pen, double J,
                                           double* viol) {
  double v = 0.0;
  const double rho = pen.rho;
  const int nblocks = cs.nblocks;
  int4 hraw[kMaxBlocks];
  ALTRO_UNROLL
  for (int bi = 0; bi < kMaxBlocks; ++bi) hraw[bi] = cs.hdr[bi];  // all block headers at once
  ALTRO_UNROLL
  for (int bi = 0; bi < kMaxBlocks; ++bi) {
    if (bi < nblocks) {
      const BlockHdr hd(hraw[bi]);
      const ConBlock& b = cs.blk[bi];
      const bool eq = hd.eq;
      const double* lb = lam + hd.row0 * LS;
      double sa = 0.0, sb = 0.0;
      for_row_chunks<n, m>(hd, b, x, u, [&](const double* c, const int* ic, const bool* ok) {
        double l[kChunk], lp[kChunk];
        ALTRO_UNROLL
        for (int q = 0; q < kChunk; ++q) {
          l[q] = lb[ic[q] * LS];
          const double arg = l[q] - rho * c[q];
          lp[q] = eq ? arg : neg_part(arg);
        }
        ALTRO_UNROLL
        for (int q = 0; q < kChunk; ++q) {
          if (ok[q]) {
            sa += lp[q] * lp[q];
            sb += l[q] * l[q];
            v = fmax(v, eq ? fabs(c[q]) : fabs(c[q] - neg_part(c[q])));
          }
        }
      });
      double Jb = sa - sb;
      Jb = div_by(Jb, pen.two_rho, pen.r_two_rho);
      J += Jb;
    }
  }
  if (viol) *viol = v;
  return J;
}

// Adds the AL gradient and Gauss-Newton Hessian of every constraint of the knot to the cost
// expansion (which already holds the QuadraticCost terms).  Per constraint the reference forms
// dx = -(dPi dc)^T Pi(lam - rho c) and dxdx = rho (dPi dc)^T (dPi dc) as matrix products (sums
// over the rows in order) and ALCost adds them to the running expansion
// (constraint_values.hpp:131-177, al_cost.hpp:280-308); here the products are accumulated
// row by row in the same order on the few entries the constraint's Jacobian touches.
template <int n, int m, int LS>
__device__ __forceinline__ void al_expansion(const ConSet& cs, const double* x, const double* u,
                                             const double* lam, double rho, double* lxx,
                                             double* lxu, double* luu, double* lx, double* lu) {
  (void)lxu;  // none of the device-capable constraints couples x and u
  const int nblocks = cs.nblocks;
  int4 hraw[kMaxBlocks];
  ALTRO_UNROLL
  for (int bi = 0; bi < kMaxBlocks; ++bi) hraw[bi] = cs.hdr[bi];
  ALTRO_UNROLL
  for (int bi = 0; bi < kMaxBlocks; ++bi) {
    if (bi >= nblocks) continue;
    const BlockHdr hd(hraw[bi]);
    const ConBlock& b = cs.blk[bi];
    const double* lb = lam + hd.row0 * LS;
    if (hd.kind == kGoal) {
      // J = [I | 0], identity dual cone: dx_i = -(lam_i - rho c_i), dxdx_ii = rho
      ALTRO_UNROLL
      for (int i = 0; i < n; ++i) {
        if (i < hd.p) {
          const double c = x[i] - b.a[i];
          const double lp = lb[i * LS] - rho * c;
          lx[i] += (-1.0) * lp;
          lxx[i + i * n] += (rho * 1.0) * 1.0;
        }
      }
    } else if (hd.kind == kControlBound) {
      // rows [0,nl): c = lb_j - u_j, J = -e_j ; rows [nl,nl+nu): c = u_j - ub_j, J = +e_j.
      // Per control j the block contributes its lower-row term first, then its upper-row term.
      double gdu[m], gduu[m];
      ALTRO_UNROLL
      for (int j = 0; j < m; ++j) { gdu[j] = 0.0; gduu[j] = 0.0; }
      const int nl = hd.nl;
      for_row_chunks<n, m>(hd, b, x, u, [&](const double* c, const int* ic, const bool* ok) {
        double g[kChunk], hh[kChunk];
        int jj[kChunk];
        ALTRO_UNROLL
        for (int q = 0; q < kChunk; ++q) {
          jj[q] = ok[q] ? b.idx[ic[q]] : -1;
          const double arg = lb[ic[q] * LS] - rho * c[q];
          const double lp = neg_part(arg);
          const double act = arg > 0 ? 0.0 : 1.0;               // constraint.hpp:112 (Q12)
          const double jp = act * (ic[q] < nl ? -1.0 : 1.0);    // proj_jac * jac entry
          g[q] = (-jp) * lp;
          hh[q] = (rho * jp) * jp;
        }
        ALTRO_UNROLL
        for (int q = 0; q < kChunk; ++q) {
          ALTRO_UNROLL
          for (int j = 0; j < m; ++j) {
            if (jj[q] == j) {
              gdu[j] += g[q];
              gduu[j] += hh[q];
            }
          }
        }
      });
      ALTRO_UNROLL
      for (int j = 0; j < m; ++j) {
        lu[j] += gdu[j];
        luu[j + j * m] += gduu[j];
      }
    } else {
      // circles: J(i, 0) = 2(cx - px), J(i, 1) = 2(cy - py) — columns 0 and 1 as the
      // reference writes them (obstacle_constraints.hpp:117-118).
      static_assert(n >= 2, "circle constraints need two position states");
      const double px = pick<n>(x, hd.xi), py = pick<n>(x, hd.yi);
      double g0 = 0.0, g1 = 0.0, h00 = 0.0, h01 = 0.0, h10 = 0.0, h11 = 0.0;
      for_row_chunks<n, m>(hd, b, x, u, [&](const double* c, const int* ic, const bool* ok) {
        double j0[kChunk], j1[kChunk], t0[kChunk], t1[kChunk], r0[kChunk], r1[kChunk];
        ALTRO_UNROLL
        for (int q = 0; q < kChunk; ++q) {
          const double arg = lb[ic[q] * LS] - rho * c[q];
          const double lp = neg_part(arg);
          const double act = (ok[q] && !(arg > 0)) ? 1.0 : 0.0;
          j0[q] = act * (2 * (b.a[ic[q]] - px));
          j1[q] = act * (2 * (b.b[ic[q]] - py));
          t0[q] = (-j0[q]) * lp;
          t1[q] = (-j1[q]) * lp;
          r0[q] = rho * j0[q];
          r1[q] = rho * j1[q];
        }
        ALTRO_UNROLL
        for (int q = 0; q < kChunk; ++q) {
          if (ok[q]) {
            g0 += t0[q];
            g1 += t1[q];
            h00 += r0[q] * j0[q];
            h01 += r0[q] * j1[q];
            h10 += r1[q] * j0[q];
            h11 += r1[q] * j1[q];
          }
        }
      });
      lx[0] += g0;
      lx[1] += g1;
      lxx[0 + 0 * n] += h00;
      lxx[0 + 1 * n] += h01;
      lxx[1 + 0 * n] += h10;
      lxx[1 + 1 * n] += h11;
    }
  }
}

// ALCost::Evaluate (al_cost.hpp:264-274) for knot k.
template <int n, int m, int LS>
__device__ __forceinline__ double knot_cost(const Desc& D, int k, const double* x,
                                            const double* u, const double* lam, const AlPen& pen,
                                            double* viol) {
  const double J = quad_eval<n, m>(D.cost(k), x, u);
  return al_value<n, m, LS>(D.conset(k), x, u, lam, pen, J, viol);
}

// Riccati step of the backward pass (knot_point_function_type.hpp:149-230) ----------------
// In:  A,B, cost expansion, P,p of knot k+1, regularisation.  Out: K,d, P,p of knot k, dV.
// Returns false when the Cholesky factorisation of Quu_reg hits a pivot <= 0 (Q19).
template <int n, int m>
__device__ __forceinline__ bool riccati_step(const double* A, const double* B, const double* lxx,
                                             const double* lxu, const double* luu,
                                             const double* lx, const double* lu, double* P,
                                             double* p, double reg, double* K, double* d,
                                             double* dV0, double* dV1) {
  double Qxx[n * n], Qxu[n * m], Quu[m * m], Qx[n], Qu[m];
  {
    double AtP[n * n], BtP[m * n], T1[n * n], T2[n * m], T3[m * m], v[n], w[m];
    mattmul<n, n, n>(A, P, AtP);          // A^T P
    matmul<n, n, n>(AtP, A, T1);          // (A^T P) A
    matmul<n, n, m>(AtP, B, T2);          // (A^T P) B
    mattmul<m, n, n>(B, P, BtP);          // B^T P
    matmul<m, n, m>(BtP, B, T3);          // (B^T P) B
    mattmul<n, n, 1>(A, p, v);
    mattmul<m, n, 1>(B, p, w);
    ALTRO_UNROLL
    for (int i = 0; i < n * n; ++i) Qxx[i] = lxx[i] + T1[i];
    ALTRO_UNROLL
    for (int i = 0; i < n * m; ++i) Qxu[i] = lxu[i] + T2[i];
    ALTRO_UNROLL
    for (int i = 0; i < m * m; ++i) Quu[i] = luu[i] + T3[i];
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) Qx[i] = lx[i] + v[i];
    ALTRO_UNROLL
    for (int i = 0; i < m; ++i) Qu[i] = lu[i] + w[i];
  }
  // RegularizeActionValue :175-186 + Eigen::LLT (lower, unblocked)
  double L[m * m], rL[m];
  ALTRO_UNROLL
  for (int j = 0; j < m; ++j)
    ALTRO_UNROLL
    for (int i = 0; i < m; ++i) L[i + j * m] = Quu[i + j * m] + (i == j ? 1.0 : 0.0) * reg;
  bool ok = true;
  ALTRO_UNROLL
  for (int k = 0; k < m; ++k) {
    double xk = L[k + k * m];
    if (k > 0) {
      double sq = 0.0;
      ALTRO_UNROLL
      for (int j = 0; j < k; ++j) sq += L[k + j * m] * L[k + j * m];
      xk -= sq;
    }
    if (xk <= 0.0) ok = false;
    xk = sqrt(xk);
    L[k + k * m] = xk;
    rL[k] = 1.0 / xk;
    ALTRO_UNROLL
    for (int i = k + 1; i < m; ++i) {
      double a = L[i + k * m];
      if (k > 0) {
        double dot = 0.0;
        ALTRO_UNROLL
        for (int j = 0; j < k; ++j) dot += L[i + j * m] * L[k + j * m];
        a -= dot;
      }
      L[i + k * m] = div_by(a, xk, rL[k]);
    }
  }
  if (!ok) return false;
  // K = -(L L^T)^-1 Qxu^T ; d = -(L L^T)^-1 Qu   (CalcGains :197-211)
  ALTRO_UNROLL
  for (int c = 0; c <= n; ++c) {
    double bvec[m];
    ALTRO_UNROLL
    for (int i = 0; i < m; ++i) bvec[i] = (c < n) ? Qxu[(c < n ? c : 0) + i * n] : Qu[i];
    ALTRO_UNROLL
    for (int i = 0; i < m; ++i) {
      double s = bvec[i];
      ALTRO_UNROLL
      for (int j = 0; j < i; ++j) s -= L[i + j * m] * bvec[j];
      bvec[i] = div_by(s, L[i + i * m], rL[i]);
    }
    ALTRO_UNROLL
    for (int i = m - 1; i >= 0; --i) {
      double s = bvec[i];
      ALTRO_UNROLL
      for (int j = i + 1; j < m; ++j) s -= L[j + i * m] * bvec[j];
      bvec[i] = div_by(s, L[i + i * m], rL[i]);
    }
    ALTRO_UNROLL
    for (int i = 0; i < m; ++i) {
      if (c < n) K[i + (c < n ? c : 0) * m] = bvec[i] * -1;
      else d[i] = bvec[i] * -1;
    }
  }
  // CalcCostToGo :220-230 — unregularised Q (Q3)
  {
    double KtQuu[n * m], v1[n], v2[n], v3[n], T1[n * n], T2[n * n], T3[n * n], Quud[m];
    mattmul<n, m, m>(K, Quu, KtQuu);
    matmul<n, m, 1>(KtQuu, d, v1);
    mattmul<n, m, 1>(K, Qu, v2);
    matmul<n, m, 1>(Qxu, d, v3);
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) p[i] = Qx[i] + v1[i] + v2[i] + v3[i];
    matmul<n, m, n>(KtQuu, K, T1);
    ALTRO_UNROLL
    for (int j = 0; j < n; ++j)
      ALTRO_UNROLL
      for (int i = 0; i < n; ++i) {
        double acc = K[i * m] * Qxu[j];
        ALTRO_UNROLL
        for (int l = 1; l < m; ++l) acc += K[l + i * m] * Qxu[j + l * n];
        T2[i + j * n] = acc;
      }
    matmul<n, m, n>(Qxu, K, T3);
    ALTRO_UNROLL
    for (int i = 0; i < n * n; ++i) P[i] = Qxx[i] + T1[i] + T2[i] + T3[i];
    double a = d[0] * Qu[0];
    ALTRO_UNROLL
    for (int i = 1; i < m; ++i) a += d[i] * Qu[i];
    matmul<m, m, 1>(Quu, d, Quud);
    double bq = d[0] * Quud[0];
    ALTRO_UNROLL
    for (int i = 1; i < m; ++i) bq += d[i] * Quud[i];
    *dV0 += a;
    *dV1 += 0.5 * bq;
  }
  return true;
}

// IncreaseRegularization / DecreaseRegularization, ilqr.hpp:770-786 (Q4)
__device__ __forceinline__ void increase_reg(const DevOptions& o, double& reg, double& dreg) {
  dreg = fmax(dreg * o.bp_reg_increase_factor, o.bp_reg_increase_factor);
  reg = fmax(reg * dreg, o.bp_reg_min);
  reg = fmin(reg, o.bp_reg_max);
}
__device__ __forceinline__ void decrease_reg(const DevOptions& o, double& reg, double& dreg) {
  dreg = fmin(dreg / o.bp_reg_increase_factor, 1 / o.bp_reg_increase_factor);
  reg = fmax(reg * dreg, o.bp_reg_min);
  reg = fmin(reg, o.bp_reg_max);
}


}  // namespace altro_b200
