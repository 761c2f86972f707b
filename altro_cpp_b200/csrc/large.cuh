// large.cuh — the large-state path (BASELINE config C5: random LQR, n = 32, m = 8): one problem
// instance per CTA, every matrix of the Riccati step resident in shared memory, every matrix
// element owned by one thread that accumulates its inner product in the reference's order
// (SURVEY.md Q20), so the results agree with the per-lane path / the oracle to rounding.
//
// Supported here: discrete LTI dynamics x+ = A x + B u (ALTRO_B200_MODEL_LINEAR), dense
// QuadraticCost, no constraints (the AL outer loop then terminates after one iLQR solve, exactly
// like the reference does for an unconstrained Problem).  fp64 FMA pipe; the DMMA tensor path
// for A'PA is a follow-up.
#pragma once

#include "device.cuh"

namespace altro_b200 {

constexpr int kLargeThreads = 256;

// C (r x c) = op(A) * B ; all column-major in shared memory; each thread owns elements
// e = tid, tid + T, ... and sums the inner index ascending.
// lda = leading dimension of A (k for a transposed k x r operand unless padded)
template <int r, int k, int c, bool kTransA, int lda>
__device__ __forceinline__ void sm_matmul(const double* A, const double* B, double* C) {
  for (int e = threadIdx.x; e < r * c; e += kLargeThreads) {
    const int i = e % r, j = e / r;
    double acc = (kTransA ? A[i * lda] : A[i]) * B[j * k];
    for (int l = 1; l < k; ++l) acc += (kTransA ? A[l + i * lda] : A[i + l * lda]) * B[l + j * k];
    C[e] = acc;
  }
}

// A, B and K are also read transposed (A'P, B'P, K'Quu, K'Qxu'): with the natural leading
// dimension consecutive lanes would hit one bank (stride 32 or 8 doubles), so they are stored
// with leading dimension n+1 / m+1.
template <int n, int m>
struct LargeSmem {
  double A[(n + 1) * n], B[(n + 1) * m], P[n * n], T[n * n], Qxx[n * n], Qxu[n * m], Quu[m * m], L[m * m];
  double BtP[m * n], K[(m + 1) * n], KtQuu[n * m];
  double x[n], xr[n], u[m], p[n], Qx[n], Qu[m], d[m], lx[n], lu[m], v1[n], v2[n], v3[n], tmp[n];
  double scal[16];
  int iscal[8];
};

// One whole AL-iLQR / iLQR solve per CTA (mode as in k_solve).
template <int n, int m>
__global__ void __launch_bounds__(kLargeThreads) k_solve_large(SolverParams P, int mode) {
  extern __shared__ __align__(128) char smem_raw[];
  LargeSmem<n, m>& S = *reinterpret_cast<LargeSmem<n, m>*>(smem_raw);
  const int b = blockIdx.x, tid = threadIdx.x;
  if (b >= P.B) return;
  constexpr int nz = n + m, nkd = m * n + m, LA = n + 1, LK = m + 1;
  const Desc D(P.blob);  // read through L1/L2: shared by all CTAs
  const DevOptions& o = P.opt;
  const int N = P.N;
  const double* mp = D.params();
  auto zptr = [&](int sel, int k) { return P.Z[sel] + (static_cast<size_t>(b) * (N + 1) + k) * nz; };
  auto kdptr = [&](int k) { return P.KD + (static_cast<size_t>(b) * N + k) * nkd; };
  double& rsc_cost = P.sc[static_cast<size_t>(S_COST) * P.Bp + b];
  for (int e = tid; e < n * n; e += kLargeThreads) S.A[e % n + (e / n) * LA] = mp[e];
  for (int e = tid; e < n * m; e += kLargeThreads) S.B[e % n + (e / n) * LA] = mp[n * n + e];
  __syncthreads();

  // ---- cost of one knot (QuadraticCost::Evaluate), result in S.scal[15]; x, u in S.x, S.u
  auto knot_cost_sm = [&](int k) {
    const double* C = D.cost(k);
    const double *Q = C, *R = C + n * n, *H = R + m * m, *q = H + n * m, *r = q + n;
    if (tid < n) {
      double a = Q[tid] * S.x[0];
      for (int j = 1; j < n; ++j) a += Q[tid + j * n] * S.x[j];
      S.v1[tid] = a;  // Qx
      double h2 = H[tid] * S.u[0];
      for (int j = 1; j < m; ++j) h2 += H[tid + j * n] * S.u[j];
      S.v2[tid] = h2;  // Hu
    } else if (tid >= 64 && tid < 64 + m) {
      const int i = tid - 64;
      double a = R[i] * S.u[0];
      for (int j = 1; j < m; ++j) a += R[i + j * m] * S.u[j];
      S.v3[i] = a;  // Ru
    }
    __syncthreads();
    if (tid == 0) {
      double xQx = S.x[0] * S.v1[0], xHu = S.x[0] * S.v2[0], qx = q[0] * S.x[0];
      for (int i = 1; i < n; ++i) {
        xQx += S.x[i] * S.v1[i];
        xHu += S.x[i] * S.v2[i];
        qx += q[i] * S.x[i];
      }
      double uRu = S.u[0] * S.v3[0], ru = r[0] * S.u[0];
      for (int i = 1; i < m; ++i) {
        uRu += S.u[i] * S.v3[i];
        ru += r[i] * S.u[i];
      }
      S.scal[15] = 0.5 * xQx + xHu + 0.5 * uRu + qx + ru + r[m];
    }
    __syncthreads();
  };

  // ---- forward sweep (kClosed: RolloutClosedLoop(alpha) into zout, else Rollout() in place)
  auto forward = [&](bool closed, int zsel, int zout, double alpha, double* J_out, double* g_out,
                     int* status) -> bool {
    if (tid < n) S.x[tid] = P.X0[static_cast<size_t>(b) * n + tid];
    __syncthreads();
    double J = 0.0, gs = 0.0;
    bool ok = true;
    for (int k = 0; k <= N; ++k) {
      const double* zc = zptr(zsel, k);
      double* zn = zptr(closed ? zout : zsel, k);
      if (k < N && closed) {
        const double* pk = kdptr(k);
        if (tid < n) S.xr[tid] = S.x[tid] - zc[tid];  // dx
        __syncthreads();
        if (tid < m) {
          double acc = pk[tid] * S.xr[0];
          for (int j = 1; j < n; ++j) acc += pk[tid + j * m] * S.xr[j];
          const double dq = pk[m * n + tid];
          S.u[tid] = zc[n + tid] + acc + dq * alpha;
          S.d[tid] = fabs(dq) / (fabs(S.u[tid]) + 1);
        }
      } else if (tid < m) {
        S.u[tid] = (k < N || !closed) ? zc[n + tid] : 0.0;
      }
      __syncthreads();
      if (tid < n) zn[tid] = S.x[tid];
      if (closed && tid < m) zn[n + tid] = S.u[tid];
      knot_cost_sm(k);
      J += S.scal[15];
      if (k < N) {
        if (closed) {
          double g = S.d[0];
          for (int i = 1; i < m; ++i) g = fmax(g, S.d[i]);
          gs += g;
        }
        if (tid < n) {  // x+ = A x + B u
          double acc = S.A[tid] * S.x[0];
          for (int j = 1; j < n; ++j) acc += S.A[tid + j * LA] * S.x[j];
          double accb = S.B[tid] * S.u[0];
          for (int j = 1; j < m; ++j) accb += S.B[tid + j * LA] * S.u[j];
          S.tmp[tid] = acc + accb;
        }
        __syncthreads();
        if (tid < n) S.x[tid] = S.tmp[tid];
        __syncthreads();
        if (closed && o.check_forwardpass_bounds) {  // ilqr.hpp:484-495
          double sx = 0.0, su = 0.0;
          for (int i = 0; i < n; ++i) sx += S.x[i] * S.x[i];
          for (int i = 0; i < m; ++i) su += S.u[i] * S.u[i];
          if (sqrt(sx) > o.state_max) {
            *status = kStateLimit;
            ok = false;
          } else if (sqrt(su) > o.control_max) {
            *status = kControlLimit;
            ok = false;
          }
          if (!ok) break;
        }
      }
    }
    if (ok && closed) *status = kUnsolved;
    *J_out = J;
    *g_out = gs;
    return ok;
  };

  // ---- backward sweep with the regularisation restart loop (ilqr.hpp:385-445)
  auto backward = [&](int zsel, double& reg, double& dreg, double& dV0, double& dV1, int& status,
                      double& gsum) {
    int max_reg_count = 0;
    dV0 = 0.0;
    dV1 = 0.0;
    bool repeat = true;
    while (repeat) {
      {  // terminal cost-to-go
        const double* C = D.cost(N);
        const double* zc = zptr(zsel, N);
        if (tid < n) S.x[tid] = zc[tid];
        if (tid < m) S.u[tid] = zc[n + tid];
        __syncthreads();
        for (int e = tid; e < n * n; e += kLargeThreads) S.P[e] = C[e];
        if (tid < n) {  // lx = Q x + q + H u
          const double *Q = C, *H = C + n * n + m * m, *q = H + n * m;
          double a = Q[tid] * S.x[0];
          for (int j = 1; j < n; ++j) a += Q[tid + j * n] * S.x[j];
          double h2 = H[tid] * S.u[0];
          for (int j = 1; j < m; ++j) h2 += H[tid + j * n] * S.u[j];
          S.p[tid] = a + q[tid] + h2;
        }
        __syncthreads();
      }
      double gs = 0.0;
      bool failed = false;
      for (int k = N - 1; k >= 0 && !failed; --k) {
        const double* C = D.cost(k);
        const double *Q = C, *R = C + n * n, *H = R + m * m, *q = H + n * m, *r = q + n;
        const double* zc = zptr(zsel, k);
        if (tid < n) S.x[tid] = zc[tid];
        if (tid < m) S.u[tid] = zc[n + tid];
        __syncthreads();
        if (tid < n) {  // cost gradient (quadratic_cost.cpp:13-18)
          double a = Q[tid] * S.x[0];
          for (int j = 1; j < n; ++j) a += Q[tid + j * n] * S.x[j];
          double h2 = H[tid] * S.u[0];
          for (int j = 1; j < m; ++j) h2 += H[tid + j * n] * S.u[j];
          S.lx[tid] = a + q[tid] + h2;
        } else if (tid >= 64 && tid < 64 + m) {
          const int i = tid - 64;
          double a = R[i] * S.u[0];
          for (int j = 1; j < m; ++j) a += R[i + j * m] * S.u[j];
          double h2 = H[i * n] * S.x[0];
          for (int j = 1; j < n; ++j) h2 += H[j + i * n] * S.x[j];
          S.lu[i] = a + r[i] + h2;
        }
        // CalcActionValueExpansion (knot_point_function_type.hpp:149-164)
        sm_matmul<n, n, n, true, LA>(S.A, S.P, S.T);    // T = A'P
        sm_matmul<m, n, n, true, LA>(S.B, S.P, S.BtP);  // B'P
        __syncthreads();
        for (int e = tid; e < n * n; e += kLargeThreads) {
          const int i = e % n, j = e / n;
          double acc = S.T[i] * S.A[j * LA];
          for (int l = 1; l < n; ++l) acc += S.T[i + l * n] * S.A[l + j * LA];
          S.Qxx[e] = Q[e] + acc;
        }
        for (int e = tid; e < n * m; e += kLargeThreads) {
          const int i = e % n, j = e / n;
          double acc = S.T[i] * S.B[j * LA];
          for (int l = 1; l < n; ++l) acc += S.T[i + l * n] * S.B[l + j * LA];
          S.Qxu[e] = H[e] + acc;
        }
        for (int e = tid; e < m * m; e += kLargeThreads) {
          const int i = e % m, j = e / m;
          double acc = S.BtP[i] * S.B[j * LA];
          for (int l = 1; l < n; ++l) acc += S.BtP[i + l * m] * S.B[l + j * LA];
          S.Quu[e] = R[e] + acc;
        }
        if (tid < n) {
          double acc = S.A[tid * LA] * S.p[0];
          for (int l = 1; l < n; ++l) acc += S.A[l + tid * LA] * S.p[l];
          S.Qx[tid] = S.lx[tid] + acc;
        } else if (tid >= 64 && tid < 64 + m) {
          const int i = tid - 64;
          double acc = S.B[i * LA] * S.p[0];
          for (int l = 1; l < n; ++l) acc += S.B[l + i * LA] * S.p[l];
          S.Qu[i] = S.lu[i] + acc;
        }
        __syncthreads();
        // RegularizeActionValue + LLT (lower, unblocked; pivot <= 0 fails, Q19) by one thread
        if (tid == 0) {
          int okc = 1;
          for (int j = 0; j < m; ++j)
            for (int i = 0; i < m; ++i) S.L[i + j * m] = S.Quu[i + j * m] + (i == j ? 1.0 : 0.0) * reg;
          for (int kk = 0; kk < m && okc; ++kk) {
            double xk = S.L[kk + kk * m];
            if (kk > 0) {
              double sq = 0.0;
              for (int j = 0; j < kk; ++j) sq += S.L[kk + j * m] * S.L[kk + j * m];
              xk -= sq;
            }
            if (xk <= 0.0) {
              okc = 0;
              break;
            }
            xk = sqrt(xk);
            S.L[kk + kk * m] = xk;
            for (int i = kk + 1; i < m; ++i) {
              double a = S.L[i + kk * m];
              if (kk > 0) {
                double dot = 0.0;
                for (int j = 0; j < kk; ++j) dot += S.L[i + j * m] * S.L[kk + j * m];
                a -= dot;
              }
              S.L[i + kk * m] = a / xk;
            }
          }
          S.iscal[0] = okc;
        }
        __syncthreads();
        if (!S.iscal[0]) {  // ilqr.hpp:409-427
          increase_reg(o, reg, dreg);
          if (reg >= o.bp_reg_max) max_reg_count++;
          if (max_reg_count >= o.bp_reg_fail_threshold) {
            status = kBackwardPassRegularizationFailed;
            repeat = false;
          }
          failed = true;
          break;
        }
        // CalcGains: column c < n of K from Qxu^T, column n is d from Qu
        if (tid <= n) {
          double bv[m];
          for (int i = 0; i < m; ++i) bv[i] = (tid < n) ? S.Qxu[tid + i * n] : S.Qu[i];
          for (int i = 0; i < m; ++i) {
            double s = bv[i];
            for (int j = 0; j < i; ++j) s -= S.L[i + j * m] * bv[j];
            bv[i] = s / S.L[i + i * m];
          }
          for (int i = m - 1; i >= 0; --i) {
            double s = bv[i];
            for (int j = i + 1; j < m; ++j) s -= S.L[j + i * m] * bv[j];
            bv[i] = s / S.L[i + i * m];
          }
          for (int i = 0; i < m; ++i) {
            if (tid < n) S.K[i + tid * LK] = bv[i] * -1;
            else S.d[i] = bv[i] * -1;
          }
        }
        __syncthreads();
        // CalcCostToGo (unregularised Q, Q3)
        sm_matmul<n, m, m, true, LK>(S.K, S.Quu, S.KtQuu);  // K'Quu (n x m)
        __syncthreads();
        if (tid < n) {
          double a1 = S.KtQuu[tid] * S.d[0], a2 = S.K[tid * LK] * S.Qu[0], a3 = S.Qxu[tid] * S.d[0];
          for (int l = 1; l < m; ++l) {
            a1 += S.KtQuu[tid + l * n] * S.d[l];
            a2 += S.K[l + tid * LK] * S.Qu[l];
            a3 += S.Qxu[tid + l * n] * S.d[l];
          }
          S.tmp[tid] = S.Qx[tid] + a1 + a2 + a3;
        }
        for (int e = tid; e < n * n; e += kLargeThreads) {
          const int i = e % n, j = e / n;
          double t1 = S.KtQuu[i] * S.K[j * LK], t2 = S.K[i * LK] * S.Qxu[j], t3 = S.Qxu[i] * S.K[j * LK];
          for (int l = 1; l < m; ++l) {
            t1 += S.KtQuu[i + l * n] * S.K[l + j * LK];
            t2 += S.K[l + i * LK] * S.Qxu[j + l * n];
            t3 += S.Qxu[i + l * n] * S.K[l + j * LK];
          }
          S.T[e] = S.Qxx[e] + t1 + t2 + t3;  // new P (into T, swapped below)
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += kLargeThreads) S.P[e] = S.T[e];
        if (tid < n) S.p[tid] = S.tmp[tid];
        {
          double a = S.d[0] * S.Qu[0];
          for (int i = 1; i < m; ++i) a += S.d[i] * S.Qu[i];
          double bq = 0.0;
          for (int i = 0; i < m; ++i) {
            double qd = S.Quu[i] * S.d[0];
            for (int l = 1; l < m; ++l) qd += S.Quu[i + l * m] * S.d[l];
            bq = (i == 0) ? S.d[0] * qd : bq + S.d[i] * qd;
          }
          dV0 += a;
          dV1 += 0.5 * bq;
          double g = fabs(S.d[0]) / (fabs(S.u[0]) + 1);
          for (int i = 1; i < m; ++i) g = fmax(g, fabs(S.d[i]) / (fabs(S.u[i]) + 1));
          gs += g;
        }
        double* pk = kdptr(k);
        for (int e = tid; e < m * n; e += kLargeThreads) pk[e] = S.K[e % m + (e / m) * LK];
        if (tid < m) pk[m * n + tid] = S.d[tid];
        __syncthreads();
        if (k == 0) repeat = false;
      }
      gsum = gs;
    }
    decrease_reg(o, reg, dreg);
  };

  // ---- the solve (same control flow as k_solve, one instance, all threads in lock-step) ----
  double penalty = P.sc[static_cast<size_t>(S_PENALTY) * P.Bp + b];
  double reg = 0.0, dreg = 0.0, dV0 = 0.0, dV1 = 0.0, cost_cur = 0.0, cost_prev = 0.0, initial_cost = 0.0;
  double dJ = 0.0, grad = 0.0, alpha_stat = 0.0, z_stat = 0.0, J0 = 0.0, viol = 0.0;
  int zsel = P.is[static_cast<size_t>(I_ZSEL) * P.Bp + b];
  int it_inner = 0, it_outer = P.is[static_cast<size_t>(I_ITERS_OUTER) * P.Bp + b];
  int it_total = P.is[static_cast<size_t>(I_ITERS_TOTAL) * P.Bp + b];
  int st = kUnsolved, st_al = kUnsolved;
  cost_cur = P.sc[static_cast<size_t>(S_COST_CUR) * P.Bp + b];
  cost_prev = P.sc[static_cast<size_t>(S_COST_PREV) * P.Bp + b];
  if (mode == 1) {  // Init()
    if (o.initial_penalty > 0) penalty = o.initial_penalty;
    it_outer = 0;
    it_total = 0;
    cost_cur = 0.0;
    cost_prev = 0.0;
  }
  for (int outer = 0;; ++outer) {
    it_inner = 0;
    st = kUnsolved;
    reg = o.bp_reg_initial;
    dreg = 0.0;
    double gtmp;
    forward(false, zsel, zsel, 0.0, &J0, &gtmp, &st);
    initial_cost = J0;
    bool run = o.max_iterations_inner > 0;
    while (run) {
      double gs_bwd = 0.0;
      backward(zsel, reg, dreg, dV0, dV1, st, gs_bwd);
      double alpha = 1.0, z = -1.0, gs_acc = 0.0;
      bool success = false;
      for (int t = 0; t < o.line_search_max_iterations; ++t) {  // ForwardPass, ilqr.hpp:512-558
        double J, gs;
        if (forward(true, zsel, zsel ^ 1, alpha, &J, &gs, &st)) {
          const double expected = -alpha * (dV0 + alpha * dV1);
          z = (expected > 0.0) ? (J0 - J) / expected : -1.0;
          if (o.line_search_lower_bound <= z && z <= o.line_search_upper_bound && J < J0) {
            success = true;
            cost_cur = J;
            alpha_stat = alpha;
            z_stat = z;
            J0 = J;
            gs_acc = gs;
            break;
          }
        }
        alpha /= o.line_search_decrease_factor;
      }
      if (success) {
        zsel ^= 1;
        grad = gs_acc / static_cast<double>(N);
      } else {
        increase_reg(o, reg, dreg);
        grad = gs_bwd / static_cast<double>(N);
      }
      dJ = (it_inner == 0) ? (initial_cost - cost_cur) : (cost_prev - cost_cur);
      it_inner++;
      it_total++;
      cost_prev = cost_cur;
      if (dJ < o.cost_tolerance && grad < o.gradient_tolerance) {
        st = kSolved;
        run = false;
      } else if (it_inner >= o.max_iterations_inner) {
        st = kMaxInnerIterations;
        run = false;
      } else if (it_total >= o.max_iterations_total) {
        st = kMaxIterations;
        run = false;
      } else if (st != kUnsolved) {
        run = false;
      }
    }
    if (mode == 0) break;
    it_outer++;  // no constraints: UpdateDuals is a no-op, violation and max penalty are 0
    viol = 0.0;
    if (st != kSolved) {
      st_al = st;
    } else {
      st_al = kSolved;  // viol (0) < constraint_tolerance
    }
    break;
  }
  {  // final Cost()
    double Jf = 0.0;
    for (int k = 0; k <= N; ++k) {
      const double* zc = zptr(zsel, k);
      if (tid < n) S.x[tid] = zc[tid];
      if (tid < m) S.u[tid] = zc[n + tid];
      __syncthreads();
      knot_cost_sm(k);
      Jf += S.scal[15];
    }
    if (tid == 0) rsc_cost = Jf;
  }
  if (tid == 0) {
    auto SC = [&](int f) -> double& { return P.sc[static_cast<size_t>(f) * P.Bp + b]; };
    auto IS = [&](int f) -> int& { return P.is[static_cast<size_t>(f) * P.Bp + b]; };
    SC(S_REG) = reg; SC(S_DREG) = dreg; SC(S_DV0) = dV0; SC(S_DV1) = dV1; SC(S_PENALTY) = penalty;
    SC(S_VIOL) = viol; SC(S_INITIAL_COST) = initial_cost; SC(S_COST_CUR) = cost_cur;
    SC(S_COST_PREV) = cost_prev; SC(S_DJ) = dJ; SC(S_GRAD) = grad; SC(S_ALPHA) = alpha_stat;
    SC(S_ZRATIO) = z_stat; SC(S_CSRC_ALPHA) = -1.0; SC(S_J0) = J0;
    IS(I_STATUS) = st; IS(I_STATUS_AL) = st_al; IS(I_ITERS_INNER) = it_inner;
    IS(I_ITERS_OUTER) = it_outer; IS(I_ITERS_TOTAL) = it_total; IS(I_ZSEL) = zsel;
    IS(I_PHASE) = kPhReported;
  }
}

}  // namespace altro_b200
