// kernels.cuh — sweeps over the horizon (one problem instance per lane) and the kernels
// built from them.
//
//   sweep_forward   iLQR::Rollout / RolloutClosedLoop fused with Cost()      (ilqr.hpp:453-499, 326-334)
//   sweep_backward  UpdateExpansions fused with BackwardPass: the knot's dynamics Jacobian and
//                   cost/constraint expansion are produced in registers right before the Riccati
//                   step consumes them — nothing is materialised            (ilqr.hpp:350-445)
//   sweep_dual      ALCost::UpdateDuals + MaxViolation                       (al_cost.hpp:314-352)
//   k_solve         the whole AugmentedLagrangianiLQR::Solve / iLQR::Solve, persistent, no host
//                   round trip                                               (al_solver.hpp:304-334)
//   k_update_expansions / k_backward_mat   the materialised-expansion data flow of the reference
//                   (SURVEY.md 8d contract layout); k_backward_mat streams the per-knot records
//                   through shared memory with TMA bulk copies
//   + step-wise kernels behind the public methods of iLQR<n,m>.
#pragma once

#include "device.cuh"

namespace altro_b200 {

// ------------------------------------------------------------------------------------------
// Lane view of the batch arrays
// ------------------------------------------------------------------------------------------
template <class M>
struct Lane {
  static constexpr int n = M::n, m = M::m, nz = M::n + M::m, nkd = M::m * M::n + M::m;
  static constexpr int nexp = M::n * (M::n + M::m) + M::n * M::n + M::n * M::m + M::m * M::m + M::n + M::m;
  static constexpr int nctg = M::n * M::n + M::n;
  const SolverParams& P;
  Desc D;
  int tile, lane, b, Bp;
  bool valid;
  __device__ __forceinline__ Lane(const SolverParams& P_, const char* blob, int tile_, int lane_)
      : P(P_), D(blob), tile(tile_), lane(lane_), b(tile_ * kTile + lane_), Bp(P_.T * kTile),
        valid(tile_ * kTile + lane_ < P_.B) {}
  __device__ __forceinline__ double* z(int sel, int k) const {
    return P.Z[sel] + (static_cast<size_t>(tile) * (P.N + 1) + k) * nz * kTile + lane;
  }
  __device__ __forceinline__ double* kd(int k) const {
    return P.KD + (static_cast<size_t>(tile) * P.N + k) * nkd * kTile + lane;
  }
  __device__ __forceinline__ double* lam(int k) const {
    return P.LAM + (static_cast<size_t>(tile) * (P.N + 1) + k) * P.pmax * kTile + lane;
  }
  __device__ __forceinline__ double* x0() const {
    return P.X0 + static_cast<size_t>(tile) * n * kTile + lane;
  }
  __device__ __forceinline__ double* exp(int k) const {
    return P.EXP + (static_cast<size_t>(tile) * (P.N + 1) + k) * nexp * kTile + lane;
  }
  __device__ __forceinline__ double* ctg(int k) const {
    return P.CTG + (static_cast<size_t>(tile) * (P.N + 1) + k) * nctg * kTile + lane;
  }
  __device__ __forceinline__ double* costs(int k) const {
    return P.COSTS + (static_cast<size_t>(tile) * (P.N + 1) + k) * kTile + lane;
  }
  __device__ __forceinline__ double& sc(int f) const { return P.sc[static_cast<size_t>(f) * Bp + b]; }
  __device__ __forceinline__ int& is(int f) const { return P.is[static_cast<size_t>(f) * Bp + b]; }
};

__device__ __forceinline__ void prefetch_rows(const double* p, int rows) {
  for (int r = 0; r < rows; ++r) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + r * kTile));
}

__device__ __forceinline__ void copy_blob(const char* __restrict__ g, char* s, int bytes) {
  const int4* src = reinterpret_cast<const int4*>(g);
  int4* dst = reinterpret_cast<int4*>(s);
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
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

// ------------------------------------------------------------------------------------------
// Forward sweep.  kClosed = false: Rollout() in place on Z (zsel) + Cost().
//                 kClosed = true : RolloutClosedLoop(alpha) from Z (zsel) into Zbar (zsel^1) +
//                                  Cost(Zbar) + the normalised feed-forward gain of the candidate.
// Returns false when the state/control bound check trips (status is set like the reference).
// ------------------------------------------------------------------------------------------
template <class M, bool kClosed>
__device__ __forceinline__ bool sweep_forward(const Lane<M>& L, int zsel, double alpha,
                                              double penalty, double& J, double& gsum,
                                              double* viol_out, int& status) {
  constexpr int n = M::n, m = M::m, nz = n + m;
  const int N = L.P.N;
  const DevOptions& o = L.P.opt;
  const Desc& D = L.D;
  const double* mp = D.params();
  const bool has_lam = L.P.pmax > 0;
  double x[n], u[m];
  {
    const double* px0 = L.x0();
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) x[i] = px0[i * kTile];
  }
  double Jsum = 0.0, g = 0.0, vmax = 0.0;
  bool ok = true;
  for (int k = 0; k < N; ++k) {
    const double* zc = L.z(zsel, k);
    double* zn = kClosed ? L.z(zsel ^ 1, k) : L.z(zsel, k);
    if (k + 1 <= N) {  // pull the next knot's rows towards L1 while this knot computes
      prefetch_rows(L.z(zsel, k + 1), nz);
      if (kClosed && k + 1 < N) prefetch_rows(L.kd(k + 1), Lane<M>::nkd);
      if (has_lam) prefetch_rows(L.lam(k + 1), L.P.pmax);
    }
    if (kClosed) {
      const double* pk = L.kd(k);
      double dx[n];
      ALTRO_UNROLL
      for (int i = 0; i < n; ++i) dx[i] = x[i] - zc[i * kTile];
      ALTRO_UNROLL
      for (int i = 0; i < m; ++i) {
        double acc = pk[i * kTile] * dx[0];
        ALTRO_UNROLL
        for (int j = 1; j < n; ++j) acc += pk[(i + j * m) * kTile] * dx[j];
        const double di = pk[(m * n + i) * kTile];
        u[i] = zc[(n + i) * kTile] + acc + di * alpha;  // ilqr.hpp:478
        g = (i == 0) ? fabs(di) / (fabs(u[i]) + 1) : fmax(g, fabs(di) / (fabs(u[i]) + 1));
      }
    } else {
      ALTRO_UNROLL
      for (int i = 0; i < m; ++i) u[i] = zc[(n + i) * kTile];
    }
    if (ok) {
      ALTRO_UNROLL
      for (int i = 0; i < n; ++i) zn[i * kTile] = x[i];
      if (kClosed) {
        ALTRO_UNROLL
        for (int i = 0; i < m; ++i) zn[(n + i) * kTile] = u[i];
      }
      double v;
      Jsum += knot_cost<n, m>(D, k, x, u, has_lam ? L.lam(k) : nullptr, penalty, &v);
      vmax = fmax(vmax, v);
      if (kClosed) gsum += g;
      double xn[n];
      rk4_step<M>(mp, x, u, D.h(k), xn);
      ALTRO_UNROLL
      for (int i = 0; i < n; ++i) x[i] = xn[i];
      if (kClosed && o.check_forwardpass_bounds) {  // ilqr.hpp:484-495
        double sx = 0.0, su = 0.0;
        ALTRO_UNROLL
        for (int i = 0; i < n; ++i) sx += x[i] * x[i];
        ALTRO_UNROLL
        for (int i = 0; i < m; ++i) su += u[i] * u[i];
        if (sqrt(sx) > o.state_max) {
          status = kStateLimit;
          ok = false;
        } else if (sqrt(su) > o.control_max) {
          status = kControlLimit;
          ok = false;
        }
      }
    }
  }
  if (ok) {
    double* zn = kClosed ? L.z(zsel ^ 1, N) : L.z(zsel, N);
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) zn[i * kTile] = x[i];
    double uN[m];
    ALTRO_UNROLL
    for (int i = 0; i < m; ++i) uN[i] = kClosed ? 0.0 : zn[(n + i) * kTile];  // Q14
    if (kClosed) {
      ALTRO_UNROLL
      for (int i = 0; i < m; ++i) zn[(n + i) * kTile] = 0.0;
    }
    double v;
    Jsum += knot_cost<n, m>(D, N, x, uN, has_lam ? L.lam(N) : nullptr, penalty, &v);
    vmax = fmax(vmax, v);
    if (kClosed) status = kUnsolved;  // ilqr.hpp:497
  }
  J = Jsum;
  if (viol_out) *viol_out = vmax;
  return ok;
}

// Cost() of a stored trajectory (no dynamics): ilqr.hpp:326-334, 758-763.
template <class M, bool kStoreCosts>
__device__ __forceinline__ double sweep_cost(const Lane<M>& L, int sel, double penalty,
                                             double* viol_out) {
  constexpr int n = M::n, m = M::m, nz = n + m;
  const int N = L.P.N;
  const bool has_lam = L.P.pmax > 0;
  double J = 0.0, vmax = 0.0;
  for (int k = 0; k <= N; ++k) {
    const double* zc = L.z(sel, k);
    if (k < N) {
      prefetch_rows(L.z(sel, k + 1), nz);
      if (has_lam) prefetch_rows(L.lam(k + 1), L.P.pmax);
    }
    double x[n], u[m];
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) x[i] = zc[i * kTile];
    ALTRO_UNROLL
    for (int i = 0; i < m; ++i) u[i] = zc[(n + i) * kTile];
    double v;
    const double c = knot_cost<n, m>(L.D, k, x, u, has_lam ? L.lam(k) : nullptr, penalty, &v);
    if (kStoreCosts) *L.costs(k) = c;
    J += c;
    vmax = fmax(vmax, v);
  }
  if (viol_out) *viol_out = vmax;
  return J;
}

// Cost + dynamics expansion of knot k in registers (UpdateExpansionsBlock body, ilqr.hpp:670-677).
template <class M>
__device__ __forceinline__ void knot_expansion(const Lane<M>& L, int k, const double* x,
                                               const double* u, double penalty, double* A,
                                               double* B, double* lxx, double* lxu, double* luu,
                                               double* lx, double* lu) {
  constexpr int n = M::n, m = M::m;
  const double* C = L.D.cost(k);
  ALTRO_UNROLL
  for (int i = 0; i < n * n; ++i) lxx[i] = C[i];                  // quadratic_cost.cpp:20-28
  ALTRO_UNROLL
  for (int i = 0; i < m * m; ++i) luu[i] = C[n * n + i];
  ALTRO_UNROLL
  for (int i = 0; i < n * m; ++i) lxu[i] = C[n * n + m * m + i];
  quad_gradient<n, m>(C, x, u, lx, lu);
  const ConSet& cs = L.D.conset(k);
  if (cs.nblocks > 0) al_expansion<n, m>(cs, x, u, L.lam(k), penalty, lxx, lxu, luu, lx, lu);
  if (k < L.P.N) rk4_jacobian<M>(L.D.params(), x, u, L.D.h(k), A, B);
}

// ------------------------------------------------------------------------------------------
// Backward sweep: expansions + Riccati recursion, with the regularisation restart loop of
// ilqr.hpp:396-445 (Q4, Q5).  gsum returns sum_k max_i |d_i|/(|u_i|+1) for the CURRENT controls
// (used only when the line search fails and Z_ stays as it is).
// ------------------------------------------------------------------------------------------
template <class M, bool kStoreCtg>
__device__ __forceinline__ void sweep_backward(const Lane<M>& L, int zsel, double penalty,
                                               double& reg, double& dreg, double& dV0,
                                               double& dV1, int& status, double& gsum) {
  constexpr int n = M::n, m = M::m, nz = n + m;
  const int N = L.P.N;
  const DevOptions& o = L.P.opt;
  const bool has_lam = L.P.pmax > 0;
  int max_reg_count = 0;
  dV0 = 0.0;
  dV1 = 0.0;
  bool repeat = true;
  while (repeat) {
    double P[n * n], p[n];
    {
      const double* zc = L.z(zsel, N);
      double x[n], u[m], A[1], B[1], lxu[n * m], luu[m * m], lu[m];
      ALTRO_UNROLL
      for (int i = 0; i < n; ++i) x[i] = zc[i * kTile];
      ALTRO_UNROLL
      for (int i = 0; i < m; ++i) u[i] = zc[(n + i) * kTile];
      knot_expansion<M>(L, N, x, u, penalty, A, B, P, lxu, luu, p, lu);  // CalcTerminalCostToGo
      if (kStoreCtg) {
        double* c = L.ctg(N);
        ALTRO_UNROLL
        for (int i = 0; i < n * n; ++i) c[i * kTile] = P[i];
        ALTRO_UNROLL
        for (int i = 0; i < n; ++i) c[(n * n + i) * kTile] = p[i];
      }
    }
    gsum = 0.0;
    for (int k = N - 1; k >= 0; --k) {
      const double* zc = L.z(zsel, k);
      if (k > 0) {
        prefetch_rows(L.z(zsel, k - 1), nz);
        if (has_lam) prefetch_rows(L.lam(k - 1), L.P.pmax);
      }
      double x[n], u[m];
      ALTRO_UNROLL
      for (int i = 0; i < n; ++i) x[i] = zc[i * kTile];
      ALTRO_UNROLL
      for (int i = 0; i < m; ++i) u[i] = zc[(n + i) * kTile];
      double A[n * n], B[n * m], lxx[n * n], lxu[n * m], luu[m * m], lx[n], lu[m];
      knot_expansion<M>(L, k, x, u, penalty, A, B, lxx, lxu, luu, lx, lu);
      double K[m * n], d[m];
      const bool ok = riccati_step<n, m>(A, B, lxx, lxu, luu, lx, lu, P, p, reg, K, d, &dV0, &dV1);
      if (!ok) {  // ilqr.hpp:409-427
        increase_reg(o, reg, dreg);
        if (reg >= o.bp_reg_max) max_reg_count++;
        if (max_reg_count >= o.bp_reg_fail_threshold) {
          status = kBackwardPassRegularizationFailed;
          repeat = false;
        }
        break;
      }
      double* pk = L.kd(k);
      ALTRO_UNROLL
      for (int i = 0; i < m * n; ++i) pk[i * kTile] = K[i];
      ALTRO_UNROLL
      for (int i = 0; i < m; ++i) pk[(m * n + i) * kTile] = d[i];
      if (kStoreCtg) {
        double* c = L.ctg(k);
        ALTRO_UNROLL
        for (int i = 0; i < n * n; ++i) c[i * kTile] = P[i];
        ALTRO_UNROLL
        for (int i = 0; i < n; ++i) c[(n * n + i) * kTile] = p[i];
      }
      double g = fabs(d[0]) / (fabs(u[0]) + 1);
      ALTRO_UNROLL
      for (int i = 1; i < m; ++i) g = fmax(g, fabs(d[i]) / (fabs(u[i]) + 1));
      gsum += g;
      if (k == 0) repeat = false;
    }
  }
  decrease_reg(o, reg, dreg);  // ilqr.hpp:443-444
}

// ------------------------------------------------------------------------------------------
// Dual update on the constraint values of trajectory buffer `sel` (Q8 decides which), returns
// the max violation of those values.  constraint_values.hpp:192-194, 216-221.
// ------------------------------------------------------------------------------------------
template <class M>
__device__ __forceinline__ double sweep_dual(const Lane<M>& L, int sel, double penalty,
                                             bool update) {
  constexpr int n = M::n, m = M::m;
  const int N = L.P.N;
  double vmax = 0.0;
  if (L.P.pmax == 0) return vmax;
  for (int k = 0; k <= N; ++k) {
    const ConSet& cs = L.D.conset(k);
    if (cs.nblocks == 0) continue;
    const double* zc = L.z(sel, k);
    double* lam = L.lam(k);
    double x[n], u[m];
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) x[i] = zc[i * kTile];
    ALTRO_UNROLL
    for (int i = 0; i < m; ++i) u[i] = zc[(n + i) * kTile];
    for (int bi = 0; bi < cs.nblocks; ++bi) {
      const ConBlock& b = cs.blk[bi];
      for (int i = 0; i < b.p; ++i) {
        const double c = con_row_fast<n, m>(b, i, x, u);
        if (update) {
          const double arg = lam[(b.row0 + i) * kTile] - penalty * c;
          lam[(b.row0 + i) * kTile] = b.equality ? arg : fmin(0.0, arg);
        }
        vmax = fmax(vmax, b.equality ? fabs(c) : fabs(c - fmin(0.0, c)));
      }
    }
  }
  return vmax;
}

// ------------------------------------------------------------------------------------------
// One inner iteration's line search + bookkeeping, shared by k_solve and k_forward_pass.
// ilqr.hpp:512-558.  All lanes of the warp must call it (warp-synchronous loop); `run` masks
// the lanes that take part.
// ------------------------------------------------------------------------------------------
struct LineSearchResult {
  bool success;
  double J, alpha, z, gsum;
};

template <class M>
__device__ __forceinline__ LineSearchResult line_search(const Lane<M>& L, bool run, int zsel,
                                                        double penalty, double J0, double dV0,
                                                        double dV1, int& status, double& csrc) {
  const DevOptions& o = L.P.opt;
  LineSearchResult r;
  r.success = false;
  r.J = J0;
  r.alpha = 1.0;
  r.z = -1.0;
  r.gsum = 0.0;
  double alpha = 1.0;
  int tries = 0;
  bool ls = run && (o.line_search_max_iterations > 0);
  while (__any_sync(kFull, ls)) {
    if (ls) {
      double J, gs = 0.0;
      const bool ok = sweep_forward<M, true>(L, zsel, alpha, penalty, J, gs, nullptr, status);
      if (ok) {
        csrc = alpha;  // Cost(*Zbar_) refreshed every constraint's stored value (Q8)
        const double expected = -alpha * (dV0 + alpha * dV1);
        double z = -1.0;
        if (expected > 0.0) z = (J0 - J) / expected;
        r.z = z;
        if (o.line_search_lower_bound <= z && z <= o.line_search_upper_bound && J < J0) {
          r.success = true;
          r.J = J;
          r.alpha = alpha;
          r.gsum = gs;
          ls = false;
        }
      }
      if (ls) {
        alpha /= o.line_search_decrease_factor;
        if (++tries >= o.line_search_max_iterations) ls = false;
      }
    }
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// k_solve: whole solve per instance, one warp per tile, persistent until every lane is done.
// mode 0 = iLQR::Solve on the current cost (duals/penalty as they are); mode 1 = AL solve.
// ------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(kTile) k_solve(SolverParams P, int mode) {
  extern __shared__ __align__(16) char s_blob[];
  copy_blob(P.blob, s_blob, P.blob_bytes);
  const Lane<M> L(P, s_blob, blockIdx.x, threadIdx.x);
  const DevOptions& o = P.opt;
  const int N = P.N;
  const bool valid = L.valid;
  const bool has_con = P.pmax > 0;

  double penalty = 1.0, reg = 0.0, dreg = 0.0, dV0 = 0.0, dV1 = 0.0;
  double cost_cur = 0.0, cost_prev = 0.0, initial_cost = 0.0, viol = 0.0;
  double dJ = 0.0, grad = 0.0, alpha_stat = 0.0, z_stat = 0.0, csrc = -1.0, J0 = 0.0;
  int zsel = 0, it_inner = 0, it_outer = 0, it_total = 0, st = kUnsolved, st_al = kUnsolved;
  if (valid) {
    penalty = L.sc(S_PENALTY);
    cost_cur = L.sc(S_COST_CUR);
    cost_prev = L.sc(S_COST_PREV);
    viol = L.sc(S_VIOL);
    alpha_stat = L.sc(S_ALPHA);
    z_stat = L.sc(S_ZRATIO);
    zsel = L.is(I_ZSEL);
    it_outer = L.is(I_ITERS_OUTER);
    it_total = L.is(I_ITERS_TOTAL);
    st_al = L.is(I_STATUS_AL);
  }
  if (mode == 1 && valid) {  // AugmentedLagrangianiLQR::Init, al_solver.hpp:287-302 (Q10)
    if (o.reset_duals && has_con) {
      for (int k = 0; k <= N; ++k) {
        double* lam = L.lam(k);
        for (int r = 0; r < P.pmax; ++r) lam[r * kTile] = 0.0;
      }
    }
    if (o.initial_penalty > 0) penalty = o.initial_penalty;
    it_outer = 0;  // stats.Reset()
    it_total = 0;
    cost_cur = 0.0;
    cost_prev = 0.0;
    st_al = kUnsolved;
  }

  bool al_run = valid;
  for (int outer = 0;; ++outer) {
    // ================================ iLQR::Solve, ilqr.hpp:284-316 =========================
    bool run = al_run;
    it_inner = 0;  // SolveSetup :629-645, ResetInternalVariables :680-690
    st = kUnsolved;
    reg = o.bp_reg_initial;
    dreg = 0.0;
    dV0 = dV1 = 0.0;
    if (run) {
      double gs = 0.0;
      sweep_forward<M, false>(L, zsel, 0.0, penalty, J0, gs, nullptr, st);  // Rollout + Cost
      initial_cost = J0;
      csrc = -1.0;
      if (o.max_iterations_inner <= 0) run = false;
    }
    while (__any_sync(kFull, run)) {
      double gs_bwd = 0.0;
      if (run) {
        csrc = -1.0;  // UpdateExpansions evaluates every constraint at Z_
        sweep_backward<M, false>(L, zsel, penalty, reg, dreg, dV0, dV1, st, gs_bwd);
      }
      const LineSearchResult ls = line_search<M>(L, run, zsel, penalty, J0, dV0, dV1, st, csrc);
      if (run) {
        if (ls.success) {
          zsel ^= 1;  // (*Z_) = (*Zbar_)
          J0 = ls.J;
          cost_cur = ls.J;  // stats.Log("cost"/"alpha"/"z")
          alpha_stat = ls.alpha;
          z_stat = ls.z;
          csrc = -1.0;  // the stored constraint values are those of the new Z_
          grad = ls.gsum / static_cast<double>(N);
        } else {
          increase_reg(o, reg, dreg);
          grad = gs_bwd / static_cast<double>(N);
        }
        // UpdateConvergenceStatistics, ilqr.hpp:568-587 (Q15)
        dJ = (it_inner == 0) ? (initial_cost - cost_cur) : (cost_prev - cost_cur);
        it_inner++;
        it_total++;
        cost_prev = cost_cur;  // NewIteration() carry-forward (Q6)
        // IsDone, ilqr.hpp:597-619
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
    }
    if (mode == 0) break;
    // ========================= AL outer step, al_solver.hpp:313-333 =========================
    if (al_run) {
      int src = zsel;
      if (csrc >= 0.0) {  // Q8: the stored values are those of the last evaluated candidate
        double Jt, gt = 0.0;
        int stt = st;
        sweep_forward<M, true>(L, zsel, csrc, penalty, Jt, gt, nullptr, stt);
        src = zsel ^ 1;
      }
      viol = sweep_dual<M>(L, src, penalty, /*update=*/true);  // UpdateDuals + GetMaxViolation
      it_outer++;
      const double max_penalty = has_con ? penalty : 0.0;
      // IsDone, al_solver.hpp:368-401 (Q16)
      if (st != kSolved) {
        st_al = st;
        al_run = false;
      } else if (viol < o.constraint_tolerance) {
        st_al = kSolved;
        al_run = false;
      } else if (max_penalty > o.maximum_penalty) {
        st_al = kMaxPenalty;
        al_run = false;
      } else if (it_outer >= o.max_iterations_outer) {
        st_al = kMaxOuterIterations;
        al_run = false;
      } else if (it_total >= o.max_iterations_total) {
        st_al = kMaxIterations;
        al_run = false;
      } else {
        penalty *= o.penalty_scaling;  // UpdatePenalties
      }
    }
    if (!__any_sync(kFull, al_run) || outer + 1 >= o.max_iterations_outer) break;
  }

  if (valid) {
    // Cost() of the final trajectory under the final duals/penalty, as printed by
    // perf/benchmark_unicycle.cpp:73-74.
    double v;
    const double Jf = sweep_cost<M, false>(L, zsel, penalty, &v);
    if (mode == 0) viol = v;  // == Cost(); GetMaxViolation()
    L.sc(S_REG) = reg;
    L.sc(S_DREG) = dreg;
    L.sc(S_DV0) = dV0;
    L.sc(S_DV1) = dV1;
    L.sc(S_PENALTY) = penalty;
    L.sc(S_VIOL) = viol;
    L.sc(S_COST) = Jf;
    L.sc(S_INITIAL_COST) = initial_cost;
    L.sc(S_COST_CUR) = cost_cur;
    L.sc(S_COST_PREV) = cost_prev;
    L.sc(S_DJ) = dJ;
    L.sc(S_GRAD) = grad;
    L.sc(S_ALPHA) = alpha_stat;
    L.sc(S_ZRATIO) = z_stat;
    L.sc(S_CSRC_ALPHA) = -1.0;
    L.is(I_STATUS) = st;
    L.is(I_STATUS_AL) = st_al;
    L.is(I_ITERS_INNER) = it_inner;
    L.is(I_ITERS_OUTER) = it_outer;
    L.is(I_ITERS_TOTAL) = it_total;
    L.is(I_ZSEL) = zsel;
  }
}

// ------------------------------------------------------------------------------------------
// Step-wise kernels (public methods of iLQR<n,m>), one warp per tile.
// ------------------------------------------------------------------------------------------
enum Phase : int {
  kPhaseRollout = 0,
  kPhaseCost,
  kPhaseBackwardFused,   // expansions in registers + Riccati (what k_solve does)
  kPhaseForward,
  kPhaseStats,
  kPhaseDuals,
  kPhasePenalties,
  kPhaseSolveSetup,
};

template <class M>
__global__ void __launch_bounds__(kTile) k_phase(SolverParams P, int phase) {
  extern __shared__ __align__(16) char s_blob[];
  copy_blob(P.blob, s_blob, P.blob_bytes);
  const Lane<M> L(P, s_blob, blockIdx.x, threadIdx.x);
  const DevOptions& o = P.opt;
  const int N = P.N;
  const bool valid = L.valid;
  int zsel = valid ? L.is(I_ZSEL) : 0;
  double penalty = valid ? L.sc(S_PENALTY) : 1.0;
  int st = valid ? L.is(I_STATUS) : kUnsolved;
  switch (phase) {
    case kPhaseSolveSetup: {  // SolveSetup(), ilqr.hpp:629-645
      if (valid) {
        L.is(I_ITERS_INNER) = 0;
        L.is(I_STATUS) = kUnsolved;
        L.sc(S_REG) = o.bp_reg_initial;
        L.sc(S_DREG) = 0.0;
        L.sc(S_DV0) = 0.0;
        L.sc(S_DV1) = 0.0;
      }
      break;
    }
    case kPhaseRollout: {
      if (valid) {
        double J, gs = 0.0;
        sweep_forward<M, false>(L, zsel, 0.0, penalty, J, gs, nullptr, st);
      }
      break;
    }
    case kPhaseCost: {
      if (valid) {
        double v;
        const double J = sweep_cost<M, true>(L, zsel, penalty, &v);
        L.sc(S_COST) = J;
        L.sc(S_VIOL) = v;
        L.sc(S_CSRC_ALPHA) = -1.0;
      }
      break;
    }
    case kPhaseBackwardFused: {
      if (valid) {
        double reg = L.sc(S_REG), dreg = L.sc(S_DREG), dV0, dV1, gs;
        sweep_backward<M, true>(L, zsel, penalty, reg, dreg, dV0, dV1, st, gs);
        L.sc(S_REG) = reg;
        L.sc(S_DREG) = dreg;
        L.sc(S_DV0) = dV0;
        L.sc(S_DV1) = dV1;
        L.is(I_STATUS) = st;
      }
      break;
    }
    case kPhaseForward: {  // ForwardPass(), ilqr.hpp:512-558
      double J0 = 0.0, dV0 = 0.0, dV1 = 0.0, csrc = -1.0;
      if (valid) {
        for (int k = 0; k <= N; ++k) J0 += *L.costs(k);  // Q7: costs_ from UpdateExpansions
        dV0 = L.sc(S_DV0);
        dV1 = L.sc(S_DV1);
        csrc = L.sc(S_CSRC_ALPHA);
      }
      const LineSearchResult ls = line_search<M>(L, valid, zsel, penalty, J0, dV0, dV1, st, csrc);
      if (valid) {
        if (ls.success) {
          L.is(I_ZSEL) = zsel ^ 1;
          L.sc(S_COST) = ls.J;
          L.sc(S_COST_CUR) = ls.J;
          L.sc(S_ALPHA) = ls.alpha;
          L.sc(S_ZRATIO) = ls.z;
          csrc = -1.0;
        } else {
          double reg = L.sc(S_REG), dreg = L.sc(S_DREG);
          increase_reg(o, reg, dreg);
          L.sc(S_REG) = reg;
          L.sc(S_DREG) = dreg;
        }
        L.sc(S_CSRC_ALPHA) = csrc;
        L.is(I_STATUS) = st;
      }
      break;
    }
    case kPhaseStats: {  // UpdateConvergenceStatistics(), ilqr.hpp:568-587, 662-668
      if (valid) {
        double gsum = 0.0;
        for (int k = 0; k < N; ++k) {
          const double* zc = L.z(zsel, k);
          const double* pk = L.kd(k);
          double g = 0.0;
          for (int i = 0; i < M::m; ++i) {
            const double gi = fabs(pk[(M::m * M::n + i) * kTile]) / (fabs(zc[(M::n + i) * kTile]) + 1);
            g = (i == 0) ? gi : fmax(g, gi);
          }
          gsum += g;
        }
        const double grad = gsum / static_cast<double>(N);
        const int it_inner = L.is(I_ITERS_INNER);
        const double cost_cur = L.sc(S_COST_CUR);
        const double dJ = (it_inner == 0) ? (L.sc(S_INITIAL_COST) - cost_cur)
                                          : (L.sc(S_COST_PREV) - cost_cur);
        L.is(I_ITERS_INNER) = it_inner + 1;
        L.is(I_ITERS_TOTAL) = L.is(I_ITERS_TOTAL) + 1;
        L.sc(S_DJ) = dJ;
        L.sc(S_GRAD) = grad;
        L.sc(S_COST_PREV) = cost_cur;
      }
      break;
    }
    case kPhaseDuals: {  // UpdateDuals(), al_solver.hpp:336-345
      if (valid) {
        int src = zsel;
        const double csrc = L.sc(S_CSRC_ALPHA);
        if (csrc >= 0.0) {
          double Jt, gt = 0.0;
          int stt = st;
          sweep_forward<M, true>(L, zsel, csrc, penalty, Jt, gt, nullptr, stt);
          src = zsel ^ 1;
        }
        L.sc(S_VIOL) = sweep_dual<M>(L, src, penalty, true);
      }
      break;
    }
    case kPhasePenalties: {  // UpdatePenalties(), al_solver.hpp:347-355
      if (valid) L.sc(S_PENALTY) = penalty * o.penalty_scaling;
      break;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Materialised data flow (the reference's UpdateExpansions -> BackwardPass hand-off).
// Record of one (tile, knot): [A | B | lxx | lxu | luu | lx | lu] x 32 lanes, contiguous.
// ------------------------------------------------------------------------------------------
// One warp per (tile, knot): fully parallel over B x (N+1), like the reference's thread-pool
// tasks (ilqr.hpp:354-365) but with the batch as the wide axis.
template <class M>
__global__ void __launch_bounds__(128) k_update_expansions(SolverParams P) {
  extern __shared__ __align__(16) char s_blob[];
  copy_blob(P.blob, s_blob, P.blob_bytes);
  constexpr int n = M::n, m = M::m;
  const int warp = threadIdx.x / kTile, lane = threadIdx.x % kTile;
  const int k = blockIdx.y * (blockDim.x / kTile) + warp;
  if (k > P.N) return;
  const Lane<M> L(P, s_blob, blockIdx.x, lane);
  if (!L.valid) return;
  const int zsel = L.is(I_ZSEL);
  const double penalty = L.sc(S_PENALTY);
  const double* zc = L.z(zsel, k);
  double x[n], u[m];
  ALTRO_UNROLL
  for (int i = 0; i < n; ++i) x[i] = zc[i * kTile];
  ALTRO_UNROLL
  for (int i = 0; i < m; ++i) u[i] = zc[(n + i) * kTile];
  double A[n * n], B[n * m], lxx[n * n], lxu[n * m], luu[m * m], lx[n], lu[m];
  ALTRO_UNROLL
  for (int i = 0; i < n * n; ++i) A[i] = 0.0;
  ALTRO_UNROLL
  for (int i = 0; i < n * m; ++i) B[i] = 0.0;
  if (k == P.N) {  // IdentityDynamics::Jacobian, problem.hpp:40-43: setIdentity on n x (n+m)
    ALTRO_UNROLL
    for (int i = 0; i < n; ++i) A[i + i * n] = 1.0;
  }
  knot_expansion<M>(L, k, x, u, penalty, A, B, lxx, lxu, luu, lx, lu);
  double* e = L.exp(k);
  int f = 0;
  ALTRO_UNROLL
  for (int i = 0; i < n * n; ++i) e[(f++) * kTile] = A[i];
  ALTRO_UNROLL
  for (int i = 0; i < n * m; ++i) e[(f++) * kTile] = B[i];
  ALTRO_UNROLL
  for (int i = 0; i < n * n; ++i) e[(f++) * kTile] = lxx[i];
  ALTRO_UNROLL
  for (int i = 0; i < n * m; ++i) e[(f++) * kTile] = lxu[i];
  ALTRO_UNROLL
  for (int i = 0; i < m * m; ++i) e[(f++) * kTile] = luu[i];
  ALTRO_UNROLL
  for (int i = 0; i < n; ++i) e[(f++) * kTile] = lx[i];
  ALTRO_UNROLL
  for (int i = 0; i < m; ++i) e[(f++) * kTile] = lu[i];
  // costs_(k) = Cost(x,u), ilqr.hpp:675
  *L.costs(k) = knot_cost<n, m>(L.D, k, x, u, P.pmax > 0 ? L.lam(k) : nullptr, penalty, nullptr);
  if (k == 0) L.sc(S_CSRC_ALPHA) = -1.0;
}

// --- TMA 1-D bulk copy + mbarrier helpers (cp.async.bulk -> SASS UBLKCP) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Backward pass over materialised expansions: the kernel BASELINE.json's "backward-pass HBM
// GB/s" is quoted on.  One warp per tile; the (tile, knot) records (nexp*256 B each, contiguous)
// are streamed N-1 .. 0 through a kStages-deep shared-memory ring by TMA bulk copies issued by
// lane 0 and tracked with mbarriers; every lane then reads its own column (conflict-free) and
// runs the Riccati step in registers.  Writes K, d (and P, p when CTG is allocated).
template <class M, int kStages, bool kStoreCtg>
__global__ void __launch_bounds__(kTile) k_backward_mat(SolverParams P) {
  constexpr int n = M::n, m = M::m, nexp = Lane<M>::nexp;
  constexpr uint32_t kRecBytes = nexp * kTile * sizeof(double);
  extern __shared__ __align__(128) char smem[];
  double* ring = reinterpret_cast<double*>(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(kStages) * kRecBytes);
  const int lane = threadIdx.x;
  const Lane<M> L(P, nullptr, blockIdx.x, lane);
  const DevOptions& o = P.opt;
  const int N = P.N;
  const bool valid = L.valid;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const double* rec0 = P.EXP + static_cast<size_t>(blockIdx.x) * (N + 1) * nexp * kTile;

  double reg = 0.0, dreg = 0.0, dV0 = 0.0, dV1 = 0.0;
  int st = kUnsolved;
  if (valid) {
    reg = L.sc(S_REG);
    dreg = L.sc(S_DREG);
    st = L.is(I_STATUS);
  }
  int max_reg_count = 0;
  bool repeat = valid;
  uint32_t issued = 0, consumed = 0;  // monotonically increasing record counters (warp-uniform)
  while (__any_sync(kFull, repeat)) {
    // terminal cost-to-go: lxx, lx of knot N (plain coalesced loads, once per pass)
    double Pm[n * n], p[n];
    {
      const double* e = L.exp(N);
      constexpr int off_lxx = n * (n + m), off_lx = off_lxx + n * n + n * m + m * m;
      ALTRO_UNROLL
      for (int i = 0; i < n * n; ++i) Pm[i] = valid ? e[(off_lxx + i) * kTile] : 0.0;
      ALTRO_UNROLL
      for (int i = 0; i < n; ++i) p[i] = valid ? e[(off_lx + i) * kTile] : 0.0;
      if (kStoreCtg && repeat) {
        double* c = L.ctg(N);
        ALTRO_UNROLL
        for (int i = 0; i < n * n; ++i) c[i * kTile] = Pm[i];
        ALTRO_UNROLL
        for (int i = 0; i < n; ++i) c[(n * n + i) * kTile] = p[i];
      }
    }
    // prologue: fill the ring
    int next_k = N - 1;
    if (lane == 0) {
      for (int s = 0; s < kStages && next_k >= 0; ++s, --next_k, ++issued) {
        const int slot = issued % kStages;
        mbar_expect_tx(&bars[slot], kRecBytes);
        tma_load_1d(ring + static_cast<size_t>(slot) * nexp * kTile,
                    rec0 + static_cast<size_t>(next_k) * nexp * kTile, kRecBytes, &bars[slot]);
      }
    }
    bool live = repeat;  // lanes still descending in this pass
    for (int k = N - 1; k >= 0; --k, ++consumed) {
      const int slot = consumed % kStages;
      mbar_wait(&bars[slot], (consumed / kStages) & 1);
      const double* e = ring + static_cast<size_t>(slot) * nexp * kTile + lane;
      double A[n * n], B[n * m], lxx[n * n], lxu[n * m], luu[m * m], lx[n], lu[m];
      int f = 0;
      ALTRO_UNROLL
      for (int i = 0; i < n * n; ++i) A[i] = e[(f++) * kTile];
      ALTRO_UNROLL
      for (int i = 0; i < n * m; ++i) B[i] = e[(f++) * kTile];
      ALTRO_UNROLL
      for (int i = 0; i < n * n; ++i) lxx[i] = e[(f++) * kTile];
      ALTRO_UNROLL
      for (int i = 0; i < n * m; ++i) lxu[i] = e[(f++) * kTile];
      ALTRO_UNROLL
      for (int i = 0; i < m * m; ++i) luu[i] = e[(f++) * kTile];
      ALTRO_UNROLL
      for (int i = 0; i < n; ++i) lx[i] = e[(f++) * kTile];
      ALTRO_UNROLL
      for (int i = 0; i < m; ++i) lu[i] = e[(f++) * kTile];
      __syncwarp();  // every lane has read slot -> it can be refilled
      if (lane == 0 && next_k >= 0) {
        const int s2 = issued % kStages;
        mbar_expect_tx(&bars[s2], kRecBytes);
        tma_load_1d(ring + static_cast<size_t>(s2) * nexp * kTile,
                    rec0 + static_cast<size_t>(next_k) * nexp * kTile, kRecBytes, &bars[s2]);
        --next_k;
        ++issued;
      }
      if (live) {
        double K[m * n], d[m];
        const bool ok =
            riccati_step<n, m>(A, B, lxx, lxu, luu, lx, lu, Pm, p, reg, K, d, &dV0, &dV1);
        if (!ok) {  // ilqr.hpp:409-427: raise the regularisation, restart from k = N-1
          increase_reg(o, reg, dreg);
          if (reg >= o.bp_reg_max) max_reg_count++;
          if (max_reg_count >= o.bp_reg_fail_threshold) {
            st = kBackwardPassRegularizationFailed;
            repeat = false;
          }
          live = false;
        } else {
          double* pk = L.kd(k);
          ALTRO_UNROLL
          for (int i = 0; i < m * n; ++i) pk[i * kTile] = K[i];
          ALTRO_UNROLL
          for (int i = 0; i < m; ++i) pk[(m * n + i) * kTile] = d[i];
          if (kStoreCtg) {
            double* c = L.ctg(k);
            ALTRO_UNROLL
            for (int i = 0; i < n * n; ++i) c[i * kTile] = Pm[i];
            ALTRO_UNROLL
            for (int i = 0; i < n; ++i) c[(n * n + i) * kTile] = p[i];
          }
          if (k == 0) repeat = false;
        }
      }
    }
  }
  if (valid) {
    decrease_reg(o, reg, dreg);
    L.sc(S_REG) = reg;
    L.sc(S_DREG) = dreg;
    L.sc(S_DV0) = dV0;
    L.sc(S_DV1) = dV1;
    L.is(I_STATUS) = st;
  }
}

// ------------------------------------------------------------------------------------------
// Layout conversion kernels (instance-major <-> tile-major)
// ------------------------------------------------------------------------------------------
// inputs: x0 [B][n], U0 [B][N][m] or nullptr (+ unom[m] by value)
struct Unom { double v[kMaxDim]; };

__global__ void k_pack_inputs(SolverParams P, const double* __restrict__ x0,
                              const double* __restrict__ U0, Unom unom) {
  const int n = P.n, m = P.m, nz = n + m, N = P.N;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (b >= P.T * kTile) return;
  const int tile = b / kTile, lane = b % kTile;
  const int src = b < P.B ? b : P.B - 1;  // padding lanes mirror the last instance
  double* z0 = P.Z[0] + (static_cast<size_t>(tile) * (N + 1) + k) * nz * kTile + lane;
  double* z1 = P.Z[1] + (static_cast<size_t>(tile) * (N + 1) + k) * nz * kTile + lane;
  for (int i = 0; i < n; ++i) {
    z0[i * kTile] = (k == 0) ? x0[static_cast<size_t>(src) * n + i] : 0.0;
    z1[i * kTile] = 0.0;
  }
  for (int i = 0; i < m; ++i) {
    double u = 0.0;
    if (k < N) u = U0 ? U0[(static_cast<size_t>(src) * N + k) * m + i] : unom.v[i];
    z0[(n + i) * kTile] = u;
    z1[(n + i) * kTile] = 0.0;  // Zbar_->SetZero(), ilqr.hpp:233-234
  }
  if (k == 0) {
    double* px0 = P.X0 + static_cast<size_t>(tile) * n * kTile + lane;
    for (int i = 0; i < n; ++i) px0[i * kTile] = x0[static_cast<size_t>(src) * n + i];
    P.is[static_cast<size_t>(I_ZSEL) * P.T * kTile + b] = 0;
    P.sc[static_cast<size_t>(S_CSRC_ALPHA) * P.T * kTile + b] = -1.0;
  }
}

__global__ void k_set_states(SolverParams P, const double* __restrict__ X) {
  const int n = P.n, nz = P.n + P.m, N = P.N;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (b >= P.B) return;
  const int tile = b / kTile, lane = b % kTile;
  const int sel = P.is[static_cast<size_t>(I_ZSEL) * P.T * kTile + b];
  double* z = P.Z[sel] + (static_cast<size_t>(tile) * (N + 1) + k) * nz * kTile + lane;
  for (int i = 0; i < n; ++i) z[i * kTile] = X[(static_cast<size_t>(b) * (N + 1) + k) * n + i];
}

// generic gather: dst[b][f] = src_tile_array[tile][k][f0+f][lane], f < nf.  For the trajectory
// buffers (sel_by_zsel != 0) the source is Z[zsel[b]].
__global__ void k_unpack(SolverParams P, const double* __restrict__ src0,
                         const double* __restrict__ src1, int K, int F, int f0, int nf,
                         int k0, int nk, double* __restrict__ dst, int sel_by_zsel) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int kk = blockIdx.y;
  if (b >= P.B || kk >= nk) return;
  const int tile = b / kTile, lane = b % kTile;
  const double* src = src0;
  if (sel_by_zsel && P.is[static_cast<size_t>(I_ZSEL) * P.T * kTile + b]) src = src1;
  const double* s = src + (static_cast<size_t>(tile) * K + (k0 + kk)) * F * kTile + lane;
  double* d = dst + (static_cast<size_t>(b) * nk + kk) * nf;
  for (int f = 0; f < nf; ++f) d[f] = s[(f0 + f) * kTile];
}

__global__ void k_fill_duals(SolverParams P, int k, const double* __restrict__ lam, int p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.T * kTile) return;
  const int tile = b / kTile, lane = b % kTile;
  double* l = P.LAM + (static_cast<size_t>(tile) * (P.N + 1) + k) * P.pmax * kTile + lane;
  for (int r = 0; r < p; ++r) l[r * kTile] = lam[r];
}

__global__ void k_fill_scalar(double* p, double v, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] = v;
}
__global__ void k_fill_int(int* p, int v, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] = v;
}

}  // namespace altro_b200
