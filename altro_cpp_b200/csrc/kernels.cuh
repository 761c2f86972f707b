// kernels.cuh — sweeps over the horizon and the kernels built from them.
//
//   sweep_forward   iLQR::Rollout / RolloutClosedLoop fused with Cost()      (ilqr.hpp:453-499, 326-334)
//   sweep_backward  UpdateExpansions fused with BackwardPass: the knot's dynamics Jacobian and
//                   cost/constraint expansion are produced in registers right before the Riccati
//                   step consumes them — nothing is materialised            (ilqr.hpp:350-445)
//   sweep_dual      ALCost::UpdateDuals + MaxViolation                       (al_cost.hpp:314-352)
//   line_search     ForwardPass: G = 32/W step lengths of every instance are rolled out at once
//                   on the warp's lane groups; the first acceptable one in the reference's
//                   order wins, so the result equals the sequential search   (ilqr.hpp:512-558)
//   k_solve         the whole AugmentedLagrangianiLQR::Solve / iLQR::Solve, persistent, no host
//                   round trip                                               (al_solver.hpp:304-334)
//   k_update_expansions / k_backward_mat   the materialised-expansion data flow of the reference
//                   (SURVEY.md 8d contract layout); k_backward_mat streams the per-knot records
//                   through shared memory with TMA bulk copies
//   + step-wise kernels behind the public methods of iLQR<n,m>.
//
// All sweeps are warp-uniform: every lane of the warp calls them and walks the same knots;
// `active` masks the lanes that own work.  Per-knot inputs (trajectory, gains, duals) are
// staged one knot ahead into a per-warp shared-memory double buffer with cp.async so that no
// global-memory latency sits on the serial chain.
#pragma once

#include "device.cuh"

#ifndef ALTRO_SOLVE_WARPS
#define ALTRO_SOLVE_WARPS 2
#endif
#ifndef ALTRO_SOLVE_MINB
#define ALTRO_SOLVE_MINB 7
#endif
// register budget of the narrow-tile instantiations (W <= 4), which only run on re-packed
// workspaces with few warps
#ifndef ALTRO_SOLVE_MINB_NARROW
#define ALTRO_SOLVE_MINB_NARROW 4
#endif

namespace altro_b200 {

// ------------------------------------------------------------------------------------------
// Lane view of the batch arrays.  lane = a*W + i : i = instance within the tile, a = group.
// ------------------------------------------------------------------------------------------
template <class M, int W>
struct Lane {
  static constexpr int n = M::n, m = M::m, nz = M::n + M::m, nkd = M::m * M::n + M::m;
  static constexpr int nexp = M::n * (M::n + M::m) + M::n * M::n + M::n * M::m + M::m * M::m + M::n + M::m;
  static constexpr int nctg = M::n * M::n + M::n;
  static constexpr int G = kWarp / W;
  const SolverParams& P;
  Desc D;
  int tile, i, a, b;
  bool valid;
  __device__ __forceinline__ Lane(const SolverParams& P_, const char* blob, int tile_, int lane_)
      : P(P_), D(blob), tile(tile_), i(lane_ % W), a(lane_ / W), b(tile_ * W + lane_ % W),
        valid(tile_ < P_.T && tile_ * W + lane_ % W < P_.B) {}
  __device__ __forceinline__ double* z(int sel, int k) const {
    return P.Z[sel] + (static_cast<size_t>(tile) * (P.N + 1) + k) * nz * W + i;
  }
  __device__ __forceinline__ double* kd(int k) const {
    return P.KD + (static_cast<size_t>(tile) * P.N + k) * nkd * W + i;
  }
  __device__ __forceinline__ double* lam(int k) const {
    return P.LAM + (static_cast<size_t>(tile) * (P.N + 1) + k) * P.pmax * W + i;
  }
  __device__ __forceinline__ double* x0() const {
    return P.X0 + static_cast<size_t>(tile) * n * W + i;
  }
  __device__ __forceinline__ double* exp(int k) const {
    return P.EXP + (static_cast<size_t>(tile) * (P.N + 1) + k) * nexp * W + i;
  }
  __device__ __forceinline__ double* ctg(int k) const {
    return P.CTG + (static_cast<size_t>(tile) * (P.N + 1) + k) * nctg * W + i;
  }
  __device__ __forceinline__ double* costs(int k) const {
    return P.COSTS + (static_cast<size_t>(tile) * (P.N + 1) + k) * W + i;
  }
  __device__ __forceinline__ double& sc(int f) const { return P.sc[static_cast<size_t>(f) * P.Bp + b]; }
  __device__ __forceinline__ int& is(int f) const { return P.is[static_cast<size_t>(f) * P.Bp + b]; }
  // trajectory buffer that holds line-search candidate `slot` while Z_ lives in buffer zsel
  __device__ __forceinline__ static int cand(int zsel, int slot) { return (zsel + 1 + slot) % (G + 1); }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
#ifdef ALTRO_STAGE_CG
  // experiment: stage through L2 only, synchronously
  *smem_dst = __ldcg(gmem_src);
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
               : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void copy_blob(const char* __restrict__ g, char* s, int bytes) {
  const int4* src = reinterpret_cast<const int4*>(g);
  int4* dst = reinterpret_cast<int4*>(s);
  for (int q = threadIdx.x; q < bytes / 16; q += blockDim.x) dst[q] = src[q];
  __syncthreads();
}

// doubles of the per-warp staging double buffer: the forward sweeps stage one knot
// (trajectory, gains, duals) x W instances, the backward sweep G knots (trajectory, duals)
template <class M>
__host__ __device__ constexpr int stage_doubles(int pmax, int W) {
  const int fwd = ((M::n + M::m) + (M::m * M::n + M::m) + pmax) * W;
  const int bwd = ((M::n + M::m) + pmax) * kWarp;
  return 2 * (fwd > bwd ? fwd : bwd);
}

// ------------------------------------------------------------------------------------------
// Forward sweep.  kClosed = false: Rollout() in place on Z (zsel) + Cost().
//                 kClosed = true : RolloutClosedLoop(alpha) from Z (zsel) into buffer zout +
//                                  Cost(Zbar) + the normalised feed-forward gain of the candidate.
// Returns false when the state/control bound check trips (status is set like the reference).
// ------------------------------------------------------------------------------------------
template <class M, int W, bool kClosed>
__device__ __forceinline__ bool sweep_forward(const Lane<M, W>& L, double* stg, bool active,
                                              int zsel, int zout, double alpha, double penalty,
                                              double& J, double& gsum, double* viol_out,
                                              int& status) {
  constexpr int n = M::n, m = M::m, nz = n + m, nkd = Lane<M, W>::nkd;
  const int N = L.P.N, pmax = L.P.pmax;
  const DevOptions& o = L.P.opt;
  const Desc& D = L.D;
  const double* mp = D.params();
  const int off_lam = nz + (kClosed ? nkd : 0);
  const int R = off_lam + pmax;
  const AlPen pen(penalty);
  auto issue = [&](int k) {
    double* s = stg + (k & 1) * R * W + L.i;
    if (L.a == 0) {
      const double* g = L.z(zsel, k);
      ALTRO_UNROLL
      for (int r = 0; r < nz; ++r) cp_async8(s + r * W, g + r * W);
      if (kClosed && k < N) {
        g = L.kd(k);
        ALTRO_UNROLL
        for (int r = 0; r < nkd; ++r) cp_async8(s + (nz + r) * W, g + r * W);
      }
      if (pmax > 0) {
        g = L.lam(k);
        for (int r = 0; r < pmax; ++r) cp_async8(s + (off_lam + r) * W, g + r * W);
      }
    }
    cp_async_commit();
  };
  __syncwarp();
  issue(0);
  double x[n], u[m];
  {
    const double* px0 = L.x0();
    ALTRO_UNROLL
    for (int q = 0; q < n; ++q) x[q] = px0[q * W];
  }
  double Jsum = 0.0, gs = 0.0, vmax = 0.0;
  bool ok = active;
  for (int k = 0; k <= N; ++k) {
    cp_async_wait_all();
    __syncwarp();
    if (k < N) issue(k + 1);
    const double* s = stg + (k & 1) * R * W + L.i;
    if (ok) {
      double* zn = L.z(kClosed ? zout : zsel, k);
      double g = 0.0;
      if (k < N) {
        if (kClosed) {
          double dx[n];
          ALTRO_UNROLL
          for (int q = 0; q < n; ++q) dx[q] = x[q] - s[q * W];
          ALTRO_UNROLL
          for (int q = 0; q < m; ++q) {
            double acc = s[(nz + q) * W] * dx[0];
            ALTRO_UNROLL
            for (int j = 1; j < n; ++j) acc += s[(nz + q + j * m) * W] * dx[j];
            const double dq = s[(nz + m * n + q) * W];
            u[q] = s[(n + q) * W] + acc + dq * alpha;  // ilqr.hpp:478
            const double gq = fabs(dq) / (fabs(u[q]) + 1);
            g = (q == 0) ? gq : fmax(g, gq);
          }
        } else {
          ALTRO_UNROLL
          for (int q = 0; q < m; ++q) u[q] = s[(n + q) * W];
        }
      } else {  // terminal knot: u_N = 0 for Zbar (SetZero), the stored u_N for Z (Q14)
        ALTRO_UNROLL
        for (int q = 0; q < m; ++q) u[q] = kClosed ? 0.0 : s[(n + q) * W];
      }
      ALTRO_UNROLL
      for (int q = 0; q < n; ++q) zn[q * W] = x[q];
      if (kClosed) {
        ALTRO_UNROLL
        for (int q = 0; q < m; ++q) zn[(n + q) * W] = u[q];
      }
      double v;
      Jsum += knot_cost<n, m, W>(D, k, x, u, s + off_lam * W, pen, &v);
      vmax = fmax(vmax, v);
      if (k < N) {
        if (kClosed) gs += g;
        double xn[n];
        rk4_step<M>(mp, x, u, D.h(k), xn);
        ALTRO_UNROLL
        for (int q = 0; q < n; ++q) x[q] = xn[q];
        if (kClosed && o.check_forwardpass_bounds) {  // ilqr.hpp:484-495
          double sx = 0.0, su = 0.0;
          ALTRO_UNROLL
          for (int q = 0; q < n; ++q) sx += x[q] * x[q];
          ALTRO_UNROLL
          for (int q = 0; q < m; ++q) su += u[q] * u[q];
          if (sx > o.state_max_sq) {  // sqrt(sx) > state_max, see DevOptions
            status = kStateLimit;
            ok = false;
          } else if (su > o.control_max_sq) {
            status = kControlLimit;
            ok = false;
          }
        }
      }
    }
  }
  if (ok && kClosed) status = kUnsolved;  // ilqr.hpp:497
  J = Jsum;
  gsum = gs;
  if (viol_out) *viol_out = vmax;
  return ok;
}

// Cost() of a stored trajectory (no dynamics): ilqr.hpp:326-334, 758-763.
template <class M, int W, bool kStoreCosts>
__device__ __forceinline__ double sweep_cost(const Lane<M, W>& L, double* stg, bool active, int sel,
                                             double penalty, double* viol_out) {
  constexpr int n = M::n, m = M::m, nz = n + m;
  const int N = L.P.N, pmax = L.P.pmax;
  const int R = nz + pmax;
  const AlPen pen(penalty);
  auto issue = [&](int k) {
    double* s = stg + (k & 1) * R * W + L.i;
    if (L.a == 0) {
      const double* g = L.z(sel, k);
      ALTRO_UNROLL
      for (int r = 0; r < nz; ++r) cp_async8(s + r * W, g + r * W);
      if (pmax > 0) {
        g = L.lam(k);
        for (int r = 0; r < pmax; ++r) cp_async8(s + (nz + r) * W, g + r * W);
      }
    }
    cp_async_commit();
  };
  __syncwarp();
  issue(0);
  double J = 0.0, vmax = 0.0;
  for (int k = 0; k <= N; ++k) {
    cp_async_wait_all();
    __syncwarp();
    if (k < N) issue(k + 1);
    const double* s = stg + (k & 1) * R * W + L.i;
    if (active) {
      double x[n], u[m];
      ALTRO_UNROLL
      for (int q = 0; q < n; ++q) x[q] = s[q * W];
      ALTRO_UNROLL
      for (int q = 0; q < m; ++q) u[q] = s[(n + q) * W];
      double v;
      const double c = knot_cost<n, m, W>(L.D, k, x, u, s + nz * W, pen, &v);
      if (kStoreCosts) *L.costs(k) = c;
      J += c;
      vmax = fmax(vmax, v);
    }
  }
  if (viol_out) *viol_out = vmax;
  return J;
}

// Cost + dynamics expansion of knot k in registers (UpdateExpansionsBlock body, ilqr.hpp:670-677).
// LS = stride between the dual rows behind `lam`.
template <class M, int LS, bool kWithDynamics = true>
__device__ __forceinline__ void knot_expansion(const Desc& D, int N, int k, const double* x,
                                               const double* u, const double* lam, double penalty,
                                               double* A, double* B, double* lxx, double* lxu,
                                               double* luu, double* lx, double* lu) {
  constexpr int n = M::n, m = M::m;
  const double* C = D.cost(k);
  ALTRO_UNROLL
  for (int q = 0; q < n * n; ++q) lxx[q] = C[q];  // quadratic_cost.cpp:20-28
  ALTRO_UNROLL
  for (int q = 0; q < m * m; ++q) luu[q] = C[n * n + q];
  ALTRO_UNROLL
  for (int q = 0; q < n * m; ++q) lxu[q] = C[n * n + m * m + q];
  quad_gradient<n, m>(C, x, u, lx, lu);
  const ConSet& cs = D.conset(k);
  if (cs.nblocks > 0) al_expansion<n, m, LS>(cs, x, u, lam, penalty, lxx, lxu, luu, lx, lu);
  if (kWithDynamics && k < N) rk4_jacobian<M>(D.params(), x, u, D.h(k), A, B);
}

// ------------------------------------------------------------------------------------------
// Backward sweep: expansions + Riccati recursion, with the regularisation restart loop of
// ilqr.hpp:396-445 (Q4, Q5).  gsum returns sum_k max_i |d_i|/(|u_i|+1) for the CURRENT controls
// (used only when the line search fails and Z_ stays as it is).
//
// The recursion itself is serial in k and runs on the instance's lead lane (a = 0).  The
// expensive, k-independent part of UpdateExpansions — the RK4 Jacobian [A|B] with its 8 sin/cos
// evaluations — is produced G knots at a time by the instance's G lane groups (group a takes
// knot top-a) and handed to the lead lane with warp shuffles, like the reference's thread pool
// spreads UpdateExpansionsBlock over knot ranges (ilqr.hpp:354-365, 670-677).
// `active` is the per-instance flag (the same on all G lanes of an instance).
// ------------------------------------------------------------------------------------------
template <class M, int W, bool kStoreCtg>
__device__ __forceinline__ void sweep_backward(const Lane<M, W>& L, double* stg, bool active,
                                               int zsel, double penalty, double& reg,
                                               double& dreg, double& dV0, double& dV1,
                                               int& status, double& gsum, double* reg_log = nullptr) {
  constexpr int n = M::n, m = M::m, nz = n + m, G = Lane<M, W>::G;
  const int N = L.P.N, pmax = L.P.pmax;
  const DevOptions& o = L.P.opt;
  const int R = nz + pmax;                  // staged rows per knot
  const bool lead = L.a == 0;
  // staging area of (buffer, group): R rows of W doubles
  auto area = [&](int buf, int g) { return stg + (static_cast<size_t>(buf) * G + g) * R * W + L.i; };
  auto issue_block = [&](int top, int cnt, int buf) {
    if (L.a < cnt) {
      const int kk = top - L.a;
      double* s = area(buf, L.a);
      const double* g = L.z(zsel, kk);
      ALTRO_UNROLL
      for (int r = 0; r < nz; ++r) cp_async8(s + r * W, g + r * W);
      if (pmax > 0) {
        g = L.lam(kk);
        for (int r = 0; r < pmax; ++r) cp_async8(s + (nz + r) * W, g + r * W);
      }
    }
    cp_async_commit();
  };
  int max_reg_count = 0;
  if (active && lead) {
    dV0 = 0.0;
    dV1 = 0.0;
  }
  bool repeat = active;  // per instance (replicated on its G lanes)
  double gs = 0.0;
  while (__any_sync(kFull, repeat)) {
    double P[n * n], p[n];
    bool live = repeat;  // instance still descending in this pass
    gs = 0.0;
    int top = N, cnt = 1, buf = 0;
    __syncwarp();
    issue_block(top, cnt, buf);
    while (top >= 0) {
      cp_async_wait_all();
      __syncwarp();
      const int ntop = (top == N) ? N - 1 : top - G;
      const int ncnt = (ntop + 1 < G) ? ntop + 1 : G;
      if (ntop >= 0) issue_block(ntop, ncnt, buf ^ 1);
      if (top == N) {
        if (live && lead) {  // CalcTerminalCostToGo: P = lxx, p = lx of the terminal expansion
          const double* s = area(buf, 0);
          double x[n], u[m], A[1], B[1], lxu[n * m], luu[m * m], lu[m];
          ALTRO_UNROLL
          for (int q = 0; q < n; ++q) x[q] = s[q * W];
          ALTRO_UNROLL
          for (int q = 0; q < m; ++q) u[q] = s[(n + q) * W];
          knot_expansion<M, W, false>(L.D, N, N, x, u, s + nz * W, penalty, A, B, P, lxu, luu, p, lu);
          if (kStoreCtg) {
            double* c = L.ctg(N);
            ALTRO_UNROLL
            for (int q = 0; q < n * n; ++q) c[q * W] = P[q];
            ALTRO_UNROLL
            for (int q = 0; q < n; ++q) c[(n * n + q) * W] = p[q];
          }
        }
      } else {
        // phase 1: every group computes the dynamics Jacobian of its knot
        double A[n * n], B[n * m];
        if (live && L.a < cnt) {
          const double* s = area(buf, L.a);
          double x[n], u[m];
          ALTRO_UNROLL
          for (int q = 0; q < n; ++q) x[q] = s[q * W];
          ALTRO_UNROLL
          for (int q = 0; q < m; ++q) u[q] = s[(n + q) * W];
          rk4_jacobian<M>(L.D.params(), x, u, L.D.h(top - L.a), A, B);
        } else {
          ALTRO_UNROLL
          for (int q = 0; q < n * n; ++q) A[q] = 0.0;
          ALTRO_UNROLL
          for (int q = 0; q < n * m; ++q) B[q] = 0.0;
        }
        // phase 2: the lead lane walks the block's knots in order
        for (int j = 0; j < cnt; ++j) {
          const int kk = top - j;
          const int src = L.i + W * j;
          double Aj[n * n], Bj[n * m];
          ALTRO_UNROLL
          for (int q = 0; q < n * n; ++q) Aj[q] = __shfl_sync(kFull, A[q], src);
          ALTRO_UNROLL
          for (int q = 0; q < n * m; ++q) Bj[q] = __shfl_sync(kFull, B[q], src);
          if (live && lead) {
            const double* s = area(buf, j);
            double x[n], u[m], lxx[n * n], lxu[n * m], luu[m * m], lx[n], lu[m], dA[1], dB[1];
            ALTRO_UNROLL
            for (int q = 0; q < n; ++q) x[q] = s[q * W];
            ALTRO_UNROLL
            for (int q = 0; q < m; ++q) u[q] = s[(n + q) * W];
            knot_expansion<M, W, false>(L.D, N, kk, x, u, s + nz * W, penalty, dA, dB, lxx, lxu, luu, lx, lu);
            double K[m * n], d[m];
            const bool ok = riccati_step<n, m>(Aj, Bj, lxx, lxu, luu, lx, lu, P, p, reg, K, d, &dV0, &dV1);
            if (!ok) {  // ilqr.hpp:409-427
              increase_reg(o, reg, dreg);
              if (reg >= o.bp_reg_max) max_reg_count++;
              if (max_reg_count >= o.bp_reg_fail_threshold) {
                status = kBackwardPassRegularizationFailed;
                repeat = false;
              }
              live = false;
            } else {
              double* pk = L.kd(kk);
              ALTRO_UNROLL
              for (int q = 0; q < m * n; ++q) pk[q * W] = K[q];
              ALTRO_UNROLL
              for (int q = 0; q < m; ++q) pk[(m * n + q) * W] = d[q];
              double g = fabs(d[0]) / (fabs(u[0]) + 1);
              ALTRO_UNROLL
              for (int q = 1; q < m; ++q) g = fmax(g, fabs(d[q]) / (fabs(u[q]) + 1));
              gs += g;
              if (kStoreCtg) {
                double* c = L.ctg(kk);
                ALTRO_UNROLL
                for (int q = 0; q < n * n; ++q) c[q * W] = P[q];
                ALTRO_UNROLL
                for (int q = 0; q < n; ++q) c[(n * n + q) * W] = p[q];
              }
              if (kk == 0) repeat = false;
            }
          }
        }
      }
      // the lead lane's verdict is the instance's verdict
      live = __shfl_sync(kFull, live, L.i);
      repeat = __shfl_sync(kFull, repeat, L.i);
      top = ntop;
      cnt = ncnt;
      buf ^= 1;
    }
  }
  if (active && lead) {
    gsum = gs;
    if (reg_log) *reg_log = reg;  // stats.Log("reg", rho_) precedes the decrease, ilqr.hpp:442
    decrease_reg(o, reg, dreg);   // ilqr.hpp:443-444
  }
}

// ------------------------------------------------------------------------------------------
// Dual update on the constraint values of trajectory buffer `sel` (Q8 decides which), returns
// the max violation of those values.  constraint_values.hpp:192-194, 216-221.
// ------------------------------------------------------------------------------------------
template <class M, int W>
__device__ __forceinline__ double sweep_dual(const Lane<M, W>& L, bool active, int sel,
                                             double penalty, bool update) {
  constexpr int n = M::n, m = M::m;
  const int N = L.P.N;
  double vmax = 0.0;
  if (L.P.pmax == 0 || !active) return vmax;
  for (int k = 0; k <= N; ++k) {
    const ConSet& cs = L.D.conset(k);
    if (cs.nblocks == 0) continue;
    const double* zc = L.z(sel, k);
    double* lam = L.lam(k);
    double x[n], u[m];
    ALTRO_UNROLL
    for (int q = 0; q < n; ++q) x[q] = zc[q * W];
    ALTRO_UNROLL
    for (int q = 0; q < m; ++q) u[q] = zc[(n + q) * W];
    for (int bi = 0; bi < cs.nblocks; ++bi) {
      const ConBlock& b = cs.blk[bi];
      const BlockHdr hd(cs.hdr[bi]);
      const bool eq = hd.eq;
      double* lb = lam + hd.row0 * W;
      for_row_chunks<n, m>(hd, b, x, u, [&](const double* c, const int* ic, const bool* ok) {
        double l[kChunk];
        if (update) {
          ALTRO_UNROLL
          for (int q = 0; q < kChunk; ++q) l[q] = lb[ic[q] * W];
        }
        ALTRO_UNROLL
        for (int q = 0; q < kChunk; ++q) {
          if (ok[q]) {
            if (update) {
              const double arg = l[q] - penalty * c[q];
              lb[ic[q] * W] = eq ? arg : neg_part(arg);
            }
            vmax = fmax(vmax, eq ? fabs(c[q]) : fabs(c[q] - neg_part(c[q])));
          }
        }
      });
    }
  }
  return vmax;
}

// ------------------------------------------------------------------------------------------
// ForwardPass line search (ilqr.hpp:512-558), G step lengths per round.  Try t of an instance
// uses alpha_t = 1 / factor^t exactly as the sequential `alpha /= factor`; group a of the warp
// evaluates try (round*G + a).  The winner is the lowest try index that passes the acceptance
// test, i.e. the one the sequential search stops at; tries after it are speculative and only
// touch their own candidate buffer.  Per-instance results are replicated on all G lanes.
// ------------------------------------------------------------------------------------------
struct LineSearchResult {
  bool success;
  int slot;  // candidate buffer slot of the accepted trajectory
  double J, alpha, z, gsum;
};

template <class M, int W>
__device__ __forceinline__ LineSearchResult line_search(const Lane<M, W>& L, double* stg, bool run,
                                                        int zsel, double penalty, double J0,
                                                        double dV0, double dV1, int& status,
                                                        double& csrc) {
  constexpr int G = Lane<M, W>::G;
  const DevOptions& o = L.P.opt;
  LineSearchResult r;
  r.success = false;
  r.slot = 0;
  r.J = J0;
  r.alpha = 1.0;
  r.z = -1.0;
  r.gsum = 0.0;
  double alpha_base = 1.0;  // step length of try index `done`
  int done = 0;
  bool searching = run && (o.line_search_max_iterations > 0);
  while (__any_sync(kFull, searching)) {
    double alpha = alpha_base;
    for (int j = 0; j < L.a; ++j) alpha /= o.line_search_decrease_factor;
    const bool mine = searching && (done + L.a < o.line_search_max_iterations);
    double J = 0.0, gs = 0.0, z = -1.0;
    int st_try = status;
    const bool ok = sweep_forward<M, W, true>(L, stg, mine, zsel, Lane<M, W>::cand(zsel, L.a), alpha,
                                              penalty, J, gs, nullptr, st_try);
    bool acc = false;
    if (mine && ok) {
      const double expected = -alpha * (dV0 + alpha * dV1);
      if (expected > 0.0) z = (J0 - J) / expected;
      acc = o.line_search_lower_bound <= z && z <= o.line_search_upper_bound && J < J0;
    }
    const unsigned accm = __ballot_sync(kFull, acc);
    const unsigned okm = __ballot_sync(kFull, mine && ok);
    const unsigned minem = __ballot_sync(kFull, mine);
    int win = -1, last = -1, lastok = -1;
    ALTRO_UNROLL
    for (int g = G - 1; g >= 0; --g) {
      const unsigned bit = 1u << (L.i + W * g);
      if (accm & bit) win = g;
      if ((minem & bit) && last < 0) last = g;
      if ((okm & bit) && lastok < 0) lastok = g;
    }
    const int src_win = L.i + W * (win < 0 ? 0 : win);
    const double Jw = __shfl_sync(kFull, J, src_win);
    const double aw = __shfl_sync(kFull, alpha, src_win);
    const double zw = __shfl_sync(kFull, z, src_win);
    const double gw = __shfl_sync(kFull, gs, src_win);
    const int st_last = __shfl_sync(kFull, st_try, L.i + W * (last < 0 ? 0 : last));
    const double a_lastok = __shfl_sync(kFull, alpha, L.i + W * (lastok < 0 ? 0 : lastok));
    if (searching) {
      if (win >= 0) {
        r.success = true;
        r.slot = win;
        r.J = Jw;
        r.alpha = aw;
        r.z = zw;
        r.gsum = gw;
        status = kUnsolved;  // the accepted rollout ran to the end (ilqr.hpp:497)
        searching = false;
      } else {
        if (last >= 0) status = st_last;      // status_ left by the last executed rollout
        if (lastok >= 0) csrc = a_lastok;     // Cost(*Zbar_) refreshed the stored constraint values (Q8)
        done += G;
        ALTRO_UNROLL
        for (int j = 0; j < G; ++j) alpha_base /= o.line_search_decrease_factor;
        if (done >= o.line_search_max_iterations) searching = false;
      }
    }
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// Tail of one inner iteration (ilqr.hpp:300-313 after ForwardPass): bookkeeping of the line
// search outcome, UpdateConvergenceStatistics (:568-587, Q15) and IsDone (:597-619).  Shared by
// the fused engine (k_solve) and the phased engine (phased.cuh) so both follow the same rules.
//
// Stall skip (option skip_repeated_iterations, off by default): when a line search fails
// completely, Z_, the duals and the penalty stay as they are, so the next iteration differs
// from this one only through (reg, dreg) at its entry.  If the failed iteration leaves (reg,
// dreg) exactly where they were at its own entry, every later iteration of this iLQR solve is
// a bit-identical repetition (same gains, same failed search, dJ = 0, same grad): the remaining
// ones are accounted for — counters, status — without being executed.  Results are identical.
// ------------------------------------------------------------------------------------------
struct InnerTail {
  bool success;
  int new_zsel;        // buffer of the accepted trajectory
  double J, alpha, z, gsum_ls, gsum_bwd;
  double reg_in, dreg_in;  // regularisation at the entry of this iteration (before BackwardPass)
};

__device__ __forceinline__ void finish_inner(const DevOptions& o, int N, int mode, const InnerTail& t,
                                             int& zsel, double& J0, double& cost_cur, double& cost_prev,
                                             double initial_cost, double& alpha_stat, double& z_stat,
                                             double& csrc, double& grad, double& dJ, double& reg,
                                             double& dreg, int& it_inner, int& it_total, int& st,
                                             int& phase, int& lsfail) {
  lsfail = t.success ? 0 : 1;
  if (t.success) {
    zsel = t.new_zsel;  // (*Z_) = (*Zbar_)
    J0 = t.J;
    cost_cur = t.J;  // stats.Log("cost"/"alpha"/"z")
    alpha_stat = t.alpha;
    z_stat = t.z;
    csrc = -1.0;  // the stored constraint values are those of the new Z_
    grad = t.gsum_ls / static_cast<double>(N);
  } else {
    increase_reg(o, reg, dreg);
    grad = t.gsum_bwd / static_cast<double>(N);
  }
  // UpdateConvergenceStatistics, ilqr.hpp:568-587 (Q15)
  dJ = (it_inner == 0) ? (initial_cost - cost_cur) : (cost_prev - cost_cur);
  it_inner++;
  it_total++;
  cost_prev = cost_cur;  // NewIteration() carry-forward (Q6)
  // IsDone, ilqr.hpp:597-619
  bool done = true;
  if (dJ < o.cost_tolerance && grad < o.gradient_tolerance) {
    st = kSolved;
  } else if (it_inner >= o.max_iterations_inner) {
    st = kMaxInnerIterations;
  } else if (it_total >= o.max_iterations_total) {
    st = kMaxIterations;
  } else if (st == kUnsolved) {
    done = false;
  }
  if (!done && o.skip_repeated_iterations && !t.success && reg == t.reg_in && dreg == t.dreg_in) {
    // iterations it_inner+1, ... repeat this one with dJ = 0 and the same grad
    dJ = 0.0;
    if (dJ < o.cost_tolerance && grad < o.gradient_tolerance) {
      it_inner += 1;
      it_total += 1;
      st = kSolved;
    } else {
      const int left_inner = o.max_iterations_inner - it_inner;
      const int left_total = o.max_iterations_total - it_total;
      const int skip = left_inner < left_total ? left_inner : left_total;
      it_inner += skip;
      it_total += skip;
      st = (it_inner >= o.max_iterations_inner) ? kMaxInnerIterations : kMaxIterations;
    }
    done = true;
  }
  if (done) phase = (mode == 1) ? kPhOuter : kPhDone;
}

// ------------------------------------------------------------------------------------------
// k_solve: the whole AugmentedLagrangianiLQR::Solve (mode 1) / iLQR::Solve (mode 0) as a
// resumable per-instance state machine, one warp per tile.  A "slot" lets every live instance
// of the tile do its pending outer-loop work (dual/penalty update, rollout of the next iLQR
// solve) and then one full inner iteration (backward sweep, parallel line search, convergence
// test).  Instances of a tile are NOT held in lock-step across iLQR solves: each advances
// through its own phases.  After `budget` slots the state is written back; the host relaunches
// (optionally after packing the unfinished instances densely) until all are reported.
// ------------------------------------------------------------------------------------------
constexpr int kSolveWarps = ALTRO_SOLVE_WARPS;

// parts: 1 = outer step + solve start, 2 = inner iteration (3 = everything; the step-wise tests
// run the halves separately).
template <class M, int W>
__global__ void __launch_bounds__(kSolveWarps* kWarp, (W <= 4 ? ALTRO_SOLVE_MINB_NARROW : ALTRO_SOLVE_MINB)) k_solve(SolverParams P, int mode,
                                                                               int budget, int parts) {
  extern __shared__ __align__(128) char smem[];
  copy_blob(P.blob, smem, P.blob_bytes);
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int tile = blockIdx.x * kSolveWarps + warp;
  if (tile >= P.T) return;
  double* stg = reinterpret_cast<double*>(smem + ((P.blob_bytes + 15) / 16) * 16) +
                static_cast<size_t>(warp) * stage_doubles<M>(P.pmax, W);
  const Lane<M, W> L(P, smem, tile, lane);
  const DevOptions& o = P.opt;
  const int N = P.N;
  const bool valid = L.valid;
  const bool lead = valid && L.a == 0;  // the lane that owns the instance's serial work
  const bool has_con = P.pmax > 0;

  // per-instance state, replicated on the G lanes of the instance
  double penalty = 1.0, reg = 0.0, dreg = 0.0, dV0 = 0.0, dV1 = 0.0;
  double cost_cur = 0.0, cost_prev = 0.0, initial_cost = 0.0, viol = 0.0;
  double dJ = 0.0, grad = 0.0, alpha_stat = 0.0, z_stat = 0.0, csrc = -1.0, J0 = 0.0;
  int zsel = 0, it_inner = 0, it_outer = 0, it_total = 0, st = kUnsolved, st_al = kUnsolved;
  int phase = kPhReported, lsfail = 0;
  if (valid) {
    penalty = L.sc(S_PENALTY);
    reg = L.sc(S_REG);
    dreg = L.sc(S_DREG);
    dV0 = L.sc(S_DV0);
    dV1 = L.sc(S_DV1);
    cost_cur = L.sc(S_COST_CUR);
    cost_prev = L.sc(S_COST_PREV);
    initial_cost = L.sc(S_INITIAL_COST);
    viol = L.sc(S_VIOL);
    dJ = L.sc(S_DJ);
    grad = L.sc(S_GRAD);
    alpha_stat = L.sc(S_ALPHA);
    z_stat = L.sc(S_ZRATIO);
    csrc = L.sc(S_CSRC_ALPHA);
    J0 = L.sc(S_J0);
    zsel = L.is(I_ZSEL);
    it_inner = L.is(I_ITERS_INNER);
    it_outer = L.is(I_ITERS_OUTER);
    it_total = L.is(I_ITERS_TOTAL);
    st = L.is(I_STATUS);
    st_al = L.is(I_STATUS_AL);
    phase = L.is(I_PHASE);
    lsfail = L.is(I_LSFAIL);
  }
  const bool was_reported = phase >= kPhReported;
  // history rows (P.HIST): the value stats.max_penalty carries into the rows of the current iLQR solve — logged by
  // Init() and after every dual update, BEFORE UpdatePenalties (al_solver.hpp:298, 362).  Lives in a register: a
  // solver that records history solves in one launch (altro_b200.cu).
  double pen_logged = 0.0;
  if (phase == kPhAlInit) {  // AugmentedLagrangianiLQR::Init, al_solver.hpp:287-302 (Q10)
    if (o.reset_duals && has_con && L.a == 0) {
      for (int k = 0; k <= N; ++k) {
        double* lam = L.lam(k);
        for (int r = 0; r < P.pmax; ++r) lam[r * W] = 0.0;
      }
    }
    if (o.initial_penalty > 0) penalty = o.initial_penalty;
    it_outer = 0;  // stats.Reset()
    it_total = 0;
    cost_cur = 0.0;
    cost_prev = 0.0;
    st_al = kUnsolved;
    phase = kPhSolveStart;
    pen_logged = has_con ? penalty : 0.0;
  }
  __syncwarp();

  for (int slot = 0; slot < budget; ++slot) {
    if (!__any_sync(kFull, phase < kPhDone)) break;
    // ------------- AL outer step for instances whose iLQR solve just ended ------------------
    // al_solver.hpp:313-333: UpdateDuals, UpdateConvergenceStatistics, IsDone, UpdatePenalties
    const bool outer = phase == kPhOuter && (parts & 1);
    if (__any_sync(kFull, outer)) {
      // Q8: after a fully failed line search the stored constraint values are those of the last
      // evaluated candidate; regenerate it (slot 0) so the dual update sees the same values.
      const bool regen = outer && csrc >= 0.0;
      if (__any_sync(kFull, regen)) {
        double Jt, gt;
        int stt = st;
        sweep_forward<M, W, true>(L, stg, regen && L.a == 0, zsel, Lane<M, W>::cand(zsel, 0), csrc, penalty,
                                  Jt, gt, nullptr, stt);
        __syncwarp();
      }
      const int src = regen ? Lane<M, W>::cand(zsel, 0) : zsel;
      double v = sweep_dual<M, W>(L, outer && L.a == 0, src, penalty, /*update=*/true);
      v = __shfl_sync(kFull, v, L.i);
      if (outer) {
        viol = v;
        it_outer++;
        const double max_penalty = has_con ? penalty : 0.0;
        pen_logged = max_penalty;
        phase = kPhDone;  // IsDone, al_solver.hpp:368-401 (Q16)
        if (st != kSolved) {
          st_al = st;
        } else if (viol < o.constraint_tolerance) {
          st_al = kSolved;
        } else if (max_penalty > o.maximum_penalty) {
          st_al = kMaxPenalty;
        } else if (it_outer >= o.max_iterations_outer) {
          st_al = kMaxOuterIterations;
        } else if (it_total >= o.max_iterations_total) {
          st_al = kMaxIterations;
        } else {
          penalty *= o.penalty_scaling;  // UpdatePenalties
          phase = kPhSolveStart;
        }
      }
      __syncwarp();
    }
    // ------------- iLQR::Solve() entry: SolveSetup, Rollout, initial Cost -------------------
    const bool start = phase == kPhSolveStart && (parts & 1);
    if (__any_sync(kFull, start)) {
      if (start) {
        it_inner = 0;  // SolveSetup :629-645, ResetInternalVariables :680-690
        lsfail = 0;
        st = kUnsolved;
        reg = o.bp_reg_initial;
        dreg = 0.0;
        dV0 = dV1 = 0.0;
      }
      double Jr = 0.0, gs = 0.0;
      int stt = st;
      sweep_forward<M, W, false>(L, stg, start && L.a == 0, zsel, zsel, 0.0, penalty, Jr, gs, nullptr, stt);
      Jr = __shfl_sync(kFull, Jr, L.i);
      if (start) {
        J0 = Jr;
        initial_cost = Jr;
        csrc = -1.0;
        phase = (o.max_iterations_inner > 0) ? kPhInner : (mode == 1 ? kPhOuter : kPhDone);
      }
    }
    // ------------- one inner iteration, ilqr.hpp:300-313 ------------------------------------
    const bool run = phase == kPhInner && (parts & 2);
    if (__any_sync(kFull, run)) {
      double gs_bwd = 0.0;
      if (run) csrc = -1.0;  // UpdateExpansions evaluates every constraint at Z_
      const double reg_in = reg, dreg_in = dreg;
      double reg_logged = reg;
      sweep_backward<M, W, false>(L, stg, run, zsel, penalty, reg, dreg, dV0, dV1, st, gs_bwd, &reg_logged);
      reg = __shfl_sync(kFull, reg, L.i);
      dreg = __shfl_sync(kFull, dreg, L.i);
      dV0 = __shfl_sync(kFull, dV0, L.i);
      dV1 = __shfl_sync(kFull, dV1, L.i);
      st = __shfl_sync(kFull, st, L.i);
      gs_bwd = __shfl_sync(kFull, gs_bwd, L.i);
      const LineSearchResult ls = line_search<M, W>(L, stg, run, zsel, penalty, J0, dV0, dV1, st, csrc);
      if (run) {
        InnerTail t;
        t.success = ls.success;
        t.new_zsel = Lane<M, W>::cand(zsel, ls.slot);
        t.J = ls.J;
        t.alpha = ls.alpha;
        t.z = ls.z;
        t.gsum_ls = ls.gsum;
        t.gsum_bwd = gs_bwd;
        t.reg_in = reg_in;
        t.dreg_in = dreg_in;
        finish_inner(o, N, mode, t, zsel, J0, cost_cur, cost_prev, initial_cost, alpha_stat, z_stat, csrc,
                     grad, dJ, reg, dreg, it_inner, it_total, st, phase, lsfail);
      }
      if (P.HIST != nullptr) {
        // One SolverStats row per inner iteration (UpdateConvergenceStatistics, ilqr.hpp:568-587).  "violations" is
        // MaxViolationStored(): of the accepted trajectory, or — after a fully failed search — of the last
        // candidate evaluated (Q8), which is regenerated here exactly like the dual update regenerates it.
        const bool rejected = run && csrc >= 0.0;
        double vs = 0.0, vr = 0.0;
        if (__any_sync(kFull, run && !rejected))
          sweep_cost<M, W, false>(L, stg, run && !rejected && L.a == 0, zsel, penalty, &vs);
        if (__any_sync(kFull, rejected)) {
          double Jt, gt;
          int stt = st;
          sweep_forward<M, W, true>(L, stg, rejected && L.a == 0, zsel, Lane<M, W>::cand(zsel, 0), csrc, penalty,
                                    Jt, gt, &vr, stt);
          __syncwarp();
        }
        const int orig = valid ? L.is(I_ORIG) : 0;
        if (run && lead && orig < P.hist_instances && it_total >= 1 && it_total <= P.hist_rows) {
          double* h = P.HIST + (static_cast<size_t>(orig) * P.hist_rows + (it_total - 1)) * kHistCols;
          h[kHistCost] = cost_cur;
          h[kHistAlpha] = alpha_stat;
          h[kHistZ] = z_stat;
          h[kHistGradient] = grad;
          h[kHistCostDecrease] = dJ;
          h[kHistRegularization] = reg_logged;
          h[kHistViolations] = rejected ? vr : vs;
          h[kHistMaxPenalty] = pen_logged;
        }
      }
    }
  }

  // Cost() of the final trajectory under the final duals/penalty, as printed by
  // perf/benchmark_unicycle.cpp:73-74 — once, when the instance terminates.
  const bool fin = phase == kPhDone;
  double Jf = 0.0, vf = 0.0;
  if (__any_sync(kFull, fin)) Jf = sweep_cost<M, W, false>(L, stg, fin && L.a == 0, zsel, penalty, &vf);
  if (lead && !was_reported) {
    if (fin) {
      if (mode == 0) viol = vf;  // == Cost(); GetMaxViolation()
      L.sc(S_COST) = Jf;
      csrc = -1.0;
      phase = kPhReported;
    }
    L.sc(S_REG) = reg;
    L.sc(S_DREG) = dreg;
    L.sc(S_DV0) = dV0;
    L.sc(S_DV1) = dV1;
    L.sc(S_PENALTY) = penalty;
    L.sc(S_VIOL) = viol;
    L.sc(S_INITIAL_COST) = initial_cost;
    L.sc(S_COST_CUR) = cost_cur;
    L.sc(S_COST_PREV) = cost_prev;
    L.sc(S_DJ) = dJ;
    L.sc(S_GRAD) = grad;
    L.sc(S_ALPHA) = alpha_stat;
    L.sc(S_ZRATIO) = z_stat;
    L.sc(S_CSRC_ALPHA) = csrc;
    L.sc(S_J0) = J0;
    L.is(I_STATUS) = st;
    L.is(I_STATUS_AL) = st_al;
    L.is(I_ITERS_INNER) = it_inner;
    L.is(I_ITERS_OUTER) = it_outer;
    L.is(I_ITERS_TOTAL) = it_total;
    L.is(I_ZSEL) = zsel;
    L.is(I_PHASE) = phase;
    L.is(I_LSFAIL) = lsfail;
    if (phase < kPhReported) atomicAdd(&P.counters[0], 1);
  }
}

// ------------------------------------------------------------------------------------------
// Re-packing of unfinished instances between k_solve launches (runtime tile widths).
// ------------------------------------------------------------------------------------------
// list[] = slots (in `src`) of the unfinished instances, `total` of them: those whose last line
// search failed completely are packed from the front, the others from the back, so that tiles of
// the re-packed workspace are homogeneous (a tile pays for its slowest member's line-search
// rounds).  Counters: [1] front cursor, [2] back cursor.
__global__ void k_list_unfinished(SolverParams src, int* __restrict__ list, int total) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= src.B) return;
  if (src.is[static_cast<size_t>(I_PHASE) * src.Bp + b] < kPhReported) {
    if (src.is[static_cast<size_t>(I_LSFAIL) * src.Bp + b]) {
      list[atomicAdd(&src.counters[1], 1)] = b;
    } else {
      list[total - 1 - atomicAdd(&src.counters[2], 1)] = b;
    }
  }
}

// dst slot j <- src slot list[j]: current trajectory (into buffer 0), gains, duals, x0, scalars.
// dir = 0: pack (src -> dst, marks the source slot kPhMoved); dir = 1: unpack results of reported
// instances back (dst slot I_ORIG <- src slot j), used at the end of a solve.
__global__ void k_move_instances(SolverParams src, SolverParams dst, const int* __restrict__ list,
                                 int count, int dir) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;  // knot, N+1 = the scalar record
  if (j >= count) return;
  const int n = src.n, m = src.m, nz = n + m, nkd = m * n + m, N = src.N, pmax = src.pmax;
  int bs, bd;
  if (dir == 0) {
    bs = list[j];
    bd = j;
  } else {
    bs = j;
    if (src.is[static_cast<size_t>(I_PHASE) * src.Bp + bs] != kPhReported) return;
    bd = src.is[static_cast<size_t>(I_ORIG) * src.Bp + bs];
  }
  const int Ws = src.W, Wd = dst.W;
  const int ts = bs / Ws, is_ = bs % Ws, td = bd / Wd, id = bd % Wd;
  if (k <= N) {
    const int zs = src.is[static_cast<size_t>(I_ZSEL) * src.Bp + bs];
    const double* z = src.Z[zs] + (static_cast<size_t>(ts) * (N + 1) + k) * nz * Ws + is_;
    double* zo = dst.Z[0] + (static_cast<size_t>(td) * (N + 1) + k) * nz * Wd + id;
    for (int q = 0; q < nz; ++q) zo[q * Wd] = z[q * Ws];
    if (k < N) {
      const double* g = src.KD + (static_cast<size_t>(ts) * N + k) * nkd * Ws + is_;
      double* go = dst.KD + (static_cast<size_t>(td) * N + k) * nkd * Wd + id;
      for (int q = 0; q < nkd; ++q) go[q * Wd] = g[q * Ws];
    }
    if (pmax > 0) {
      const double* l = src.LAM + (static_cast<size_t>(ts) * (N + 1) + k) * pmax * Ws + is_;
      double* lo = dst.LAM + (static_cast<size_t>(td) * (N + 1) + k) * pmax * Wd + id;
      for (int q = 0; q < pmax; ++q) lo[q * Wd] = l[q * Ws];
    }
  } else {
    const double* x = src.X0 + static_cast<size_t>(ts) * n * Ws + is_;
    double* xo = dst.X0 + static_cast<size_t>(td) * n * Wd + id;
    for (int q = 0; q < n; ++q) xo[q * Wd] = x[q * Ws];
    for (int f = 0; f < S_NUM; ++f)  // (the candidate buffers do not move with the instance)
      dst.sc[static_cast<size_t>(f) * dst.Bp + bd] = (f == S_CAND_ALPHA) ? -1.0 : src.sc[static_cast<size_t>(f) * src.Bp + bs];
    for (int f = 0; f < I_NUM; ++f) {
      int v = src.is[static_cast<size_t>(f) * src.Bp + bs];
      if (f == I_ZSEL) v = 0;
      if (dir == 0 && f == I_ORIG) v = src.is[static_cast<size_t>(I_ORIG) * src.Bp + bs];
      if (dir == 1 && f == I_ORIG) v = bd;
      dst.is[static_cast<size_t>(f) * dst.Bp + bd] = v;
    }
    if (dir == 0) src.is[static_cast<size_t>(I_PHASE) * src.Bp + bs] = kPhMoved;
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
  kPhaseAlInit,
};

template <class M, int W>
__global__ void __launch_bounds__(kSolveWarps* kWarp) k_phase(SolverParams P, int phase) {
  extern __shared__ __align__(128) char smem[];
  copy_blob(P.blob, smem, P.blob_bytes);
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int tile = blockIdx.x * kSolveWarps + warp;
  if (tile >= P.T) return;
  double* stg = reinterpret_cast<double*>(smem + ((P.blob_bytes + 15) / 16) * 16) +
                static_cast<size_t>(warp) * stage_doubles<M>(P.pmax, W);
  const Lane<M, W> L(P, smem, tile, lane);
  const DevOptions& o = P.opt;
  const int N = P.N;
  const bool valid = L.valid;
  const bool lead = valid && L.a == 0;
  int zsel = valid ? L.is(I_ZSEL) : 0;
  double penalty = valid ? L.sc(S_PENALTY) : 1.0;
  int st = valid ? L.is(I_STATUS) : kUnsolved;
  __syncwarp();
  switch (phase) {
    case kPhaseSolveSetup: {  // SolveSetup(), ilqr.hpp:629-645
      if (lead) {
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
      double J, gs;
      sweep_forward<M, W, false>(L, stg, lead, zsel, zsel, 0.0, penalty, J, gs, nullptr, st);
      break;
    }
    case kPhaseCost: {
      double v;
      const double J = sweep_cost<M, W, true>(L, stg, lead, zsel, penalty, &v);
      if (lead) {
        L.sc(S_COST) = J;
        L.sc(S_VIOL) = v;
        L.sc(S_CSRC_ALPHA) = -1.0;
      }
      break;
    }
    case kPhaseBackwardFused: {
      double reg = valid ? L.sc(S_REG) : 0.0, dreg = valid ? L.sc(S_DREG) : 0.0, dV0 = 0.0, dV1 = 0.0, gs = 0.0;
      __syncwarp();
      sweep_backward<M, W, true>(L, stg, valid, zsel, penalty, reg, dreg, dV0, dV1, st, gs);
      if (lead) {
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
      __syncwarp();
      const LineSearchResult ls = line_search<M, W>(L, stg, valid, zsel, penalty, J0, dV0, dV1, st, csrc);
      if (lead) {
        if (ls.success) {
          L.is(I_ZSEL) = Lane<M, W>::cand(zsel, ls.slot);
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
      if (lead) {
        double gsum = 0.0;
        for (int k = 0; k < N; ++k) {
          const double* zc = L.z(zsel, k);
          const double* pk = L.kd(k);
          double g = 0.0;
          for (int q = 0; q < M::m; ++q) {
            const double gq = fabs(pk[(M::m * M::n + q) * W]) / (fabs(zc[(M::n + q) * W]) + 1);
            g = (q == 0) ? gq : fmax(g, gq);
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
      const double csrc = valid ? L.sc(S_CSRC_ALPHA) : -1.0;
      const bool regen = valid && csrc >= 0.0;
      __syncwarp();
      if (__any_sync(kFull, regen)) {
        double Jt, gt;
        int stt = st;
        sweep_forward<M, W, true>(L, stg, regen && L.a == 0, zsel, Lane<M, W>::cand(zsel, 0), csrc, penalty,
                                  Jt, gt, nullptr, stt);
        __syncwarp();
      }
      const double v = sweep_dual<M, W>(L, lead, regen ? Lane<M, W>::cand(zsel, 0) : zsel, penalty, true);
      if (lead) L.sc(S_VIOL) = v;
      break;
    }
    case kPhasePenalties: {  // UpdatePenalties(), al_solver.hpp:347-355
      if (lead) L.sc(S_PENALTY) = penalty * o.penalty_scaling;
      break;
    }
    case kPhaseAlInit: {  // AugmentedLagrangianiLQR::Init(), al_solver.hpp:287-302 (what a whole solve does first)
      if (lead) {
        if (o.reset_duals && P.pmax > 0) {
          for (int k = 0; k <= N; ++k) {
            double* lam = L.lam(k);
            for (int r = 0; r < P.pmax; ++r) lam[r * W] = 0.0;
          }
        }
        if (o.initial_penalty > 0) L.sc(S_PENALTY) = o.initial_penalty;
        L.is(I_ITERS_OUTER) = 0;  // stats.Reset()
        L.is(I_ITERS_TOTAL) = 0;
        L.sc(S_COST_CUR) = 0.0;
        L.sc(S_COST_PREV) = 0.0;
        L.is(I_STATUS_AL) = kUnsolved;
      }
      break;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Materialised data flow (the reference's UpdateExpansions -> BackwardPass hand-off).
// Record of one (tile, knot): [A | B | lxx | lxu | luu | lx | lu] x W instances, contiguous.
// ------------------------------------------------------------------------------------------
// One thread per (instance, knot): fully parallel over B x (N+1), like the reference's
// thread-pool tasks (ilqr.hpp:354-365) but with the batch as the wide axis.
// kPhased: only instances about to run an inner iteration (kPhInner) are expanded and the
// per-knot costs_ are not stored (the phased engine carries J0 like k_solve does).
#ifndef ALTRO_EXP_MINB
#define ALTRO_EXP_MINB 4
#endif
template <class M, int W, bool kPhased = false>
__global__ void __launch_bounds__(128, ALTRO_EXP_MINB) k_update_expansions(SolverParams P) {
  // all threads of a CTA work on one knot: descriptor words are read from the global blob
  // (uniform addresses, L1-resident); no per-CTA staging of the whole blob
  const char* s_blob = P.blob;
  constexpr int n = M::n, m = M::m;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (b >= P.B) return;
  const Lane<M, W> L(P, s_blob, b / W, b % W);
  if (kPhased && L.is(I_PHASE) != kPhInner) return;
  const int zsel = L.is(I_ZSEL);
  const double penalty = L.sc(S_PENALTY);
  const double* zc = L.z(zsel, k);
  double x[n], u[m];
  ALTRO_UNROLL
  for (int q = 0; q < n; ++q) x[q] = zc[q * W];
  ALTRO_UNROLL
  for (int q = 0; q < m; ++q) u[q] = zc[(n + q) * W];
  const double* lam = P.pmax > 0 ? L.lam(k) : nullptr;
  double* e = L.exp(k);
  constexpr int off_cost = n * n + n * m;  // record: [A | B | lxx | lxu | luu | lx | lu]
  {
    // cost / constraint expansion first, written out before the dynamics Jacobian is formed: the two
    // halves never hold their registers at the same time
    double lxx[n * n], lxu[n * m], luu[m * m], lx[n], lu[m], dA[1], dB[1];
    knot_expansion<M, W, false>(L.D, P.N, k, x, u, lam, penalty, dA, dB, lxx, lxu, luu, lx, lu);
    int f = off_cost;
    ALTRO_UNROLL
    for (int q = 0; q < n * n; ++q) e[(f++) * W] = lxx[q];
    ALTRO_UNROLL
    for (int q = 0; q < n * m; ++q) e[(f++) * W] = lxu[q];
    ALTRO_UNROLL
    for (int q = 0; q < m * m; ++q) e[(f++) * W] = luu[q];
    ALTRO_UNROLL
    for (int q = 0; q < n; ++q) e[(f++) * W] = lx[q];
    ALTRO_UNROLL
    for (int q = 0; q < m; ++q) e[(f++) * W] = lu[q];
  }
  {
    double A[n * n], B[n * m];
    if (k < P.N) {
      rk4_jacobian<M>(L.D.params(), x, u, L.D.h(k), A, B);
    } else {  // IdentityDynamics::Jacobian, problem.hpp:40-43: setIdentity on n x (n+m)
      ALTRO_UNROLL
      for (int q = 0; q < n * n; ++q) A[q] = (q % (n + 1) == 0) ? 1.0 : 0.0;
      ALTRO_UNROLL
      for (int q = 0; q < n * m; ++q) B[q] = 0.0;
    }
    int f = 0;
    ALTRO_UNROLL
    for (int q = 0; q < n * n; ++q) e[(f++) * W] = A[q];
    ALTRO_UNROLL
    for (int q = 0; q < n * m; ++q) e[(f++) * W] = B[q];
  }
  if (!kPhased) {
    // costs_(k) = Cost(x,u), ilqr.hpp:675
    *L.costs(k) = knot_cost<n, m, W>(L.D, k, x, u, lam, AlPen(penalty), nullptr);
  }
  if (k == 0) L.sc(S_CSRC_ALPHA) = -1.0;
}

// Constraint values c(x_k, u_k) of the current trajectory at one knot, rows in ALCost order
// (what ConstraintValues::GetConstraintValue() holds after Cost(), constraint_values.hpp:216-221;
// used by GetConstraintInfo / PrintViolations, al_solver.hpp:68-104).  out = [B][pmax], rows past
// the knot's own count are zero.
template <class M, int W>
__global__ void __launch_bounds__(128) k_constraint_values(SolverParams P, int k, double* out) {
  extern __shared__ __align__(128) char s_blob[];
  copy_blob(P.blob, s_blob, P.blob_bytes);
  constexpr int n = M::n, m = M::m;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const Lane<M, W> L(P, s_blob, b / W, b % W);
  const double* zc = L.z(L.is(I_ZSEL), k);
  double x[n], u[m];
  ALTRO_UNROLL
  for (int q = 0; q < n; ++q) x[q] = zc[q * W];
  ALTRO_UNROLL
  for (int q = 0; q < m; ++q) u[q] = zc[(n + q) * W];
  const ConSet& cs = L.D.conset(k);
  double* o = out + static_cast<size_t>(b) * P.pmax;
  for (int r = 0; r < P.pmax; ++r) o[r] = 0.0;
  for (int bi = 0; bi < cs.nblocks; ++bi) {
    const ConBlock& blk = cs.blk[bi];
    for (int i = 0; i < blk.p; ++i) o[blk.row0 + i] = con_row_fast<n, m>(blk, i, x, u);
  }
}

// --- TMA 1-D bulk copy + mbarrier helpers (cp.async.bulk -> SASS UBLKCP) -------------------
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
// GB/s" is quoted on.  Every lane is one instance: a warp owns 32/W consecutive tiles (lane =
// t*W + i).  The (tile, knot) records (nexp*W*8 B each, contiguous) are streamed N-1 .. 0 through
// a kStages-deep shared-memory ring by TMA bulk copies (one per tile per knot) issued by lane 0
// and tracked with mbarriers; every lane then reads its own column (conflict-free) and runs the
// Riccati step in registers.  Writes K, d (and P, p when kStoreCtg).
// kPhased: only instances in kPhInner run; additionally hands the entry regularisation to the
// line-search kernels (phased.cuh).  The kernel touches nothing but the records, K, d and a few scalars per
// instance: the gain statistic of a failed search (sum_k max_i |d_i|/(|u_i|+1), which needs the
// controls) is formed by the line-search kernel that observes the failure.
#ifndef ALTRO_BP_REFILL_EARLY
#define ALTRO_BP_REFILL_EARLY 0
#endif
template <class M, int W, int kStages, bool kStoreCtg, bool kPhased = false>
__global__ void __launch_bounds__(kWarp) k_backward_mat(SolverParams P) {
  constexpr int n = M::n, m = M::m, nexp = Lane<M, W>::nexp, TPW = kWarp / W;
  constexpr uint32_t kRecBytes = nexp * W * sizeof(double);       // one tile's record
  constexpr int kSlotDoubles = nexp * kWarp;                       // TPW records per ring slot
  extern __shared__ __align__(128) char smem[];
  double* ring = reinterpret_cast<double*>(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(kStages) * kSlotDoubles * sizeof(double));
  const int lane = threadIdx.x;
  const int tile0 = blockIdx.x * TPW;
  const int ntiles = (P.T - tile0 < TPW) ? (P.T - tile0) : TPW;   // tiles of this warp that exist
  const Lane<M, W> L(P, nullptr, tile0 + lane / W, lane % W);    // a == 0 for every lane
  const DevOptions& o = P.opt;
  const int N = P.N;
  const bool valid = L.valid && (!kPhased || L.is(I_PHASE) == kPhInner);
  if (kPhased && !__any_sync(kFull, valid)) return;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const size_t tile_stride = static_cast<size_t>(N + 1) * nexp * W;  // doubles between tiles
  const double* rec0 = P.EXP + static_cast<size_t>(tile0) * tile_stride;
  auto load_slot = [&](int slot, int k) {  // lane 0 only
    mbar_expect_tx(&bars[slot], kRecBytes * ntiles);
    for (int t = 0; t < ntiles; ++t)
      tma_load_1d(ring + static_cast<size_t>(slot) * kSlotDoubles + t * nexp * W,
                  rec0 + t * tile_stride + static_cast<size_t>(k) * nexp * W, kRecBytes, &bars[slot]);
  };

  double reg = 0.0, dreg = 0.0, dV0 = 0.0, dV1 = 0.0;
  int st = kUnsolved;
  if (valid) {
    reg = L.sc(S_REG);
    dreg = L.sc(S_DREG);
    st = L.is(I_STATUS);
  }
  const double reg_in = reg, dreg_in = dreg;
  int max_reg_count = 0;
  bool repeat = valid;
  uint32_t issued = 0, consumed = 0;  // monotonically increasing slot counters (warp-uniform)
  while (__any_sync(kFull, repeat)) {
    // terminal cost-to-go: lxx, lx of knot N (plain loads, once per pass)
    double Pm[n * n], p[n];
    {
      const double* e = L.exp(N);
      constexpr int off_lxx = n * (n + m), off_lx = off_lxx + n * n + n * m + m * m;
      ALTRO_UNROLL
      for (int q = 0; q < n * n; ++q) Pm[q] = valid ? e[(off_lxx + q) * W] : 0.0;
      ALTRO_UNROLL
      for (int q = 0; q < n; ++q) p[q] = valid ? e[(off_lx + q) * W] : 0.0;
      if (kStoreCtg && repeat) {
        double* c = L.ctg(N);
        ALTRO_UNROLL
        for (int q = 0; q < n * n; ++q) c[q * W] = Pm[q];
        ALTRO_UNROLL
        for (int q = 0; q < n; ++q) c[(n * n + q) * W] = p[q];
      }
    }
    // prologue: fill the ring
    int next_k = N - 1;
    for (int s = 0; s < kStages && next_k >= 0; ++s, --next_k, ++issued)
      if (lane == 0) load_slot(issued % kStages, next_k);
    bool live = repeat;  // lanes still descending in this pass
    for (int k = N - 1; k >= 0; --k, ++consumed) {
      const int slot = consumed % kStages;
      mbar_wait(&bars[slot], (consumed / kStages) & 1);
      const double* e = ring + static_cast<size_t>(slot) * kSlotDoubles + (lane / W) * nexp * W + L.i;
      double A[n * n], B[n * m], lxx[n * n], lxu[n * m], luu[m * m], lx[n], lu[m];
      int f = 0;
      ALTRO_UNROLL
      for (int q = 0; q < n * n; ++q) A[q] = e[(f++) * W];
      ALTRO_UNROLL
      for (int q = 0; q < n * m; ++q) B[q] = e[(f++) * W];
      ALTRO_UNROLL
      for (int q = 0; q < n * n; ++q) lxx[q] = e[(f++) * W];
      ALTRO_UNROLL
      for (int q = 0; q < n * m; ++q) lxu[q] = e[(f++) * W];
      ALTRO_UNROLL
      for (int q = 0; q < m * m; ++q) luu[q] = e[(f++) * W];
      ALTRO_UNROLL
      for (int q = 0; q < n; ++q) lx[q] = e[(f++) * W];
      ALTRO_UNROLL
      for (int q = 0; q < m; ++q) lu[q] = e[(f++) * W];
#if ALTRO_BP_REFILL_EARLY
      // (experiment) refill right after the reads, ordered by a generic -> async proxy fence only
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (next_k >= 0) {
        if (lane == 0) load_slot(issued % kStages, next_k);
        --next_k;
        ++issued;
      }
#endif
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
          for (int q = 0; q < m * n; ++q) pk[q * W] = K[q];
          ALTRO_UNROLL
          for (int q = 0; q < m; ++q) pk[(m * n + q) * W] = d[q];
          if (kStoreCtg) {
            double* c = L.ctg(k);
            ALTRO_UNROLL
            for (int q = 0; q < n * n; ++q) c[q * W] = Pm[q];
            ALTRO_UNROLL
            for (int q = 0; q < n; ++q) c[(n * n + q) * W] = p[q];
          }
          if (k == 0) repeat = false;
        }
      }
#if !ALTRO_BP_REFILL_EARLY
      // Refill the slot only now.  The TMA write is an async-proxy access: nothing orders it after the
      // generic-proxy loads above except their COMPLETION, and a load is only known to be complete once its
      // value has been used — the Riccati step has consumed all 47 of them.  Issuing the refill right
      // after the loads (round 1) let a bulk copy that hit L2 overtake loads still queued behind a busy
      // LSU: with other kernels resident on the SM, one tile of a warp occasionally read the last rows
      // of its record (lx, lu — the last loads issued) from the NEXT occupant of the slot.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (next_k >= 0) {
        if (lane == 0) load_slot(issued % kStages, next_k);
        --next_k;
        ++issued;
      }
#endif
    }
  }
  if (valid) {
    decrease_reg(o, reg, dreg);
    L.sc(S_REG) = reg;
    L.sc(S_DREG) = dreg;
    L.sc(S_DV0) = dV0;
    L.sc(S_DV1) = dV1;
    L.is(I_STATUS) = st;
    if (kPhased) {
      L.sc(S_REG_IN) = reg_in;
      L.sc(S_DREG_IN) = dreg_in;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Layout conversion kernels (instance-major <-> tile-major), runtime tile width P.W
// ------------------------------------------------------------------------------------------
struct Unom { double v[kMaxDim]; };

// inputs: x0 [B][n], U0 [B][N][m] or nullptr (+ unom[m] by value).  Z_ := buffer 0 with the
// initial guess; every other buffer is zeroed (Zbar_->SetZero(), ilqr.hpp:233-234).
__global__ void k_pack_inputs(SolverParams P, const double* __restrict__ x0,
                              const double* __restrict__ U0, Unom unom, int nbuf) {
  const int n = P.n, m = P.m, nz = n + m, N = P.N, W = P.W;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (b >= P.Bp) return;
  const int tile = b / W, i = b % W;
  const int src = b < P.B ? b : P.B - 1;  // padding instances mirror the last one (never read back)
  const size_t off = (static_cast<size_t>(tile) * (N + 1) + k) * nz * W + i;
  for (int q = 0; q < n; ++q) {
    P.Z[0][off + q * W] = (k == 0) ? x0[static_cast<size_t>(src) * n + q] : 0.0;
    for (int zb = 1; zb < nbuf; ++zb) P.Z[zb][off + q * W] = 0.0;
  }
  for (int q = 0; q < m; ++q) {
    double u = 0.0;
    if (k < N) u = U0 ? U0[(static_cast<size_t>(src) * N + k) * m + q] : unom.v[q];
    P.Z[0][off + (n + q) * W] = u;
    for (int zb = 1; zb < nbuf; ++zb) P.Z[zb][off + (n + q) * W] = 0.0;
  }
  if (k == 0) {
    double* px0 = P.X0 + static_cast<size_t>(tile) * n * W + i;
    for (int q = 0; q < n; ++q) px0[q * W] = x0[static_cast<size_t>(src) * n + q];
    P.is[static_cast<size_t>(I_ZSEL) * P.Bp + b] = 0;
    P.is[static_cast<size_t>(I_ROLLED) * P.Bp + b] = 0;
    P.sc[static_cast<size_t>(S_CSRC_ALPHA) * P.Bp + b] = -1.0;
    P.sc[static_cast<size_t>(S_CAND_ALPHA) * P.Bp + b] = -1.0;
  }
}

__global__ void k_set_states(SolverParams P, const double* __restrict__ X) {
  const int n = P.n, nz = P.n + P.m, N = P.N, W = P.W;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (b >= P.B) return;
  const int tile = b / W, i = b % W;
  const int sel = P.is[static_cast<size_t>(I_ZSEL) * P.Bp + b];
  double* z = P.Z[sel] + (static_cast<size_t>(tile) * (N + 1) + k) * nz * W + i;
  for (int q = 0; q < n; ++q) z[q * W] = X[(static_cast<size_t>(b) * (N + 1) + k) * n + q];
  if (k == 0) P.is[static_cast<size_t>(I_ROLLED) * P.Bp + b] = 0;
}

// generic gather: dst[b][kk][f] = src[tile][k0+kk][f0+f][i], f < nf.  sel_by_zsel: the source is
// the trajectory buffer P.Z[zsel[b]].
__global__ void k_unpack(SolverParams P, const double* __restrict__ src0, int K, int F, int f0,
                         int nf, int k0, int nk, double* __restrict__ dst, int sel_by_zsel) {
  const int W = P.W;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int kk = blockIdx.y;
  if (b >= P.B || kk >= nk) return;
  const int tile = b / W, i = b % W;
  const double* src = src0;
  if (sel_by_zsel) src = P.Z[P.is[static_cast<size_t>(I_ZSEL) * P.Bp + b]];
  const double* s = src + (static_cast<size_t>(tile) * K + (k0 + kk)) * F * W + i;
  double* d = dst + (static_cast<size_t>(b) * nk + kk) * nf;
  for (int f = 0; f < nf; ++f) d[f] = s[(f0 + f) * W];
}

__global__ void k_fill_duals(SolverParams P, int k, const double* __restrict__ lam, int p) {
  const int W = P.W;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.Bp) return;
  const int tile = b / W, i = b % W;
  double* l = P.LAM + (static_cast<size_t>(tile) * (P.N + 1) + k) * P.pmax * W + i;
  for (int r = 0; r < p; ++r) l[r * W] = lam[r];
}

__global__ void k_fill_scalar(double* p, double v, int count) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < count) p[q] = v;
}
__global__ void k_iota(int* p, int count) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < count) p[q] = q;
}
__global__ void k_fill_int(int* p, int v, int count) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < count) p[q] = v;
}

}  // namespace altro_b200
