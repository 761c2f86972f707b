// outer.cuh — the outer-loop work of the phased engine as a short sequence of dense kernels.
//
// Between two inner iterations an instance may have outer-loop work pending:
//   kPhAlInit      AugmentedLagrangianiLQR::Init()                      (al_solver.hpp:287-302)
//   kPhOuter       UpdateDuals, UpdateConvergenceStatistics, IsDone,
//                  UpdatePenalties                                       (al_solver.hpp:313-355, 368-401)
//   kPhSolveStart  iLQR::Solve() entry: SolveSetup, Rollout, initial Cost (ilqr.hpp:284-299, 453-459)
//   kPhDone        Cost() of the final trajectory, once                  (ilqr.hpp:326-334)
// Only a few instances are in these phases at any time, scattered over the tiles.  They are put on
// a dense list; everything that is independent per knot (dual update, constraint violation, knot
// costs) runs as one thread per (entry, knot), and only the two genuinely serial chains — the
// re-rollout of a rejected candidate for Q8 and the open-loop Rollout — run as one lane per entry:
//
//   k_outer_select    list the instances with pending outer work, count the unfinished ones
//   k_outer_regen     Q8: regenerate the last evaluated candidate of a failed search (lane per entry)
//   k_outer_duals     Init(): duals <- 0; UpdateDuals + max violation   (thread per entry x knot)
//   k_outer_decide    Init() scalars; IsDone / UpdatePenalties; SolveSetup (thread per entry)
//   k_outer_rollout   Rollout()                                          (lane per entry)
//   k_outer_cost      knot costs + violation of the trajectory           (thread per entry x knot)
//   k_outer_finish    ordered cost sums, phase changes, final report     (thread per entry)
//
// Every per-instance quantity is computed by the same device functions, on the same values and in
// the same order as in k_solve, so the results are bit-identical to the fused engine.
#pragma once

#include "kernels.cuh"

namespace altro_b200 {

constexpr int kOuterThreads = 128;

// running maximum of non-negative doubles through their bit patterns (exact, order-independent)
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(__double_as_longlong(v)));
}

// per-instance addressing by batch index (the lanes of a warp hold unrelated instances here)
template <class M, int W>
struct Inst {
  static constexpr int n = M::n, m = M::m, nz = M::n + M::m, nkd = M::m * M::n + M::m;
  const SolverParams& P;
  int b, tile, i;
  __device__ __forceinline__ Inst(const SolverParams& P_, int b_) : P(P_), b(b_), tile(b_ / W), i(b_ % W) {}
  __device__ __forceinline__ double* z(int sel, int k) const {
    return P.Z[sel] + (static_cast<size_t>(tile) * (P.N + 1) + k) * nz * W + i;
  }
  __device__ __forceinline__ double* kd(int k) const {
    return P.KD + (static_cast<size_t>(tile) * P.N + k) * nkd * W + i;
  }
  __device__ __forceinline__ double* lam(int k) const {
    return P.LAM + (static_cast<size_t>(tile) * (P.N + 1) + k) * P.pmax * W + i;
  }
  __device__ __forceinline__ double* x0() const { return P.X0 + static_cast<size_t>(tile) * n * W + i; }
  __device__ __forceinline__ double& sc(int f) const { return P.sc[static_cast<size_t>(f) * P.Bp + b]; }
  __device__ __forceinline__ int& is(int f) const { return P.is[static_cast<size_t>(f) * P.Bp + b]; }
};

// (a) ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_outer_select(SolverParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int ph = kPhReported;
  if (b < P.B) ph = P.is[static_cast<size_t>(I_PHASE) * P.Bp + b];
  const bool pending = ph == kPhAlInit || ph == kPhSolveStart || ph == kPhOuter || ph == kPhDone;
  const bool unfinished = ph < kPhReported;
  const unsigned lane = threadIdx.x % kWarp;
  const unsigned mp = __ballot_sync(kFull, pending), mu = __ballot_sync(kFull, unfinished);
  int base = 0;
  if (lane == 0) {
    if (mp) base = atomicAdd(&P.counters[4], __popc(mp));
    if (mu) atomicAdd(&P.counters[0], __popc(mu));
  }
  base = __shfl_sync(kFull, base, 0);
  if (pending) {
    P.olist[base + __popc(mp & ((1u << lane) - 1u))] = b;
    P.sc[static_cast<size_t>(S_VTMP) * P.Bp + b] = 0.0;
  }
}

// (b) Q8: after a fully failed line search the stored constraint values are those of the last
// evaluated candidate; it is regenerated into candidate buffer (zsel + 1) so that the dual update
// sees the same values (RolloutClosedLoop(alpha), ilqr.hpp:468-499).
// (e) Rollout(), ilqr.hpp:453-459.
// Both are skipped where they would only reproduce bits that are already there: the deep line search
// leaves the last try of a failed search in buffer zsel + 1 (S_CAND_ALPHA says which step length it
// holds); and Z_ of an instance that has been through an iLQR solve was produced by a rollout of
// the same discrete dynamics from the same x0 and controls (I_ROLLED), so rolling it out again returns it
// unchanged (tests: the fused engine, which always rolls out, is bit-identical).  force != 0 runs
// them regardless (ALTRO_B200_ALWAYS_ROLLOUT=1).
template <class M, int W, bool kClosed>
__global__ void __launch_bounds__(kOuterThreads) k_outer_rollout(SolverParams P, int force) {
  constexpr int n = M::n, m = M::m, nz = n + m, nkd = Inst<M, W>::nkd, GZ = kWarp / W;
  const int count = P.counters[4];
  const Desc D(P.blob);  // uniform descriptor words straight from global memory (L1-resident)
  const DevOptions& o = P.opt;
  const int N = P.N;
  const double* mp = D.params();
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    const Inst<M, W> I(P, P.olist[e]);
    const int ph = I.is(I_PHASE);
    const double alpha = kClosed ? I.sc(S_CSRC_ALPHA) : 0.0;
    if (kClosed ? !(ph == kPhOuter && alpha >= 0.0) : ph != kPhSolveStart) continue;
    if (!force && (kClosed ? I.sc(S_CAND_ALPHA) == alpha : I.is(I_ROLLED) != 0)) continue;
    const int zsel = I.is(I_ZSEL);
    const int zout = kClosed ? (zsel + 1) % (GZ + 1) : zsel;
    double x[n], u[m];
    {
      const double* px0 = I.x0();
      ALTRO_UNROLL
      for (int q = 0; q < n; ++q) x[q] = px0[q * W];
    }
    // knot record one knot ahead of the chain: reference state, control and (closed loop) gains
    double rx[n], ru[m], rk[kClosed ? nkd : 1];
    auto fetch = [&](int k) {
      const double* g = I.z(zsel, k);
      if (kClosed) {
        ALTRO_UNROLL
        for (int q = 0; q < n; ++q) rx[q] = g[q * W];
      }
      ALTRO_UNROLL
      for (int q = 0; q < m; ++q) ru[q] = g[(n + q) * W];
      if (kClosed && k < N) {
        const double* kd = I.kd(k);
        ALTRO_UNROLL
        for (int q = 0; q < nkd; ++q) rk[q] = kd[q * W];
      }
    };
    fetch(0);
    bool ok = true;
    for (int k = 0; k <= N && ok; ++k) {
      double cx[n], cu[m], ck[kClosed ? nkd : 1];
      ALTRO_UNROLL
      for (int q = 0; q < n; ++q) cx[q] = kClosed ? rx[q] : 0.0;
      ALTRO_UNROLL
      for (int q = 0; q < m; ++q) cu[q] = ru[q];
      if (kClosed) {
        ALTRO_UNROLL
        for (int q = 0; q < nkd; ++q) ck[q] = rk[q];
      }
      if (k < N) fetch(k + 1);
      double* zn = I.z(zout, k);
      if (k < N) {
        if (kClosed) {
          double dx[n];
          ALTRO_UNROLL
          for (int q = 0; q < n; ++q) dx[q] = x[q] - cx[q];
          ALTRO_UNROLL
          for (int q = 0; q < m; ++q) {
            double acc = ck[q] * dx[0];
            ALTRO_UNROLL
            for (int j = 1; j < n; ++j) acc += ck[q + j * m] * dx[j];
            const double dq = ck[m * n + q];
            u[q] = cu[q] + acc + dq * alpha;  // ilqr.hpp:478
          }
        } else {
          ALTRO_UNROLL
          for (int q = 0; q < m; ++q) u[q] = cu[q];
        }
      } else {  // terminal knot of Zbar: u_N = 0 (SetZero, Q14); Z_ keeps its stored u_N
        ALTRO_UNROLL
        for (int q = 0; q < m; ++q) u[q] = 0.0;
      }
      ALTRO_UNROLL
      for (int q = 0; q < n; ++q) zn[q * W] = x[q];
      if (kClosed) {
        ALTRO_UNROLL
        for (int q = 0; q < m; ++q) zn[(n + q) * W] = u[q];
      }
      if (k < N) {
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
          if (sx > o.state_max_sq || su > o.control_max_sq) ok = false;
        }
      }
    }
  }
}

// (c) AugmentedLagrangianiLQR::Init(): duals <- 0 (al_solver.hpp:292-294);
//     UpdateDuals on the stored constraint values + their max violation
//     (constraint_values.hpp:192-194, 216-221; Q8 decides which trajectory they come from).
template <class M, int W>
__global__ void __launch_bounds__(kOuterThreads) k_outer_duals(SolverParams P) {
  constexpr int n = M::n, m = M::m, GZ = kWarp / W;
  if (P.pmax == 0) return;
  const int count = P.counters[4];
  const int nk = P.N + 1;
  const long items = static_cast<long>(count) * nk;
  const Desc D(P.blob);
  for (long w = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; w < items;
       w += static_cast<long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(w / nk), k = static_cast<int>(w % nk);
    const Inst<M, W> I(P, P.olist[e]);
    const int ph = I.is(I_PHASE);
    double* lam = I.lam(k);
    if (ph == kPhAlInit) {
      if (P.opt.reset_duals)
        for (int r = 0; r < P.pmax; ++r) lam[r * W] = 0.0;
      continue;
    }
    if (ph != kPhOuter) continue;
    const ConSet& cs = D.conset(k);
    if (cs.nblocks == 0) continue;
    const int zsel = I.is(I_ZSEL);
    const double penalty = I.sc(S_PENALTY);
    const int sel = (I.sc(S_CSRC_ALPHA) >= 0.0) ? (zsel + 1) % (GZ + 1) : zsel;
    const double* zc = I.z(sel, k);
    double x[n], u[m];
    ALTRO_UNROLL
    for (int q = 0; q < n; ++q) x[q] = zc[q * W];
    ALTRO_UNROLL
    for (int q = 0; q < m; ++q) u[q] = zc[(n + q) * W];
    double vmax = 0.0;
    for (int bi = 0; bi < cs.nblocks; ++bi) {
      const ConBlock& blk = cs.blk[bi];
      const BlockHdr hd(cs.hdr[bi]);
      const bool eq = hd.eq;
      double* lb = lam + hd.row0 * W;
      for_row_chunks<n, m>(hd, blk, x, u, [&](const double* c, const int* ic, const bool* ok) {
        double l[kChunk];
        ALTRO_UNROLL
        for (int q = 0; q < kChunk; ++q) l[q] = lb[ic[q] * W];
        ALTRO_UNROLL
        for (int q = 0; q < kChunk; ++q) {
          if (ok[q]) {
            const double arg = l[q] - penalty * c[q];
            lb[ic[q] * W] = eq ? arg : neg_part(arg);
            vmax = fmax(vmax, eq ? fabs(c[q]) : fabs(c[q] - neg_part(c[q])));
          }
        }
      });
    }
    if (vmax > 0.0) atomic_max_nonneg(&I.sc(S_VTMP), vmax);
  }
}

// (d) Init() scalars; AL convergence test and penalty update; SolveSetup of the next iLQR solve.
__global__ void __launch_bounds__(kOuterThreads) k_outer_decide(SolverParams P) {
  const int count = P.counters[4];
  const DevOptions& o = P.opt;
  const size_t Bp = P.Bp;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    const int b = P.olist[e];
    auto sc = [&](int f) -> double& { return P.sc[static_cast<size_t>(f) * Bp + b]; };
    auto is = [&](int f) -> int& { return P.is[static_cast<size_t>(f) * Bp + b]; };
    int ph = is(I_PHASE);
    if (ph == kPhAlInit) {  // al_solver.hpp:287-302 (Q10)
      if (o.initial_penalty > 0) sc(S_PENALTY) = o.initial_penalty;
      is(I_ITERS_OUTER) = 0;  // stats.Reset()
      is(I_ITERS_TOTAL) = 0;
      sc(S_COST_CUR) = 0.0;
      sc(S_COST_PREV) = 0.0;
      is(I_STATUS_AL) = kUnsolved;
      ph = kPhSolveStart;
    } else if (ph == kPhOuter) {  // al_solver.hpp:313-333, IsDone :368-401 (Q16)
      const double viol = sc(S_VTMP);
      const double penalty = sc(S_PENALTY);
      const int st = is(I_STATUS);
      const int it_outer = is(I_ITERS_OUTER) + 1;
      sc(S_VIOL) = viol;
      is(I_ITERS_OUTER) = it_outer;
      const double max_penalty = P.pmax > 0 ? penalty : 0.0;
      ph = kPhDone;
      if (st != kSolved) {
        is(I_STATUS_AL) = st;
      } else if (viol < o.constraint_tolerance) {
        is(I_STATUS_AL) = kSolved;
      } else if (max_penalty > o.maximum_penalty) {
        is(I_STATUS_AL) = kMaxPenalty;
      } else if (it_outer >= o.max_iterations_outer) {
        is(I_STATUS_AL) = kMaxOuterIterations;
      } else if (is(I_ITERS_TOTAL) >= o.max_iterations_total) {
        is(I_STATUS_AL) = kMaxIterations;
      } else {
        sc(S_PENALTY) = penalty * o.penalty_scaling;  // UpdatePenalties
        ph = kPhSolveStart;
      }
    }
    if (ph == kPhSolveStart) {  // SolveSetup :629-645, ResetInternalVariables :680-690
      is(I_ITERS_INNER) = 0;
      is(I_LSFAIL) = 0;
      is(I_STATUS) = kUnsolved;
      sc(S_REG) = o.bp_reg_initial;
      sc(S_DREG) = 0.0;
      sc(S_DV0) = 0.0;
      sc(S_DV1) = 0.0;
    }
    is(I_PHASE) = ph;
    sc(S_VTMP) = 0.0;  // reused by k_outer_cost
  }
}

// (f) ALCost::Evaluate of every knot of the trajectories that were just rolled out (initial cost
// of an iLQR solve) or that terminated (final Cost()): COSTK[entry][k]; max violation -> S_VTMP.
template <class M, int W>
__global__ void __launch_bounds__(kOuterThreads) k_outer_cost(SolverParams P) {
  constexpr int n = M::n, m = M::m;
  const int count = P.counters[4];
  const int nk = P.N + 1;
  const long items = static_cast<long>(count) * nk;
  const Desc D(P.blob);
  for (long w = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; w < items;
       w += static_cast<long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(w / nk), k = static_cast<int>(w % nk);
    const Inst<M, W> I(P, P.olist[e]);
    const int ph = I.is(I_PHASE);
    if (ph != kPhSolveStart && ph != kPhDone) continue;
    const double* zc = I.z(I.is(I_ZSEL), k);
    double x[n], u[m];
    ALTRO_UNROLL
    for (int q = 0; q < n; ++q) x[q] = zc[q * W];
    ALTRO_UNROLL
    for (int q = 0; q < m; ++q) u[q] = zc[(n + q) * W];
    const AlPen pen(I.sc(S_PENALTY));
    const double* lam = P.pmax > 0 ? I.lam(k) : nullptr;
    double v = 0.0;
    P.COSTK[w] = knot_cost<n, m, W>(D, k, x, u, lam, pen, &v);
    if (v > 0.0) atomic_max_nonneg(&I.sc(S_VTMP), v);
  }
}

// (g) cost sums in knot order, phase changes, report of terminated instances.
__global__ void __launch_bounds__(kOuterThreads) k_outer_finish(SolverParams P, int mode) {
  const int count = P.counters[4];
  const DevOptions& o = P.opt;
  const size_t Bp = P.Bp;
  const int nk = P.N + 1;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    const int b = P.olist[e];
    auto sc = [&](int f) -> double& { return P.sc[static_cast<size_t>(f) * Bp + b]; };
    auto is = [&](int f) -> int& { return P.is[static_cast<size_t>(f) * Bp + b]; };
    const int ph = is(I_PHASE);
    if (ph != kPhSolveStart && ph != kPhDone) continue;
    const double* ck = P.COSTK + static_cast<size_t>(e) * nk;
    double J = 0.0;
    for (int k = 0; k < nk; ++k) J += ck[k];
    if (ph == kPhSolveStart) {
      sc(S_J0) = J;
      sc(S_INITIAL_COST) = J;
      sc(S_CSRC_ALPHA) = -1.0;
      is(I_ROLLED) = 1;
      is(I_PHASE) = (o.max_iterations_inner > 0) ? kPhInner : (mode == 1 ? kPhOuter : kPhDone);
    } else {  // Cost() of the final trajectory under the final duals / penalty
      if (mode == 0) sc(S_VIOL) = sc(S_VTMP);  // == GetMaxViolation() after Cost()
      sc(S_COST) = J;
      sc(S_CSRC_ALPHA) = -1.0;
      is(I_PHASE) = kPhReported;
      atomicSub(&P.counters[0], 1);
    }
  }
}

}  // namespace altro_b200
