// phased.cuh — the phased engine: one inner iteration of the whole batch as a short sequence
// of throughput kernels, each with the thread mapping that suits its phase, instead of one
// persistent warp-per-tile state machine (k_solve).
//
//   k_solve(parts = outer|start)   AL outer step, SolveSetup + Rollout of the next iLQR solve,
//                                  final Cost() of terminated instances      (kernels.cuh)
//   k_update_expansions<phased>    UpdateExpansions: one thread per (instance, knot), writes the
//                                  47-double records the reference hands to BackwardPass
//                                  (ilqr.hpp:350-366, 670-677)
//   k_backward_mat<phased>         BackwardPass: one lane per instance, records streamed through
//                                  shared memory by TMA bulk copies            (ilqr.hpp:385-445)
//   k_ls_wide                      ForwardPass, first G = 32/W step lengths of every instance at
//                                  once (lane = a*W + i)                        (ilqr.hpp:512-558)
//   k_ls_deep                      ForwardPass, the remaining step lengths, one instance per warp
//                                  (32 step lengths per round), only for the instances on `list`:
//                                  those whose first G tries failed and those whose previous
//                                  search failed completely
//
// An instance's arithmetic is the same device code in the same order whichever kernel runs it,
// so results do not depend on the engine, the tile width or the list order.  The per-instance
// state lives in the SoA scalar arrays between kernels; the tail of an inner iteration
// (finish_inner) runs in whichever line-search kernel concludes the instance's search.
#pragma once

#include "kernels.cuh"

namespace altro_b200 {

constexpr int kLsWarps = 4;  // warps per CTA of the line-search kernels

// Lane view for the line-search kernels: the warp holds WI instances x G = 32/WI step lengths.
// (tile, i) locate the instance in the tile-major arrays of width W; si is its column in the
// warp's private staging buffer (width WI); a is the lane's step-length group.
template <class M, int W, int WI>
struct PLane {
  static constexpr int n = M::n, m = M::m, nz = M::n + M::m, nkd = M::m * M::n + M::m;
  static constexpr int G = kWarp / WI;
  const SolverParams& P;
  Desc D;
  int b, tile, i, a, si;
  bool valid;
  __device__ __forceinline__ PLane(const SolverParams& P_, const char* blob, int b_, int a_, int si_, bool valid_)
      : P(P_), D(blob), b(b_), tile(b_ / W), i(b_ % W), a(a_), si(si_), valid(valid_) {}
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

template <class M>
__host__ __device__ constexpr int p_stage_doubles(int pmax, int WI) {
  return 2 * ((M::n + M::m) + (M::m * M::n + M::m) + pmax) * WI;
}

// RolloutClosedLoop(alpha) + Cost(Zbar) + normalised feed-forward gain of the candidate
// (ilqr.hpp:468-499, 326-334, 662-668).  The candidate is written to zo[k*zknot + f*zrow].
// Returns false when the state/control bound check trips (status is set like the reference).
// kDbg (latency probe only, 0 in every solve kernel): 1 no cost, 2 no dynamics, 4 no staging
// after knot 0, 8 no candidate stores, 16 no normalised-gain division
// store = false: the candidate is evaluated but not written anywhere (the deep search only needs
// its cost; the winner is rolled out again to its destination).  kGain = false: no normalised
// feed-forward gain (gsum = 0).
template <class M, int W, int WI, int kDbg = 0, bool kGain = true>
__device__ __forceinline__ bool p_rollout(const PLane<M, W, WI>& L, double* stg, bool active, int zsel,
                                          double* zo, int zrow, int zknot, double alpha, double penalty,
                                          double& J, double& gsum, int& status, bool store = true) {
  constexpr int n = M::n, m = M::m, nz = n + m, nkd = PLane<M, W, WI>::nkd;
  const int N = L.P.N, pmax = L.P.pmax;
  const DevOptions& o = L.P.opt;
  const Desc& D = L.D;
  const double* mp = D.params();
  const int off_lam = nz + nkd;
  const int R = off_lam + pmax;
  const AlPen pen(penalty);
  auto issue = [&](int k) {
    double* s = stg + (k & 1) * R * WI;
    if (WI == 1) {  // one instance per warp: lane r stages row r
      for (int r = L.a; r < R; r += kWarp) {
        const double* g;
        if (r < nz) {
          g = L.z(zsel, k) + r * W;
        } else if (r < off_lam) {
          if (k >= N) continue;
          g = L.kd(k) + (r - nz) * W;
        } else {
          g = L.lam(k) + (r - off_lam) * W;
        }
        cp_async8(s + r, g);
      }
    } else if (WI <= 4) {  // few instances per warp: the G lanes of an instance share its rows
      for (int r = L.a; r < R; r += PLane<M, W, WI>::G) {
        const double* g;
        if (r < nz) {
          g = L.z(zsel, k) + r * W;
        } else if (r < off_lam) {
          if (k >= N) continue;
          g = L.kd(k) + (r - nz) * W;
        } else {
          g = L.lam(k) + (r - off_lam) * W;
        }
        cp_async8(s + r * WI + L.si, g);
      }
    } else if (L.a == 0) {
      s += L.si;
      const double* g = L.z(zsel, k);
      ALTRO_UNROLL
      for (int r = 0; r < nz; ++r) cp_async8(s + r * WI, g + r * W);
      if (k < N) {
        g = L.kd(k);
        ALTRO_UNROLL
        for (int r = 0; r < nkd; ++r) cp_async8(s + (nz + r) * WI, g + r * W);
      }
      if (pmax > 0) {
        g = L.lam(k);
        for (int r = 0; r < pmax; ++r) cp_async8(s + (off_lam + r) * WI, g + r * W);
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
  double Jsum = 0.0, gs = 0.0;
  bool ok = active;
  for (int k = 0; k <= N; ++k) {
    cp_async_wait_all();
    __syncwarp();
    if (k < N && !(kDbg & 4)) issue(k + 1);
    const double* s = stg + ((kDbg & 4) ? 0 : (k & 1)) * R * WI + L.si;
    if (ok) {
      double* zn = zo + static_cast<size_t>(k) * zknot;
      double g = 0.0;
      if (k < N) {
        double dx[n];
        ALTRO_UNROLL
        for (int q = 0; q < n; ++q) dx[q] = x[q] - s[q * WI];
        ALTRO_UNROLL
        for (int q = 0; q < m; ++q) {
          double acc = s[(nz + q) * WI] * dx[0];
          ALTRO_UNROLL
          for (int j = 1; j < n; ++j) acc += s[(nz + q + j * m) * WI] * dx[j];
          const double dq = s[(nz + m * n + q) * WI];
          u[q] = s[(n + q) * WI] + acc + dq * alpha;  // ilqr.hpp:478
          if (kGain) {
            const double gq = (kDbg & 16) ? fabs(dq) : fabs(dq) / (fabs(u[q]) + 1);
            g = (q == 0) ? gq : fmax(g, gq);
          }
        }
      } else {  // terminal knot of Zbar: u_N = 0 (SetZero, Q14)
        ALTRO_UNROLL
        for (int q = 0; q < m; ++q) u[q] = 0.0;
      }
      if (store && !(kDbg & 8)) {
        ALTRO_UNROLL
        for (int q = 0; q < n; ++q) zn[q * zrow] = x[q];
        ALTRO_UNROLL
        for (int q = 0; q < m; ++q) zn[(n + q) * zrow] = u[q];
      }
      double v;
      if (!(kDbg & 1)) Jsum += knot_cost<n, m, WI>(D, k, x, u, s + off_lam * WI, pen, &v);
      else Jsum += u[0];
      if (k < N) {
        gs += g;
        double xn[n];
        if (!(kDbg & 2)) {
          rk4_step<M>(mp, x, u, D.h(k), xn);
        } else {
          ALTRO_UNROLL
          for (int q = 0; q < n; ++q) xn[q] = x[q] + 1e-3 * u[q % m];
        }
        ALTRO_UNROLL
        for (int q = 0; q < n; ++q) x[q] = xn[q];
        if (o.check_forwardpass_bounds) {  // ilqr.hpp:484-495
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
  if (ok) status = kUnsolved;  // ilqr.hpp:497
  J = Jsum;
  gsum = gs;
  return ok;
}

// ForwardPass line search (ilqr.hpp:512-558) over the tries done0, done0+1, ... in rounds of G:
// group a evaluates try done + a.  The lowest accepted try wins, which is where the sequential
// search stops.  Stops after max_rounds rounds even if tries remain (`exhausted` says whether
// every try allowed by line_search_max_iterations has been evaluated).
struct PLsResult {
  bool success, exhausted;
  int slot;  // winning group
  double J, alpha, z, gsum;
};

// kStore = false (deep search): candidates are not written, except the LAST try allowed by
// line_search_max_iterations, which goes to zo: if the whole search fails that is the candidate whose
// constraint values the reference keeps (Q8), and the outer step finds it there.
template <class M, int W, int WI, bool kStore = true>
__device__ __forceinline__ PLsResult p_line_search(const PLane<M, W, WI>& L, double* stg, bool run, int zsel,
                                                   double* zo, int zrow, int zknot, double penalty,
                                                   double J0, double dV0, double dV1, int& status,
                                                   double& csrc, int done0, int max_rounds) {
  constexpr int G = PLane<M, W, WI>::G;
  const DevOptions& o = L.P.opt;
  PLsResult r;
  r.success = false;
  r.slot = 0;
  r.J = J0;
  r.alpha = 1.0;
  r.z = -1.0;
  r.gsum = 0.0;
  double alpha_base = 1.0;  // step length of try index `done`: repeated division like `alpha /= factor`
  for (int j = 0; j < done0; ++j) alpha_base /= o.line_search_decrease_factor;
  int done = done0;
  bool searching = run && (done < o.line_search_max_iterations);
  r.exhausted = run && !searching;
  for (int round = 0; round < max_rounds && __any_sync(kFull, searching); ++round) {
    double alpha = alpha_base;
    for (int j = 0; j < L.a; ++j) alpha /= o.line_search_decrease_factor;
    const bool mine = searching && (done + L.a < o.line_search_max_iterations);
    double J = 0.0, gs = 0.0, z = -1.0;
    int st_try = status;
    const bool ok = p_rollout<M, W, WI, 0, kStore>(L, stg, mine, zsel, zo, zrow, zknot, alpha, penalty, J, gs, st_try,
                                                   kStore || done + L.a == o.line_search_max_iterations - 1);
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
      const unsigned bit = 1u << (L.si + WI * g);
      if (accm & bit) win = g;
      if ((minem & bit) && last < 0) last = g;
      if ((okm & bit) && lastok < 0) lastok = g;
    }
    const int src_win = L.si + WI * (win < 0 ? 0 : win);
    const double Jw = __shfl_sync(kFull, J, src_win);
    const double aw = __shfl_sync(kFull, alpha, src_win);
    const double zw = __shfl_sync(kFull, z, src_win);
    const double gw = __shfl_sync(kFull, gs, src_win);
    const int st_last = __shfl_sync(kFull, st_try, L.si + WI * (last < 0 ? 0 : last));
    const double a_lastok = __shfl_sync(kFull, alpha, L.si + WI * (lastok < 0 ? 0 : lastok));
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
        if (last >= 0) status = st_last;   // status_ left by the last executed rollout
        if (lastok >= 0) csrc = a_lastok;  // Cost(*Zbar_) refreshed the stored constraint values (Q8)
        done += G;
        ALTRO_UNROLL
        for (int j = 0; j < G; ++j) alpha_base /= o.line_search_decrease_factor;
        if (done >= o.line_search_max_iterations) {
          searching = false;
          r.exhausted = true;
        }
      }
    }
  }
  return r;
}

// Tail of the inner iteration for one instance (called by the lane that owns it): loads the rest
// of the instance state, applies finish_inner and writes everything back.
// gsum_bwd: sum_k max_i |d_i|/(|u_i|+1) over the current Z_ (only read when the search failed).
template <class LaneT>
__device__ __forceinline__ void p_finish(const LaneT& L, int mode, const PLsResult& r, int new_zsel, int zsel,
                                         double J0, double csrc, int st, double gsum_bwd, double cand_alpha = -1.0) {
  const DevOptions& o = L.P.opt;
  double cost_cur = L.sc(S_COST_CUR), cost_prev = L.sc(S_COST_PREV);
  const double initial_cost = L.sc(S_INITIAL_COST);
  double alpha_stat = L.sc(S_ALPHA), z_stat = L.sc(S_ZRATIO), grad = L.sc(S_GRAD), dJ = L.sc(S_DJ);
  double reg = L.sc(S_REG), dreg = L.sc(S_DREG);
  int it_inner = L.is(I_ITERS_INNER), it_total = L.is(I_ITERS_TOTAL);
  int phase = kPhInner, lsfail = 0;
  InnerTail t;
  t.success = r.success;
  t.new_zsel = new_zsel;
  t.J = r.J;
  t.alpha = r.alpha;
  t.z = r.z;
  t.gsum_ls = r.gsum;
  t.gsum_bwd = gsum_bwd;
  t.reg_in = L.sc(S_REG_IN);
  t.dreg_in = L.sc(S_DREG_IN);
  finish_inner(o, L.P.N, mode, t, zsel, J0, cost_cur, cost_prev, initial_cost, alpha_stat, z_stat, csrc, grad,
               dJ, reg, dreg, it_inner, it_total, st, phase, lsfail);
  L.is(I_ZSEL) = zsel;
  L.sc(S_J0) = J0;
  L.sc(S_COST_CUR) = cost_cur;
  L.sc(S_COST_PREV) = cost_prev;
  L.sc(S_ALPHA) = alpha_stat;
  L.sc(S_ZRATIO) = z_stat;
  L.sc(S_CSRC_ALPHA) = csrc;
  L.sc(S_CAND_ALPHA) = r.success ? -1.0 : cand_alpha;
  L.sc(S_GRAD) = grad;
  L.sc(S_DJ) = dJ;
  L.sc(S_REG) = reg;
  L.sc(S_DREG) = dreg;
  L.is(I_ITERS_INNER) = it_inner;
  L.is(I_ITERS_TOTAL) = it_total;
  L.is(I_STATUS) = st;
  L.is(I_PHASE) = phase;
  L.is(I_LSFAIL) = lsfail;
}

// ------------------------------------------------------------------------------------------
// Split line search (default): the serial part of a candidate — RolloutClosedLoop, a chain of
// N feedback + RK4 steps — runs alone in k_roll_*; Cost(Zbar) is independent per knot and runs
// as one thread per (candidate, knot) in k_cost_*; k_acc_* sums the knot costs in knot order
// (the same sum the fused kernels form), applies the acceptance test and, for the winner only,
// computes the normalised feed-forward gain.  Same arithmetic per candidate as k_ls_wide /
// k_ls_deep, but the serial chain is ~2.3x shorter.
// ------------------------------------------------------------------------------------------
constexpr int kRollStages = 4;  // knots in flight in the staging ring of the rollout kernels

template <class M>
__host__ __device__ constexpr int p_roll_stage_doubles(int WI) {
  return kRollStages * ((M::n + M::m) + (M::m * M::n + M::m)) * WI;
}

template <int NKeep>
__device__ __forceinline__ void cp_async_wait_keep() {
  asm volatile("cp.async.wait_group %0;" ::"n"(NKeep) : "memory");
}

// RolloutClosedLoop(alpha) only (ilqr.hpp:468-499): candidate to zo[k*zknot + f*zrow].
template <class M, int W, int WI>
__device__ __forceinline__ bool p_rollout_only(const PLane<M, W, WI>& L, double* stg, bool active, int zsel,
                                               double* zo, int zrow, int zknot, double alpha, int& status) {
  constexpr int n = M::n, m = M::m, nz = n + m, nkd = PLane<M, W, WI>::nkd, R = nz + nkd, NS = kRollStages;
  const int N = L.P.N;
  const DevOptions& o = L.P.opt;
  const Desc& D = L.D;
  const double* mp = D.params();
  auto issue = [&](int k) {
    if (k <= N) {
      double* s = stg + (k % NS) * R * WI;
      if (WI == 1) {  // one instance per warp: lane r stages row r
        const int r = L.a;
        if (r < nz) cp_async8(s + r, L.z(zsel, k) + r * W);
        else if (r < R && k < N) cp_async8(s + r, L.kd(k) + (r - nz) * W);
      } else if (L.a == 0) {
        s += L.si;
        const double* g = L.z(zsel, k);
        ALTRO_UNROLL
        for (int r = 0; r < nz; ++r) cp_async8(s + r * WI, g + r * W);
        if (k < N) {
          g = L.kd(k);
          ALTRO_UNROLL
          for (int r = 0; r < nkd; ++r) cp_async8(s + (nz + r) * WI, g + r * W);
        }
      }
    }
    cp_async_commit();  // one group per knot, empty past the end, so the wait count stays uniform
  };
  static_assert(WI != 1 || R <= kWarp, "row-per-lane staging needs R <= 32");
  __syncwarp();
  ALTRO_UNROLL
  for (int k = 0; k < NS - 1; ++k) issue(k);
  double x[n], u[m];
  {
    const double* px0 = L.x0();
    ALTRO_UNROLL
    for (int q = 0; q < n; ++q) x[q] = px0[q * W];
  }
  bool ok = active;
  for (int k = 0; k <= N; ++k) {
    cp_async_wait_keep<NS - 2>();  // knot k has landed
    __syncwarp();
    issue(k + NS - 1);
    const double* s = stg + (k % NS) * R * WI + L.si;
    if (ok) {
      double* zn = zo + static_cast<size_t>(k) * zknot;
      if (k < N) {
        double dx[n];
        ALTRO_UNROLL
        for (int q = 0; q < n; ++q) dx[q] = x[q] - s[q * WI];
        ALTRO_UNROLL
        for (int q = 0; q < m; ++q) {
          double acc = s[(nz + q) * WI] * dx[0];
          ALTRO_UNROLL
          for (int j = 1; j < n; ++j) acc += s[(nz + q + j * m) * WI] * dx[j];
          const double dq = s[(nz + m * n + q) * WI];
          u[q] = s[(n + q) * WI] + acc + dq * alpha;  // ilqr.hpp:478
        }
      } else {  // terminal knot of Zbar: u_N = 0 (SetZero, Q14)
        ALTRO_UNROLL
        for (int q = 0; q < m; ++q) u[q] = 0.0;
      }
      ALTRO_UNROLL
      for (int q = 0; q < n; ++q) zn[q * zrow] = x[q];
      ALTRO_UNROLL
      for (int q = 0; q < m; ++q) zn[(n + q) * zrow] = u[q];
      if (k < N) {
        double xn[n];
        rk4_step<M>(mp, x, u, D.h(k), xn);
        ALTRO_UNROLL
        for (int q = 0; q < n; ++q) x[q] = xn[q];
        if (o.check_forwardpass_bounds) {  // ilqr.hpp:484-495
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
  cp_async_wait_all();
  if (ok) status = kUnsolved;  // ilqr.hpp:497
  return ok;
}

// step length of try t: 1 / factor^t by repeated division, like `alpha /= factor` (ilqr.hpp:553)
__device__ __forceinline__ double try_alpha(const DevOptions& o, int t) {
  double alpha = 1.0;
  for (int j = 0; j < t; ++j) alpha /= o.line_search_decrease_factor;
  return alpha;
}

// max_i |d_i| / (|u_i| + 1) of one knot (NormalizedFeedforwardGain, ilqr.hpp:662-668)
template <int m>
__device__ __forceinline__ double knot_gain(const double* d, int ds, const double* u, int us) {
  double g = 0.0;
  ALTRO_UNROLL
  for (int q = 0; q < m; ++q) {
    const double gq = fabs(d[q * ds]) / (fabs(u[q * us]) + 1);
    g = (q == 0) ? gq : fmax(g, gq);
  }
  return g;
}

// sum_k max_i |d_i|/(|u_i|+1) over the CURRENT trajectory (Z_ in buffer zsel) with the gains of the
// last backward pass, summed in knot order: the gradient statistic of an inner iteration whose line
// search failed, i.e. with Z_ unchanged (ilqr.hpp:571-575).  One lane, serial.
template <class LaneT>
__device__ __forceinline__ double gain_sum_serial(const LaneT& L, int zsel) {
  constexpr int n = LaneT::n, m = LaneT::m;
  const int N = L.P.N, W = L.P.W;
  double gs = 0.0;
  for (int k = 0; k < N; ++k) gs += knot_gain<m>(L.kd(k) + m * n * W, W, L.z(zsel, k) + n * W, W);
  return gs;
}
// The same sum formed by the G lanes that share an instance (the warp holds WI = 32/G instances):
// they take the knots round-robin into the instance's gbuf[N] (shared), lane a = 0 adds them up
// in knot order (same additions as the serial version).  Warp-collective; `want` per instance.
template <class LaneT>
__device__ __forceinline__ double gain_sum_warp(const LaneT& L, bool want, int zsel, double* gbuf) {
  constexpr int n = LaneT::n, m = LaneT::m, G = LaneT::G;
  const int N = L.P.N, W = L.P.W;
  if (want)
    for (int k = L.a; k < N; k += G) gbuf[k] = knot_gain<m>(L.kd(k) + m * n * W, W, L.z(zsel, k) + n * W, W);
  __syncwarp();
  double gs = 0.0;
  if (want && L.a == 0)
    for (int k = 0; k < N; ++k) gs += gbuf[k];
  __syncwarp();
  return gs;
}

// Acceptance over the G candidates of every instance of the warp (tries done0 + a), given each
// candidate's cost J and rollout outcome; same rules and tie-breaking as p_line_search.
template <int WI>
__device__ __forceinline__ PLsResult p_accept(const DevOptions& o, int si, int a, bool run, bool mine, bool ok,
                                              double J, double alpha, int st_try, double J0, double dV0,
                                              double dV1, int done0, int& status, double& csrc) {
  constexpr int G = kWarp / WI;
  PLsResult r;
  r.success = false;
  r.exhausted = false;
  r.slot = 0;
  r.J = J0;
  r.alpha = 1.0;
  r.z = -1.0;
  r.gsum = 0.0;
  double z = -1.0;
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
    const unsigned bit = 1u << (si + WI * g);
    if (accm & bit) win = g;
    if ((minem & bit) && last < 0) last = g;
    if ((okm & bit) && lastok < 0) lastok = g;
  }
  const int src_win = si + WI * (win < 0 ? 0 : win);
  const double Jw = __shfl_sync(kFull, J, src_win);
  const double aw = __shfl_sync(kFull, alpha, src_win);
  const double zw = __shfl_sync(kFull, z, src_win);
  const int st_last = __shfl_sync(kFull, st_try, si + WI * (last < 0 ? 0 : last));
  const double a_lastok = __shfl_sync(kFull, alpha, si + WI * (lastok < 0 ? 0 : lastok));
  if (run) {
    if (win >= 0) {
      r.success = true;
      r.slot = win;
      r.J = Jw;
      r.alpha = aw;
      r.z = zw;
      status = kUnsolved;  // the accepted rollout ran to the end (ilqr.hpp:497)
    } else {
      if (last >= 0) status = st_last;   // status_ left by the last executed rollout
      if (lastok >= 0) csrc = a_lastok;  // Cost(*Zbar_) refreshed the stored constraint values (Q8)
      r.exhausted = done0 + G >= o.line_search_max_iterations;
    }
  }
  (void)a;
  return r;
}

// --- wide: tries 0 .. G-1 of every instance whose previous search did not fail completely -----
template <class M, int W>
__global__ void __launch_bounds__(kLsWarps* kWarp) k_roll_wide(SolverParams P) {
  extern __shared__ __align__(128) char smem[];
  copy_blob(P.blob, smem, P.blob_bytes);
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int tile = blockIdx.x * kLsWarps + warp;
  if (tile >= P.T) return;
  using LaneT = PLane<M, W, W>;
  constexpr int G = LaneT::G, nz = LaneT::nz;
  double* stg = reinterpret_cast<double*>(smem + ((P.blob_bytes + 15) / 16) * 16) +
                static_cast<size_t>(warp) * p_roll_stage_doubles<M>(W);
  const int b = tile * W + lane % W;
  const LaneT L(P, smem, b, lane / W, lane % W, b < P.B);
  const bool run = L.valid && L.is(I_PHASE) == kPhInner;
  if (!__any_sync(kFull, run)) return;
  const int zsel = run ? L.is(I_ZSEL) : 0;
  int st_try = run ? L.is(I_STATUS) : kUnsolved;
  const bool mine = run && L.a < P.opt.line_search_max_iterations;
  const double alpha = try_alpha(P.opt, L.a);
  double* zo = L.z((zsel + 1 + L.a) % (G + 1), 0);
  p_rollout_only<M, W, W>(L, stg, mine, zsel, zo, W, nz * W, alpha, st_try);
  if (mine) P.TRYST[tile * kWarp + lane] = st_try;
}

// ALCost::Evaluate (al_cost.hpp:264-274) of every knot of every candidate: a work item is one
// (tile, knot) row of 32 lanes (lane = a*W + i, like the rollout and acceptance kernels); the
// persistent CTAs walk the items with a grid stride, the problem blob is staged once per CTA.
// COSTK[(tile*(N+1) + k)*32 + lane].
template <class M, int W>
__global__ void __launch_bounds__(kLsWarps* kWarp) k_cost_wide(SolverParams P) {
  extern __shared__ __align__(128) char s_blob[];
  copy_blob(P.blob, s_blob, P.blob_bytes);
  constexpr int n = M::n, m = M::m, G = kWarp / W;
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int nk = P.N + 1;
  const long items = static_cast<long>(P.T) * nk;
  const int i = lane % W, a = lane / W;
  for (long w = static_cast<long>(blockIdx.x) * kLsWarps + warp; w < items; w += static_cast<long>(gridDim.x) * kLsWarps) {
    const int tile = static_cast<int>(w / nk), k = static_cast<int>(w % nk);
    const int b = tile * W + i;
    if (b >= P.B) continue;
    const PLane<M, W, W> L(P, s_blob, b, a, i, true);
    if (L.is(I_PHASE) != kPhInner || a >= P.opt.line_search_max_iterations) continue;
    const int zsel = L.is(I_ZSEL);
    const double* zc = L.z((zsel + 1 + a) % (G + 1), k);
    double x[n], u[m];
    ALTRO_UNROLL
    for (int q = 0; q < n; ++q) x[q] = zc[q * W];
    ALTRO_UNROLL
    for (int q = 0; q < m; ++q) u[q] = zc[(n + q) * W];
    const AlPen pen(L.sc(S_PENALTY));
    const double* lam = P.pmax > 0 ? L.lam(k) : nullptr;
    P.COSTK[static_cast<size_t>(w) * kWarp + lane] = knot_cost<n, m, W>(L.D, k, x, u, lam, pen, nullptr);
  }
}

template <class M, int W>
__global__ void __launch_bounds__(kLsWarps* kWarp) k_acc_wide(SolverParams P, int mode) {
  extern __shared__ __align__(128) char smem[];  // per warp: W x N doubles for the winner's gains
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int tile = blockIdx.x * kLsWarps + warp;
  if (tile >= P.T) return;
  using LaneT = PLane<M, W, W>;
  constexpr int G = LaneT::G, n = M::n, m = M::m;
  const int N = P.N;
  double* gbuf = reinterpret_cast<double*>(smem) + static_cast<size_t>(warp) * W * N;
  const int b = tile * W + lane % W;
  const LaneT L(P, nullptr, b, lane / W, lane % W, b < P.B);
  const bool run = L.valid && L.is(I_PHASE) == kPhInner;
  if (!__any_sync(kFull, run)) return;
  int zsel = 0, st = kUnsolved, st_try = kUnsolved;
  double J0 = 0.0, dV0 = 0.0, dV1 = 0.0, csrc = -1.0;
  const bool mine = run && L.a < P.opt.line_search_max_iterations;
  if (run) {
    zsel = L.is(I_ZSEL);
    st = L.is(I_STATUS);
    J0 = L.sc(S_J0);
    dV0 = L.sc(S_DV0);
    dV1 = L.sc(S_DV1);
    csrc = L.sc(S_CSRC_ALPHA);
  }
  double J = 0.0;
  if (mine) {
    st_try = P.TRYST[tile * kWarp + lane];
    const double* ck = P.COSTK + static_cast<size_t>(tile) * (N + 1) * kWarp + lane;
    for (int k = 0; k <= N; ++k) J += ck[k * kWarp];  // Cost(): sum in knot order
  }
  const double alpha = try_alpha(P.opt, L.a);
  PLsResult r = p_accept<W>(P.opt, L.si, L.a, run, mine, mine && st_try == kUnsolved, J, alpha, st_try, J0, dV0,
                            dV1, 0, st, csrc);
  // normalised feed-forward gain of the accepted candidate: its G lanes take the knots round-robin,
  // the owner sums them in knot order
  const bool succ = run && r.success;
  if (__any_sync(kFull, succ)) {
    if (succ) {
      const int sel = (zsel + 1 + r.slot) % (G + 1);
      for (int k = L.a; k < N; k += G) gbuf[L.si * N + k] = knot_gain<m>(L.kd(k) + m * n * W, W, L.z(sel, k) + n * W, W);
    }
    __syncwarp();
    if (succ && L.a == 0) {
      double gs = 0.0;
      for (int k = 0; k < N; ++k) gs += gbuf[L.si * N + k];
      r.gsum = gs;
    }
  }
  if (run && L.a == 0) {
    if (r.success || r.exhausted) {
      p_finish(L, mode, r, (zsel + 1 + r.slot) % (G + 1), zsel, J0, csrc, st,
               r.success ? 0.0 : gain_sum_serial(L, zsel));
    } else {  // continue with try G in the deep kernels
      L.sc(S_CSRC_ALPHA) = csrc;
      L.is(I_STATUS) = st;
      P.list[atomicAdd(&P.counters[3], 1)] = (b << 1) | 1;
    }
  }
}

// --- deep: the next 32 tries of the instances on P.list, one instance per warp ------------------
template <class M, int W>
__global__ void __launch_bounds__(kLsWarps* kWarp) k_roll_deep(SolverParams P, int wide_tries) {
  extern __shared__ __align__(128) char smem[];
  const int count = P.counters[3];
  if (static_cast<int>(blockIdx.x) * kLsWarps >= count) return;
  copy_blob(P.blob, smem, P.blob_bytes);
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int j = blockIdx.x * kLsWarps + warp;
  if (j >= count) return;
  using LaneT = PLane<M, W, 1>;
  constexpr int nz = LaneT::nz;
  double* stg = reinterpret_cast<double*>(smem + ((P.blob_bytes + 15) / 16) * 16) +
                static_cast<size_t>(warp) * p_roll_stage_doubles<M>(1);
  const int entry = P.list[j];
  const LaneT L(P, smem, entry >> 1, lane, 0, true);
  const int t = ((entry & 1) ? wide_tries : 0) + lane;
  const bool mine = t < P.opt.line_search_max_iterations;
  int st_try = L.is(I_STATUS);
  double* cand = P.CAND + static_cast<size_t>(j) * (P.N + 1) * nz * kWarp + lane;
  p_rollout_only<M, W, 1>(L, stg, mine, L.is(I_ZSEL), cand, kWarp, nz * kWarp, try_alpha(P.opt, t), st_try);
  if (mine) P.TRYST[j * kWarp + lane] = st_try;
}

// same for the candidates of k_roll_deep: item = (list entry j, knot), lane = try.
// COSTK[(j*(N+1) + k)*32 + lane].
template <class M, int W>
__global__ void __launch_bounds__(kLsWarps* kWarp) k_cost_deep(SolverParams P, int wide_tries) {
  extern __shared__ __align__(128) char s_blob[];
  const int count = P.counters[3];
  const int nk = P.N + 1;
  const long items = static_cast<long>(count) * nk;
  if (static_cast<long>(blockIdx.x) * kLsWarps >= items) return;
  copy_blob(P.blob, s_blob, P.blob_bytes);
  constexpr int n = M::n, m = M::m, nz = n + m;
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  for (long w = static_cast<long>(blockIdx.x) * kLsWarps + warp; w < items; w += static_cast<long>(gridDim.x) * kLsWarps) {
    const int j = static_cast<int>(w / nk), k = static_cast<int>(w % nk);
    const int entry = P.list[j];
    if (((entry & 1) ? wide_tries : 0) + lane >= P.opt.line_search_max_iterations) continue;
    const PLane<M, W, 1> L(P, s_blob, entry >> 1, lane, 0, true);
    const double* zc = P.CAND + static_cast<size_t>(w) * nz * kWarp + lane;
    double x[n], u[m];
    ALTRO_UNROLL
    for (int q = 0; q < n; ++q) x[q] = zc[q * kWarp];
    ALTRO_UNROLL
    for (int q = 0; q < m; ++q) u[q] = zc[(n + q) * kWarp];
    const AlPen pen(L.sc(S_PENALTY));
    const double* lam = P.pmax > 0 ? L.lam(k) : nullptr;
    P.COSTK[static_cast<size_t>(w) * kWarp + lane] = knot_cost<n, m, W>(L.D, k, x, u, lam, pen, nullptr);
  }
}

template <class M, int W>
__global__ void __launch_bounds__(kLsWarps* kWarp) k_acc_deep(SolverParams P, int mode, int wide_tries) {
  extern __shared__ __align__(128) char smem[];  // per warp: N doubles for the winner's gains
  const int count = P.counters[3];
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int j = blockIdx.x * kLsWarps + warp;
  if (j >= count) return;
  using LaneT = PLane<M, W, 1>;
  constexpr int nz = LaneT::nz, GZ = kWarp / W, n = M::n, m = M::m;
  const int N = P.N;
  double* gbuf = reinterpret_cast<double*>(smem) + static_cast<size_t>(warp) * N;
  const int entry = P.list[j];
  const LaneT L(P, nullptr, entry >> 1, lane, 0, true);
  const int done0 = (entry & 1) ? wide_tries : 0;
  const int t = done0 + lane;
  const bool mine = t < P.opt.line_search_max_iterations;
  const int zsel = L.is(I_ZSEL);
  int st = L.is(I_STATUS), st_try = st;
  const double J0 = L.sc(S_J0), dV0 = L.sc(S_DV0), dV1 = L.sc(S_DV1);
  double csrc = L.sc(S_CSRC_ALPHA);
  double J = 0.0;
  if (mine) {
    st_try = P.TRYST[j * kWarp + lane];
    const double* ck = P.COSTK + static_cast<size_t>(j) * (N + 1) * kWarp + lane;
    for (int k = 0; k <= N; ++k) J += ck[k * kWarp];
  }
  PLsResult r = p_accept<1>(P.opt, 0, lane, true, mine, mine && st_try == kUnsolved, J, try_alpha(P.opt, t), st_try,
                            J0, dV0, dV1, done0, st, csrc);
  r.exhausted = true;  // one round covers every remaining try (the host guarantees max <= wide + 32)
  const double* cand = P.CAND + static_cast<size_t>(j) * (N + 1) * nz * kWarp;
  const int new_zsel = (zsel + 1) % (GZ + 1);
  if (r.success) {
    for (int k = lane; k < N; k += kWarp)
      gbuf[k] = knot_gain<m>(L.kd(k) + m * n * W, W, cand + (static_cast<size_t>(k) * nz + n) * kWarp + r.slot, kWarp);
    const double* src = cand + r.slot;
    double* dst = L.z(new_zsel, 0);
    const int total = (N + 1) * nz;
    for (int e = lane; e < total; e += kWarp) dst[static_cast<size_t>(e) * W] = src[static_cast<size_t>(e) * kWarp];
    __syncwarp();
    if (lane == 0) {
      double gs = 0.0;
      for (int k = 0; k < N; ++k) gs += gbuf[k];
      r.gsum = gs;
    }
  }
  double gs_bwd = 0.0;
  if (!r.success) gs_bwd = gain_sum_warp(L, true, zsel, gbuf);
  if (lane == 0) p_finish(L, mode, r, new_zsel, zsel, J0, csrc, st, gs_bwd);
}

// First G = 32/W tries of every instance in kPhInner whose previous search did not fail
// completely.  One warp per tile, lane = a*W + i.
#ifndef ALTRO_WIDE_MINB
#define ALTRO_WIDE_MINB 1
#endif
template <class M, int W>
__global__ void __launch_bounds__(kLsWarps* kWarp, ALTRO_WIDE_MINB) k_ls_wide(SolverParams P, int mode) {
  extern __shared__ __align__(128) char smem[];
  copy_blob(P.blob, smem, P.blob_bytes);
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int tile = blockIdx.x * kLsWarps + warp;
  if (tile >= P.T) return;
  using LaneT = PLane<M, W, W>;
  constexpr int G = LaneT::G, nz = LaneT::nz;
  double* stg = reinterpret_cast<double*>(smem + ((P.blob_bytes + 15) / 16) * 16) +
                static_cast<size_t>(warp) * p_stage_doubles<M>(P.pmax, W);
  const int b = tile * W + lane % W;
  const LaneT L(P, smem, b, lane / W, lane % W, b < P.B);
  const bool run = L.valid && L.is(I_PHASE) == kPhInner;
  if (!__any_sync(kFull, run)) return;
  int zsel = 0, st = kUnsolved;
  double penalty = 1.0, J0 = 0.0, dV0 = 0.0, dV1 = 0.0, csrc = -1.0;
  if (run) {
    zsel = L.is(I_ZSEL);
    st = L.is(I_STATUS);
    penalty = L.sc(S_PENALTY);
    J0 = L.sc(S_J0);
    dV0 = L.sc(S_DV0);
    dV1 = L.sc(S_DV1);
    csrc = L.sc(S_CSRC_ALPHA);
  }
  double* zo = L.z((zsel + 1 + L.a) % (G + 1), 0);
  const PLsResult r = p_line_search<M, W, W>(L, stg, run, zsel, zo, W, nz * W, penalty, J0, dV0, dV1, st, csrc, 0, 1);
  if (run && L.a == 0) {
    if (r.success || r.exhausted) {
      p_finish(L, mode, r, (zsel + 1 + r.slot) % (G + 1), zsel, J0, csrc, st,
               r.success ? 0.0 : gain_sum_serial(L, zsel));
    } else {  // continue with try G in k_ls_deep
      L.sc(S_CSRC_ALPHA) = csrc;
      L.is(I_STATUS) = st;
      P.list[atomicAdd(&P.counters[3], 1)] = (b << 1) | 1;
    }
  }
}

// The remaining tries of the instances on P.list, one instance per warp, 32 tries per round.
// entry = (instance << 1) | from_wide: from_wide = 1 -> the first `wide_tries` tries are done.
// The search only needs each candidate's cost: nothing is stored while searching (most entries are
// instances whose search fails at every step length).  The accepted candidate, if any, is rolled
// out once more by its lane straight into the instance's next trajectory buffer, which also yields
// its normalised feed-forward gain — the same arithmetic on the same inputs, hence the same
// trajectory the search evaluated.
#ifndef ALTRO_DEEP_MINB
#define ALTRO_DEEP_MINB 4
#endif
template <class M>
__host__ __device__ constexpr int p_deep_warp_doubles(int pmax, int N, int WI) {
  return p_stage_doubles<M>(pmax, WI) + WI * N;
}
// WI instances per warp, G = 32/WI tries each per round: WI = 2 when the tries left after the wide
// kernel fit 16 lanes (the default 20 - 4), WI = 1 otherwise.
template <class M, int W, int WI>
__global__ void __launch_bounds__(kLsWarps* kWarp, ALTRO_DEEP_MINB) k_ls_deep(SolverParams P, int mode, int wide_tries) {
  extern __shared__ __align__(128) char smem[];
  const int count = P.counters[3];
  if (static_cast<int>(blockIdx.x) * kLsWarps * WI >= count) return;
  copy_blob(P.blob, smem, P.blob_bytes);
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int j0 = (blockIdx.x * kLsWarps + warp) * WI;
  if (j0 >= count) return;
  using LaneT = PLane<M, W, WI>;
  constexpr int nz = LaneT::nz, GZ = kWarp / W;  // GZ + 1 trajectory buffers exist per instance
  double* stg = reinterpret_cast<double*>(smem + ((P.blob_bytes + 15) / 16) * 16) +
                static_cast<size_t>(warp) * p_deep_warp_doubles<M>(P.pmax, P.N, WI);
  const int si = lane % WI, a = lane / WI;
  const bool valid = j0 + si < count;
  const int entry = P.list[valid ? j0 + si : j0];
  const LaneT L(P, smem, entry >> 1, a, si, valid);
  double* gbuf = stg + p_stage_doubles<M>(P.pmax, WI) + si * P.N;
  const int done0 = (entry & 1) ? wide_tries : 0;
  int zsel = L.is(I_ZSEL), st = L.is(I_STATUS);
  const double penalty = L.sc(S_PENALTY), J0 = L.sc(S_J0), dV0 = L.sc(S_DV0), dV1 = L.sc(S_DV1);
  double csrc = L.sc(S_CSRC_ALPHA);
  const int new_zsel = (zsel + 1) % (GZ + 1);
  PLsResult r = p_line_search<M, W, WI, false>(L, stg, valid, zsel, L.z(new_zsel, 0), W, nz * W, penalty, J0, dV0, dV1,
                                               st, csrc, done0, 1 << 30);
  // the last try sits in buffer new_zsel; it is the Q8 candidate iff it was the last rollout that ran to the end
  const double cand_alpha = (csrc == try_alpha(P.opt, P.opt.line_search_max_iterations - 1)) ? csrc : -1.0;
  const bool succ = valid && r.success;
  if (__any_sync(kFull, succ)) {
    double J2, gs;
    int st2 = kUnsolved;
    p_rollout<M, W, WI>(L, stg, succ && a == r.slot, zsel, L.z(new_zsel, 0), W, nz * W, r.alpha, penalty, J2, gs, st2);
    r.gsum = __shfl_sync(kFull, gs, si + WI * (succ ? r.slot : 0));
  }
  const double gs_bwd = gain_sum_warp(L, valid && !r.success, zsel, gbuf);
  if (valid && a == 0) p_finish(L, mode, r, new_zsel, zsel, J0, csrc, st, gs_bwd, cand_alpha);
}

// ------------------------------------------------------------------------------------------
// Latency probe (tools/gpu_microbench.py): cycles of the per-knot device functions for ONE warp,
// each measured as a dependent chain of `reps` calls with clock64().  Not part of any solve.
// out[0..] = cycles per call: dfma, sincos, rk4_step, knot_cost(k=1), knot_cost(k=0),
// quad_eval, al_value(k=1), rk4_jacobian, knot_expansion(k=1, no dynamics), riccati_step
// ------------------------------------------------------------------------------------------
template <class M, int W>
__global__ void k_microbench(SolverParams P, double* sink, long long* out, int reps) {
  extern __shared__ __align__(128) char smem[];
  copy_blob(P.blob, smem, P.blob_bytes);
  constexpr int n = M::n, m = M::m;
  const Desc D(smem);
  const int lane = threadIdx.x;
  double x[n], u[m];
  for (int q = 0; q < n; ++q) x[q] = 0.1 * (q + 1) + 1e-3 * lane;
  for (int q = 0; q < m; ++q) u[q] = 0.05 * (q + 1);
  __shared__ double lam_s[128];
  for (int q = lane; q < 128; q += 32) lam_s[q] = -0.01 * q;
  __syncwarp();
  const double* mp = D.params();
  const AlPen pen10(10.0 + 1e-9 * lane);
  double acc = 0.0;
  long long t0, t1;
  int o = 0;
  // dfma chain
  t0 = clock64();
  { double a = x[0]; for (int r = 0; r < reps * 16; ++r) a = fma(a, 1.0000001, 1e-9); acc += a; }
  t1 = clock64(); if (lane == 0) out[o] = (t1 - t0) / (reps * 16); ++o;
  // sincos chain
  t0 = clock64();
  { double a = x[0], s, c; for (int r = 0; r < reps; ++r) { sincos(a, &s, &c); a = s + c; } acc += a; }
  t1 = clock64(); if (lane == 0) out[o] = (t1 - t0) / reps; ++o;
  // rk4_step chain
  t0 = clock64();
  { double xx[n]; for (int q = 0; q < n; ++q) xx[q] = x[q];
    for (int r = 0; r < reps; ++r) { double xn[n]; rk4_step<M>(mp, xx, u, D.h(1), xn); for (int q = 0; q < n; ++q) xx[q] = xn[q]; }
    acc += xx[0]; }
  t1 = clock64(); if (lane == 0) out[o] = (t1 - t0) / reps; ++o;
  // knot_cost k=1 (chain through x[0])
  for (int kk = 1; kk >= 0; --kk) {
    t0 = clock64();
    { double xx[n]; for (int q = 0; q < n; ++q) xx[q] = x[q];
      for (int r = 0; r < reps; ++r) { double v; const double J = knot_cost<n, m, 1>(D, kk, xx, u, lam_s, pen10, &v); xx[0] += 1e-12 * J; }
      acc += xx[0]; }
    t1 = clock64(); if (lane == 0) out[o] = (t1 - t0) / reps; ++o;
  }
  // quad_eval
  t0 = clock64();
  { double xx[n]; for (int q = 0; q < n; ++q) xx[q] = x[q];
    for (int r = 0; r < reps; ++r) { const double J = quad_eval<n, m>(D.cost(1), xx, u); xx[0] += 1e-12 * J; }
    acc += xx[0]; }
  t1 = clock64(); if (lane == 0) out[o] = (t1 - t0) / reps; ++o;
  // al_value k=1
  t0 = clock64();
  { double xx[n]; for (int q = 0; q < n; ++q) xx[q] = x[q];
    for (int r = 0; r < reps; ++r) { double v; const double J = al_value<n, m, 1>(D.conset(1), xx, u, lam_s, pen10, 0.0, &v); xx[0] += 1e-12 * J; }
    acc += xx[0]; }
  t1 = clock64(); if (lane == 0) out[o] = (t1 - t0) / reps; ++o;
  // rk4_jacobian
  t0 = clock64();
  { double xx[n]; for (int q = 0; q < n; ++q) xx[q] = x[q];
    for (int r = 0; r < reps; ++r) { double A[n * n], B[n * m]; rk4_jacobian<M>(mp, xx, u, D.h(1), A, B); xx[0] += 1e-12 * (A[0] + B[0]); }
    acc += xx[0]; }
  t1 = clock64(); if (lane == 0) out[o] = (t1 - t0) / reps; ++o;
  // knot_expansion without dynamics
  t0 = clock64();
  { double xx[n]; for (int q = 0; q < n; ++q) xx[q] = x[q];
    for (int r = 0; r < reps; ++r) {
      double A[1], B[1], lxx[n * n], lxu[n * m], luu[m * m], lx[n], lu[m];
      knot_expansion<M, 1, false>(D, P.N, 1, xx, u, lam_s, 10.0, A, B, lxx, lxu, luu, lx, lu);
      xx[0] += 1e-12 * (lxx[0] + luu[0] + lx[0] + lu[0] + lxu[0]); }
    acc += xx[0]; }
  t1 = clock64(); if (lane == 0) out[o] = (t1 - t0) / reps; ++o;
  // riccati_step
  t0 = clock64();
  { double A[n * n], B[n * m], lxx[n * n], lxu[n * m], luu[m * m], lx[n], lu[m], Pm[n * n], p[n];
    for (int q = 0; q < n * n; ++q) { A[q] = (q % (n + 1) == 0) ? 1.0 : 0.01 * q; lxx[q] = (q % (n + 1) == 0) ? 1.0 : 0.0; Pm[q] = lxx[q] * 10; }
    for (int q = 0; q < n * m; ++q) { B[q] = 0.1 + 0.01 * q; lxu[q] = 0.0; }
    for (int q = 0; q < m * m; ++q) luu[q] = (q % (m + 1) == 0) ? 0.5 : 0.0;
    for (int q = 0; q < n; ++q) { lx[q] = 0.1 * q; p[q] = 0.2 * q + 1e-3 * lane; }
    for (int q = 0; q < m; ++q) lu[q] = 0.05 * q;
    double dV0 = 0, dV1 = 0;
    for (int r = 0; r < reps; ++r) { double K[m * n], d[m]; riccati_step<n, m>(A, B, lxx, lxu, luu, lx, lu, Pm, p, 1e-8, K, d, &dV0, &dV1);
      for (int q = 0; q < n * n; ++q) Pm[q] = 0.5 * Pm[q] + 5.0 * lxx[q]; }
    acc += Pm[0] + p[0] + dV0 + dV1; }
  t1 = clock64(); if (lane == 0) out[o] = (t1 - t0) / reps; ++o;
  // whole closed-loop rollouts of instance 0 (one instance per warp mapping), cycles per knot
  {
    __shared__ double stg_s[2 * 160];
    const PLane<M, W, 1> L(P, smem, 0, lane, 0, true);
    double* cand = P.CAND + lane;
    const int nzz = n + m;
    double J, g;
    int st = kUnsolved;
#define ALTRO_PROBE(mask)                                                                              \
    t0 = clock64();                                                                                  \
    p_rollout<M, W, 1, mask>(L, stg_s, true, 0, cand, kWarp, nzz * kWarp, 0.5, 10.0, J, g, st);      \
    t1 = clock64();                                                                                  \
    acc += J + g;                                                                                    \
    if (lane == 0) out[o] = (t1 - t0) / (P.N + 1);                                                   \
    ++o;
    ALTRO_PROBE(0)
    ALTRO_PROBE(0)
    ALTRO_PROBE(1)
    ALTRO_PROBE(2)
    ALTRO_PROBE(4)
    ALTRO_PROBE(8)
    ALTRO_PROBE(16)
    ALTRO_PROBE(31)
#undef ALTRO_PROBE
  }
  sink[lane] = acc;
}

}  // namespace altro_b200
