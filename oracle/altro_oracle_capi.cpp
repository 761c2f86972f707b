// oracle/altro_oracle_capi.cpp — C ABI over altro_oracle.hpp.
// TEST INFRASTRUCTURE (see the header of altro_oracle.hpp).
#include "altro_oracle.h"

#include <atomic>
#include <thread>

#include "altro_oracle.hpp"

using namespace altro_oracle;

namespace {

Options FromC(const altro_oracle_options& o) {
  Options r;
  r.max_iterations_total = o.max_iterations_total;
  r.max_iterations_outer = o.max_iterations_outer;
  r.max_iterations_inner = o.max_iterations_inner;
  r.cost_tolerance = o.cost_tolerance;
  r.gradient_tolerance = o.gradient_tolerance;
  r.bp_reg_increase_factor = o.bp_reg_increase_factor;
  r.bp_reg_initial = o.bp_reg_initial;
  r.bp_reg_max = o.bp_reg_max;
  r.bp_reg_min = o.bp_reg_min;
  r.bp_reg_fail_threshold = o.bp_reg_fail_threshold;
  r.check_forwardpass_bounds = o.check_forwardpass_bounds;
  r.state_max = o.state_max;
  r.control_max = o.control_max;
  r.line_search_max_iterations = o.line_search_max_iterations;
  r.line_search_lower_bound = o.line_search_lower_bound;
  r.line_search_upper_bound = o.line_search_upper_bound;
  r.line_search_decrease_factor = o.line_search_decrease_factor;
  r.constraint_tolerance = o.constraint_tolerance;
  r.maximum_penalty = o.maximum_penalty;
  r.initial_penalty = o.initial_penalty;
  r.reset_duals = o.reset_duals;
  r.penalty_scaling = o.penalty_scaling;
  return r;
}

struct ISolver {
  virtual ~ISolver() = default;
  virtual void SetOptions(const Options& o) = 0;
  virtual void SetControls(const double* U) = 0;
  virtual void SetStates(const double* X) = 0;
  virtual void SetInitialState(const double* x0) = 0;
  virtual void SetPenalty(double rho) = 0;
  virtual void SetDuals(int k, const double* lam) = 0;
  virtual void Rollout() = 0;
  virtual double Cost() = 0;
  virtual void UpdateExpansions() = 0;
  virtual void BackwardPass() = 0;
  virtual void ForwardPass() = 0;
  virtual void UpdateConvergenceStatistics() = 0;
  virtual void SolveILQR() = 0;
  virtual void SolveAL() = 0;
  virtual void UpdateDuals() = 0;
  virtual void UpdatePenalties() = 0;
  virtual double MaxViolationStored() = 0;
  virtual double MaxPenalty() = 0;
  virtual void GetTrajectory(double* X, double* U) = 0;
  virtual void GetGains(double* K, double* d) = 0;
  virtual void GetCtg(int k, double* P, double* p) = 0;
  virtual void GetExpansion(int k, double*, double*, double*, double*, double*, double*) = 0;
  virtual void GetActionValue(int k, double*, double*, double*, double*, double*) = 0;
  virtual int NumDuals(int k) = 0;
  virtual void GetDuals(int k, double* lam) = 0;
  virtual void GetStatus(int* out4) = 0;
  virtual void GetScalars(double* out5) = 0;
  virtual int GetStat(int which, double* out, int cap) = 0;
  virtual void GetCounters(long* out4) = 0;
  virtual void ResetStats() = 0;
};

template <int n, int m>
struct SolverT final : ISolver {
  ALSolver<n, m> al;
  int N;
  SolverT(const Problem& P, bool use_constraints) : al(P, use_constraints), N(P.N) {}
  iLQR<n, m>& il() { return al.ilqr; }

  void SetOptions(const Options& o) override { il().opts = o; }
  void SetControls(const double* U) override {
    std::memcpy(il().U.data(), U, sizeof(double) * N * m);
    for (int i = 0; i < m; ++i) il().U[N * m + i] = 0.0;
  }
  void SetStates(const double* X) override {
    std::memcpy(il().X.data(), X, sizeof(double) * (N + 1) * n);
  }
  void SetInitialState(const double* x0) override { il().x0.assign(x0, x0 + n); }
  void SetPenalty(double rho) override { al.SetPenalty(rho); }
  void SetDuals(int k, const double* lam) override {
    for (auto* vec : {&il().costfun[k].eq, &il().costfun[k].ineq})
      for (auto& cv : *vec) {
        for (int i = 0; i < cv.p; ++i) cv.lambda[i] = lam[i];
        lam += cv.p;
      }
  }
  void Rollout() override { il().Rollout(); }
  double Cost() override { return il().Cost(); }
  void UpdateExpansions() override { il().UpdateExpansions(); }
  void BackwardPass() override { il().BackwardPass(); }
  void ForwardPass() override { il().ForwardPass(); }
  void UpdateConvergenceStatistics() override { il().UpdateConvergenceStatistics(); }
  void SolveILQR() override { il().Solve(); }
  void SolveAL() override { al.Solve(); }
  void UpdateDuals() override { al.UpdateDuals(); }
  void UpdatePenalties() override { al.UpdatePenalties(); }
  double MaxViolationStored() override { return al.GetMaxViolation(); }
  double MaxPenalty() override { return al.GetMaxPenalty(); }
  void GetTrajectory(double* X, double* U) override {
    if (X) std::memcpy(X, il().X.data(), sizeof(double) * (N + 1) * n);
    if (U) std::memcpy(U, il().U.data(), sizeof(double) * N * m);
  }
  void GetGains(double* K, double* d) override {
    for (int k = 0; k < N; ++k) {
      if (K) std::memcpy(K + k * m * n, il().kp[k].K, sizeof(double) * m * n);
      if (d) std::memcpy(d + k * m, il().kp[k].d, sizeof(double) * m);
    }
  }
  void GetCtg(int k, double* P, double* p) override {
    if (P) std::memcpy(P, il().kp[k].P, sizeof(double) * n * n);
    if (p) std::memcpy(p, il().kp[k].p, sizeof(double) * n);
  }
  void GetExpansion(int k, double* lxx, double* lxu, double* luu, double* lx, double* lu,
                    double* jac) override {
    auto& f = il().kp[k];
    if (lxx) std::memcpy(lxx, f.lxx, sizeof(f.lxx));
    if (lxu) std::memcpy(lxu, f.lxu, sizeof(f.lxu));
    if (luu) std::memcpy(luu, f.luu, sizeof(f.luu));
    if (lx) std::memcpy(lx, f.lx, sizeof(f.lx));
    if (lu) std::memcpy(lu, f.lu, sizeof(f.lu));
    if (jac) std::memcpy(jac, f.jac, sizeof(f.jac));
  }
  void GetActionValue(int k, double* Qxx, double* Qxu, double* Quu, double* Qx,
                      double* Qu) override {
    auto& f = il().kp[k];
    if (Qxx) std::memcpy(Qxx, f.Qxx, sizeof(f.Qxx));
    if (Qxu) std::memcpy(Qxu, f.Qxu, sizeof(f.Qxu));
    if (Quu) std::memcpy(Quu, f.Quu, sizeof(f.Quu));
    if (Qx) std::memcpy(Qx, f.Qx, sizeof(f.Qx));
    if (Qu) std::memcpy(Qu, f.Qu, sizeof(f.Qu));
  }
  int NumDuals(int k) override {
    int p = 0;
    for (auto& cv : il().costfun[k].eq) p += cv.p;
    for (auto& cv : il().costfun[k].ineq) p += cv.p;
    return p;
  }
  void GetDuals(int k, double* lam) override {
    for (auto* vec : {&il().costfun[k].eq, &il().costfun[k].ineq})
      for (auto& cv : *vec) {
        for (int i = 0; i < cv.p; ++i) lam[i] = cv.lambda[i];
        lam += cv.p;
      }
  }
  void GetStatus(int* o) override {
    o[0] = il().costfun[0].plain ? il().status : al.status;
    o[1] = il().stats.iterations_inner;
    o[2] = il().stats.iterations_outer;
    o[3] = il().stats.iterations_total;
  }
  void GetScalars(double* o) override {
    o[0] = il().rho;
    o[1] = il().drho;
    o[2] = il().deltaV[0];
    o[3] = il().deltaV[1];
    o[4] = il().stats.initial_cost;
  }
  int GetStat(int which, double* out, int cap) override {
    std::vector<double>& v = *il().stats.all[which];
    const int len = static_cast<int>(v.size());
    for (int i = 0; i < std::min(len, cap); ++i) out[i] = v[i];
    return len;
  }
  void ResetStats() override { il().stats.Reset(); }
  void GetCounters(long* o) override {
    o[0] = il().n_backward;
    o[1] = il().n_rollout_cl;
    o[2] = il().n_cost;
    o[3] = il().n_expansions;
  }
};

ISolver* MakeSolver(const Problem& P, bool use_constraints) {
#define ALTRO_ORACLE_CASE(N_, M_) \
  if (P.n == N_ && P.m == M_) return new SolverT<N_, M_>(P, use_constraints);
  ALTRO_ORACLE_CASE(3, 2)
  ALTRO_ORACLE_CASE(6, 2)
  ALTRO_ORACLE_CASE(4, 1)
  ALTRO_ORACLE_CASE(3, 1)
  ALTRO_ORACLE_CASE(9, 3)
  ALTRO_ORACLE_CASE(32, 8)
#undef ALTRO_ORACLE_CASE
  return nullptr;
}

}  // namespace

extern "C" {

void altro_oracle_default_options(altro_oracle_options* o) {
  Options d;
  std::memset(o, 0, sizeof(*o));
  o->max_iterations_total = d.max_iterations_total;
  o->max_iterations_outer = d.max_iterations_outer;
  o->max_iterations_inner = d.max_iterations_inner;
  o->bp_reg_fail_threshold = d.bp_reg_fail_threshold;
  o->check_forwardpass_bounds = d.check_forwardpass_bounds;
  o->line_search_max_iterations = d.line_search_max_iterations;
  o->reset_duals = d.reset_duals;
  o->cost_tolerance = d.cost_tolerance;
  o->gradient_tolerance = d.gradient_tolerance;
  o->bp_reg_increase_factor = d.bp_reg_increase_factor;
  o->bp_reg_initial = d.bp_reg_initial;
  o->bp_reg_max = d.bp_reg_max;
  o->bp_reg_min = d.bp_reg_min;
  o->state_max = d.state_max;
  o->control_max = d.control_max;
  o->line_search_lower_bound = d.line_search_lower_bound;
  o->line_search_upper_bound = d.line_search_upper_bound;
  o->line_search_decrease_factor = d.line_search_decrease_factor;
  o->constraint_tolerance = d.constraint_tolerance;
  o->maximum_penalty = d.maximum_penalty;
  o->initial_penalty = d.initial_penalty;
  o->penalty_scaling = d.penalty_scaling;
}

int altro_oracle_problem_create(int n, int m, int N, void** out) {
  *out = new Problem(n, m, N);
  return 0;
}
void altro_oracle_problem_destroy(void* p) { delete static_cast<Problem*>(p); }

int altro_oracle_problem_set_model(void* p, int kind, const double* params, int nparams) {
  Problem& P = *static_cast<Problem*>(p);
  P.model = kind;
  P.model_params.assign(params, params + nparams);
  return 0;
}
int altro_oracle_problem_set_uniform_step(void* p, float h) {
  static_cast<Problem*>(p)->SetUniformStep(h);
  return 0;
}
int altro_oracle_problem_set_steps(void* p, const float* t, const float* h) {
  Problem& P = *static_cast<Problem*>(p);
  for (int k = 0; k <= P.N; ++k) {  // Trajectory::SetTime / SetStep per knot, altro/common/trajectory.hpp:119-120
    P.t[k] = t[k];
    P.h[k] = h[k];
  }
  return 0;
}
int altro_oracle_problem_set_cost(void* p, int k0, int k1, const double* Q, const double* R,
                                  const double* H, const double* q, const double* r, double c) {
  Problem& P = *static_cast<Problem*>(p);
  if (k0 < 0 || k1 > P.N + 1 || k0 >= k1) return -1;
  const int n = P.n, m = P.m;
  auto qc = std::make_shared<QuadCost>();
  qc->Q.assign(Q, Q + n * n);
  qc->R.assign(R, R + m * m);
  qc->H.assign(H, H + n * m);
  qc->q.assign(q, q + n);
  qc->r.assign(r, r + m);
  qc->c = c;
  for (int k = k0; k < k1; ++k) P.cost[k] = qc;
  return 0;
}
int altro_oracle_problem_add_goal(void* p, int k, const double* xf) {
  Problem& P = *static_cast<Problem*>(p);
  ConstraintDef d;
  d.kind = kConGoal;
  d.equality = true;
  d.p = P.n;
  d.xf.assign(xf, xf + P.n);
  P.eq[k].push_back(d);
  return 0;
}
int altro_oracle_problem_add_control_bound(void* p, int k, const double* lb, const double* ub) {
  Problem& P = *static_cast<Problem*>(p);
  ConstraintDef d;
  d.kind = kConControlBound;
  d.equality = false;
  d.lb.assign(lb, lb + P.m);
  d.ub.assign(ub, ub + P.m);
  for (int i = 0; i < P.m; ++i) {  // basic_constraints.hpp:136-143
    if (std::fabs(ub[i]) < DBL_MAX) d.idx_ub.push_back(i);
    if (std::fabs(lb[i]) < DBL_MAX) d.idx_lb.push_back(i);
  }
  d.p = static_cast<int>(d.idx_lb.size() + d.idx_ub.size());
  P.ineq[k].push_back(d);
  return 0;
}
int altro_oracle_problem_add_circles(void* p, int k, int nc, const double* cx, const double* cy,
                                     const double* cr, int xi, int yi) {
  Problem& P = *static_cast<Problem*>(p);
  ConstraintDef d;
  d.kind = kConCircle;
  d.equality = false;
  d.p = nc;
  d.cx.assign(cx, cx + nc);
  d.cy.assign(cy, cy + nc);
  d.cr.assign(cr, cr + nc);
  d.xi = xi;
  d.yi = yi;
  P.ineq[k].push_back(d);
  return 0;
}
int altro_oracle_problem_set_initial_state(void* p, const double* x0) {
  Problem& P = *static_cast<Problem*>(p);
  P.x0.assign(x0, x0 + P.n);
  return 0;
}

void* altro_oracle_solver_create(const void* prob, int use_constraints) {
  return MakeSolver(*static_cast<const Problem*>(prob), use_constraints != 0);
}
void altro_oracle_solver_destroy(void* s) { delete static_cast<ISolver*>(s); }
#define S(s) static_cast<ISolver*>(s)
int altro_oracle_solver_set_options(void* s, const altro_oracle_options* o) {
  S(s)->SetOptions(FromC(*o));
  return 0;
}
int altro_oracle_solver_set_controls(void* s, const double* U) { S(s)->SetControls(U); return 0; }
int altro_oracle_solver_set_states(void* s, const double* X) { S(s)->SetStates(X); return 0; }
int altro_oracle_solver_set_initial_state(void* s, const double* x0) {
  S(s)->SetInitialState(x0);
  return 0;
}
int altro_oracle_solver_set_penalty(void* s, double rho) { S(s)->SetPenalty(rho); return 0; }
int altro_oracle_solver_set_duals(void* s, int k, const double* lam) {
  S(s)->SetDuals(k, lam);
  return 0;
}
void altro_oracle_solver_rollout(void* s) { S(s)->Rollout(); }
double altro_oracle_solver_cost(void* s) { return S(s)->Cost(); }
void altro_oracle_solver_update_expansions(void* s) { S(s)->UpdateExpansions(); }
void altro_oracle_solver_backward_pass(void* s) { S(s)->BackwardPass(); }
void altro_oracle_solver_forward_pass(void* s) { S(s)->ForwardPass(); }
void altro_oracle_solver_update_convergence_statistics(void* s) {
  S(s)->UpdateConvergenceStatistics();
}
void altro_oracle_solver_solve_ilqr(void* s) { S(s)->SolveILQR(); }
void altro_oracle_solver_solve_al(void* s) { S(s)->SolveAL(); }
void altro_oracle_solver_update_duals(void* s) { S(s)->UpdateDuals(); }
void altro_oracle_solver_update_penalties(void* s) { S(s)->UpdatePenalties(); }
double altro_oracle_solver_max_violation_stored(void* s) { return S(s)->MaxViolationStored(); }
double altro_oracle_solver_max_penalty(void* s) { return S(s)->MaxPenalty(); }
void altro_oracle_solver_get_trajectory(void* s, double* X, double* U) {
  S(s)->GetTrajectory(X, U);
}
void altro_oracle_solver_get_gains(void* s, double* K, double* d) { S(s)->GetGains(K, d); }
void altro_oracle_solver_get_ctg(void* s, int k, double* P, double* p) { S(s)->GetCtg(k, P, p); }
void altro_oracle_solver_get_expansion(void* s, int k, double* lxx, double* lxu, double* luu,
                                       double* lx, double* lu, double* jac) {
  S(s)->GetExpansion(k, lxx, lxu, luu, lx, lu, jac);
}
void altro_oracle_solver_get_action_value(void* s, int k, double* Qxx, double* Qxu, double* Quu,
                                          double* Qx, double* Qu) {
  S(s)->GetActionValue(k, Qxx, Qxu, Quu, Qx, Qu);
}
int altro_oracle_solver_num_duals(void* s, int k) { return S(s)->NumDuals(k); }
void altro_oracle_solver_get_duals(void* s, int k, double* lam) { S(s)->GetDuals(k, lam); }
void altro_oracle_solver_get_status(void* s, int* out4) { S(s)->GetStatus(out4); }
void altro_oracle_solver_get_scalars(void* s, double* out5) { S(s)->GetScalars(out5); }
int altro_oracle_solver_get_stat(void* s, int which, double* out, int cap) {
  return S(s)->GetStat(which, out, cap);
}
void altro_oracle_solver_get_counters(void* s, long* out4) { S(s)->GetCounters(out4); }
#undef S

int altro_oracle_solve_batch(const void* prob, const altro_oracle_options* o, int use_al, int B,
                             const double* x0s, const double* U0s, const double* U0,
                             int nthreads, double* X, double* U, double* K, double* d,
                             double* cost, double* viol, int* status, int* iters) {
  const Problem& P = *static_cast<const Problem*>(prob);
  const int n = P.n, m = P.m, N = P.N;
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0);
  std::atomic<int> err(0);
  const Options opts = FromC(*o);
  auto worker = [&]() {
    // One solver object per thread, re-used across instances exactly like
    // perf/benchmark_unicycle.cpp:45-78 re-uses one solver across runs.
    std::unique_ptr<ISolver> s(MakeSolver(P, use_al != 0));
    if (!s) {
      err = 1;
      return;
    }
    s->SetOptions(opts);
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      s->SetInitialState(x0s + static_cast<size_t>(b) * n);
      s->SetControls(U0s ? U0s + static_cast<size_t>(b) * N * m : U0);
      if (use_al) {
        s->SolveAL();
      } else {
        s->ResetStats();  // a fresh iLQR<n,m> per instance (iLQR::Solve never resets its stats)
        s->SolveILQR();
      }
      if (X || U)
        s->GetTrajectory(X ? X + static_cast<size_t>(b) * (N + 1) * n : nullptr,
                         U ? U + static_cast<size_t>(b) * N * m : nullptr);
      if (K || d)
        s->GetGains(K ? K + static_cast<size_t>(b) * N * m * n : nullptr,
                    d ? d + static_cast<size_t>(b) * N * m : nullptr);
      if (viol) viol[b] = s->MaxViolationStored();  // stored c_ (Q8), before Cost() refreshes it
      if (cost) cost[b] = s->Cost();
      int st[4];
      s->GetStatus(st);
      if (status) status[b] = st[0];
      if (iters) {
        iters[3 * b + 0] = st[1];
        iters[3 * b + 1] = st[2];
        iters[3 * b + 2] = st[3];
      }
    }
  };
  std::vector<std::thread> pool;
  for (int i = 1; i < nthreads; ++i) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  return err.load();
}

}  // extern "C"
