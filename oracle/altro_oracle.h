/* oracle/altro_oracle.h — C ABI of the CPU oracle (TEST INFRASTRUCTURE, see
 * altro_oracle.hpp).  Loaded with ctypes by tests/ and by bench.py's
 * cpu_baseline / --impl reference legs only.
 *
 * The problem-builder calls take the same arguments as the product ABI in
 * include/altro_b200.h so one Python description drives both.  All matrices are
 * column-major doubles; trajectories are instance-major [B][N+1][n] / [B][N][m].
 */
#ifndef ALTRO_ORACLE_H_
#define ALTRO_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct altro_oracle_options {
  int max_iterations_total;
  int max_iterations_outer;
  int max_iterations_inner;
  int bp_reg_fail_threshold;
  int check_forwardpass_bounds;
  int line_search_max_iterations;
  int reset_duals;
  int _pad;
  double cost_tolerance;
  double gradient_tolerance;
  double bp_reg_increase_factor;
  double bp_reg_initial;
  double bp_reg_max;
  double bp_reg_min;
  double state_max;
  double control_max;
  double line_search_lower_bound;
  double line_search_upper_bound;
  double line_search_decrease_factor;
  double constraint_tolerance;
  double maximum_penalty;
  double initial_penalty;
  double penalty_scaling;
} altro_oracle_options;

void altro_oracle_default_options(altro_oracle_options* o);

/* ---- problem description ---- */
int altro_oracle_problem_create(int n, int m, int N, void** out);
void altro_oracle_problem_destroy(void* p);
int altro_oracle_problem_set_model(void* p, int kind, const double* params, int nparams);
int altro_oracle_problem_set_uniform_step(void* p, float h);
int altro_oracle_problem_set_steps(void* p, const float* t, const float* h);
int altro_oracle_problem_set_cost(void* p, int k0, int k1, const double* Q, const double* R,
                                  const double* H, const double* q, const double* r, double c);
int altro_oracle_problem_add_goal(void* p, int k, const double* xf);
int altro_oracle_problem_add_control_bound(void* p, int k, const double* lb, const double* ub);
int altro_oracle_problem_add_circles(void* p, int k, int ncircles, const double* cx,
                                     const double* cy, const double* cr, int xi, int yi);
int altro_oracle_problem_set_initial_state(void* p, const double* x0);

/* ---- single-instance solver (step-wise, mirrors iLQR / AugmentedLagrangianiLQR) ---- */
void* altro_oracle_solver_create(const void* prob, int use_constraints);
void altro_oracle_solver_destroy(void* s);
int altro_oracle_solver_set_options(void* s, const altro_oracle_options* o);
int altro_oracle_solver_set_controls(void* s, const double* U);   /* [N][m] */
int altro_oracle_solver_set_states(void* s, const double* X);     /* [N+1][n] */
int altro_oracle_solver_set_initial_state(void* s, const double* x0);
int altro_oracle_solver_set_penalty(void* s, double rho);
int altro_oracle_solver_set_duals(void* s, int k, const double* lambda); /* eq rows then ineq rows */
void altro_oracle_solver_rollout(void* s);
double altro_oracle_solver_cost(void* s);
void altro_oracle_solver_update_expansions(void* s);
void altro_oracle_solver_backward_pass(void* s);
void altro_oracle_solver_forward_pass(void* s);
void altro_oracle_solver_update_convergence_statistics(void* s);
void altro_oracle_solver_solve_ilqr(void* s);
void altro_oracle_solver_solve_al(void* s);
void altro_oracle_solver_update_duals(void* s);
void altro_oracle_solver_update_penalties(void* s);
double altro_oracle_solver_max_violation_stored(void* s);
double altro_oracle_solver_max_penalty(void* s);
/* getters */
void altro_oracle_solver_get_trajectory(void* s, double* X, double* U); /* [N+1][n], [N][m] */
void altro_oracle_solver_get_gains(void* s, double* K, double* d);      /* [N][m*n] col-major, [N][m] */
void altro_oracle_solver_get_ctg(void* s, int k, double* P, double* p);
/* lxx,lxu,luu,lx,lu,jac(n x (n+m)) of knot k */
void altro_oracle_solver_get_expansion(void* s, int k, double* lxx, double* lxu, double* luu,
                                       double* lx, double* lu, double* jac);
/* Qxx,Qxu,Quu,Qx,Qu of knot k */
void altro_oracle_solver_get_action_value(void* s, int k, double* Qxx, double* Qxu, double* Quu,
                                          double* Qx, double* Qu);
int altro_oracle_solver_num_duals(void* s, int k);
void altro_oracle_solver_get_duals(void* s, int k, double* lambda);
/* status, iterations_inner, iterations_outer, iterations_total */
void altro_oracle_solver_get_status(void* s, int* out4);
/* regularization rho, drho, deltaV[0], deltaV[1], initial_cost */
void altro_oracle_solver_get_scalars(void* s, double* out5);
/* copies min(len,cap) entries of the named stat vector; returns len.
 * which: 0 cost,1 alpha,2 z,3 gradient,4 cost_decrease,5 regularization,6 violations,7 max_penalty */
int altro_oracle_solver_get_stat(void* s, int which, double* out, int cap);
/* n_backward, n_rollout_cl, n_cost, n_expansions */
void altro_oracle_solver_get_counters(void* s, long* out4);

/* ---- batched solve over host threads (one independent solve per instance) ----
 * x0s [B][n]; U0s [B][N][m] or NULL (then U0 [N][m] is used for every instance).
 * Outputs (any may be NULL): X [B][N+1][n], U [B][N][m], K [B][N][m*n], d [B][N][m],
 * cost [B], viol [B], status [B], iters [B][3] (inner,outer,total).
 * use_al: 1 = AugmentedLagrangianiLQR::Solve, 0 = iLQR::Solve on the bare cost.
 * Returns 0 on success. */
int altro_oracle_solve_batch(const void* prob, const altro_oracle_options* o, int use_al, int B,
                             const double* x0s, const double* U0s, const double* U0,
                             int nthreads, double* X, double* U, double* K, double* d,
                             double* cost, double* viol, int* status, int* iters);

#ifdef __cplusplus
}
#endif
#endif
