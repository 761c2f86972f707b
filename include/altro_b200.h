/* include/altro_b200.h — C ABI of the B200-native batched AL-iLQR solver.
 *
 * The reference (optimusride/altro-cpp @ d5e8cfe) has no FFI: its boundary is the
 * public C++ class API.  Each entry point below names the reference
 * interface it replaces (file:line relative to the reference root).  A reference
 * maintainer binds these from host C++; the Python package
 * altro_cpp_b200 binds the same symbols with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative altro_b200_status on failure;
 *     altro_b200_last_error() returns a thread-local message for the last failure.
 *   - matrices are column-major double (Eigen's default, altro/eigentypes.hpp:8-27);
 *     `t` and `h` are float exactly as in the reference (altro/common/knotpoint.hpp:179-180).
 *   - batch arrays are instance-major:  x0 [B][n], U [B][N][m], X [B][N+1][n],
 *     K [B][N][m*n] (each K col-major m x n), d [B][N][m].
 *   - pointers named *_dev are device pointers on the solver's device, all others host.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - a problem handle is immutable once a solver has been created from it; a solver
 *     handle is not thread-safe, distinct handles are independent (reference threading
 *     contract, altro/problem/problem.hpp:129-137).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     ALTRO_B200_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef ALTRO_B200_H_
#define ALTRO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum altro_b200_status {
  ALTRO_B200_OK = 0,
  ALTRO_B200_ERR_ARG = -1,         /* bad argument (the reference would ALTRO_ASSERT) */
  ALTRO_B200_ERR_UNSUPPORTED = -2, /* (n,m,model) has no device instantiation */
  ALTRO_B200_ERR_CUDA = -3,        /* CUDA runtime error / no device */
  ALTRO_B200_ERR_STATE = -4        /* call order violated (e.g. solve before inputs) */
} altro_b200_status;

/* altro/common/solver_stats.hpp:20-31 (same numeric values) */
typedef enum altro_b200_solver_status {
  ALTRO_B200_SOLVED = 0,
  ALTRO_B200_UNSOLVED = 1,
  ALTRO_B200_STATE_LIMIT = 2,
  ALTRO_B200_CONTROL_LIMIT = 3,
  ALTRO_B200_COST_INCREASE = 4,
  ALTRO_B200_MAX_ITERATIONS = 5,
  ALTRO_B200_MAX_OUTER_ITERATIONS = 6,
  ALTRO_B200_MAX_INNER_ITERATIONS = 7,
  ALTRO_B200_MAX_PENALTY = 8,
  ALTRO_B200_BACKWARD_PASS_REGULARIZATION_FAILED = 9
} altro_b200_solver_status;

/* Continuous-time models with a device functor (closed registry). */
typedef enum altro_b200_model {
  ALTRO_B200_MODEL_UNICYCLE = 0,          /* examples/unicycle.cpp:12-33, n=3 m=2 */
  ALTRO_B200_MODEL_TRIPLE_INTEGRATOR = 1, /* examples/triple_integrator.cpp:9-33, n=3m */
  ALTRO_B200_MODEL_CARTPOLE = 2,          /* new (BASELINE config C4), params mc,mp,l,g */
  ALTRO_B200_MODEL_LINEAR = 3             /* new (BASELINE config C5), discrete x+=Ax+Bu; n=32, m=8,
                                             unconstrained, whole solves only (one instance per CTA) */
} altro_b200_model;

/* altro/common/solver_options.hpp:19-65, numeric fields only (logging/profiler/thread
 * knobs have no device meaning).  penalty_scaling = ConstraintValues::kDefaultPenaltyScaling
 * (altro/constraints/constraint_values.hpp:30), settable like SetPenaltyScaling
 * (altro/augmented_lagrangian/al_solver.hpp:278-284). */
typedef struct altro_b200_options {
  int32_t max_iterations_total;
  int32_t max_iterations_outer;
  int32_t max_iterations_inner;
  int32_t bp_reg_fail_threshold;
  int32_t check_forwardpass_bounds;
  int32_t line_search_max_iterations;
  int32_t reset_duals;
  int32_t skip_repeated_iterations; /* extension, default 0: account for provably identical repeated
                                       inner iterations without executing them (same results) */
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
} altro_b200_options;

typedef struct altro_b200_problem altro_b200_problem;
typedef struct altro_b200_solver altro_b200_solver;

const char* altro_b200_last_error(void);
const char* altro_b200_version(void);
/* SolverOptions() defaults, altro/common/solver_options.hpp:23-56 */
void altro_b200_default_options(altro_b200_options* o);
/* 1 if (n, m, model) has a device instantiation */
int altro_b200_is_supported(int n, int m, int model);

/* ------------------------------------------------------------------------------------
 * Problem description  (replaces problem::Problem, altro/problem/problem.hpp:65-309)
 * ---------------------------------------------------------------------------------- */
/* Problem(N) with state/control dimensions; altro/problem/problem.hpp:76-77 */
int altro_b200_problem_create(int n, int m, int N, altro_b200_problem** out);
void altro_b200_problem_destroy(altro_b200_problem* p);
/* SetDynamics(DiscretizedModel<Model, RungeKutta4>, k) for every k < N
 * (altro/problem/problem.hpp:155-166, discretized_model.hpp:25, integration.hpp:113).
 * One model per problem; ALTRO_B200_MODEL_LINEAR takes params = [A (n*n), B (n*m)]. */
int altro_b200_problem_set_model(altro_b200_problem* p, int model, const double* params,
                                 int nparams);
/* Trajectory::SetUniformStep(h), altro/common/trajectory.hpp:122-130 */
int altro_b200_problem_set_uniform_step(altro_b200_problem* p, float h);
/* Per-knot times and steps (Trajectory::SetTime / SetStep, altro/common/trajectory.hpp:119-120): t[N+1], h[N+1]
 * as the trajectory stores them (h[N] is the terminal knot's, normally 0); h[k] > 0 for k < N. */
int altro_b200_problem_set_steps(altro_b200_problem* p, const float* t, const float* h);
/* SetCostFunction(QuadraticCost(Q,R,H,q,r,c), k) for k0 <= k < k1
 * (altro/problem/problem.hpp:113-121; examples/quadratic_cost.hpp:12-27) */
int altro_b200_problem_set_cost(altro_b200_problem* p, int k0, int k1, const double* Q,
                                const double* R, const double* H, const double* q,
                                const double* r, double c);
/* ------------------------------------------------------------------------------------
 * Plug-in dynamics models.  The reference is extended by subclassing
 * ContinuousDynamics (altro/problem/dynamics.hpp:59-99) — host virtuals the device cannot call.
 * Here a model is the CUDA source text of a functor struct (the concept is documented in
 * altro_cpp_b200/csrc/device.cuh, an example is altro_cpp_b200/plugins/cartpole.cuh); the library
 * instantiates its kernel templates on it with NVRTC when a solver for it is created, caches the
 * module on disk, and from then on treats it like a built-in model (RK4-discretised).
 *   name        the struct's name (a C++ identifier); registering the same (name, source) twice
 *               returns the same id
 *   *model_id   the id to pass to altro_b200_problem_set_model
 * ---------------------------------------------------------------------------------- */
int altro_b200_register_model(const char* name, const char* cuda_source, int n, int m, int nparams,
                              int* model_id);
/* compile the model's module into the on-disk cache now (NVRTC only: needs no GPU); cache_path
 * (optional) receives the file name */
int altro_b200_precompile_model(int model_id, char* cache_path, int cache_path_cap);
/* the same for a built-in model id whose (n, m) instantiation is produced at run time
 * (ALTRO_B200_MODEL_TRIPLE_INTEGRATOR with dof = m other than 2; ALTRO_B200_MODEL_CARTPOLE) */
int altro_b200_precompile_builtin_model(int model, int n, int m, char* cache_path, int cache_path_cap);

/* SetConstraint(GoalConstraint(xf), k)  — Equality; examples/basic_constraints.hpp:15-40 */
int altro_b200_problem_add_goal(altro_b200_problem* p, int k, const double* xf);
/* SetConstraint(ControlBound(lb,ub), k) — Inequality; examples/basic_constraints.hpp:42-150.
 * Infinite entries produce no row (GetFiniteIndices :136-143). */
int altro_b200_problem_add_control_bound(altro_b200_problem* p, int k, const double* lb,
                                         const double* ub);
/* SetConstraint(CircleConstraint{AddObstacle(cx,cy,r)...}, k) — Inequality;
 * examples/obstacle_constraints.hpp:69-126 */
int altro_b200_problem_add_circles(altro_b200_problem* p, int k, int ncircles, const double* cx,
                                   const double* cy, const double* cr, int xi, int yi);
/* the same with squared radii (what a CircleConstraint evaluates at a circle's centre): lets a caller
 * that only sees the functor's virtual interface describe it without losing a bit */
int altro_b200_problem_add_circles_r2(altro_b200_problem* p, int k, int ncircles, const double* cx,
                                      const double* cy, const double* cr2, int xi, int yi);
/* SetInitialState(x0): the nominal initial state, altro/problem/problem.hpp:195-202 */
int altro_b200_problem_set_initial_state(altro_b200_problem* p, const double* x0);

/* ------------------------------------------------------------------------------------
 * Batched solver (replaces AugmentedLagrangianiLQR<n,m> / iLQR<n,m> for B independent
 * instances of one Problem; altro/augmented_lagrangian/al_solver.hpp:28, altro/ilqr/ilqr.hpp:47)
 * ---------------------------------------------------------------------------------- */
/* AugmentedLagrangianiLQR<n,m>(prob) x batch.  use_constraints = 0 builds the plain
 * iLQR<n,m>(prob) that ignores constraints (ilqr.hpp:51-54). device = CUDA ordinal. */
int altro_b200_solver_create(const altro_b200_problem* p, int batch, int use_constraints,
                             int device, altro_b200_solver** out);
void altro_b200_solver_destroy(altro_b200_solver* s);
/* solver.GetOptions() = *o */
int altro_b200_solver_set_options(altro_b200_solver* s, const altro_b200_options* o);
int altro_b200_solver_batch(const altro_b200_solver* s);

/* --- inputs: SetTrajectory(initial guess) + per-instance initial state
 * (ilqr.hpp:231-235, problem.hpp:195-202).  U0 may be NULL together with u_nominal != NULL:
 * every knot of every instance gets u_nominal[m] (InitialTrajectory(),
 * examples/problems/unicycle.hpp:84-93). */
int altro_b200_solver_set_inputs_host(altro_b200_solver* s, const double* x0, const double* U0,
                                      const double* u_nominal, void* stream);
int altro_b200_solver_set_inputs_dev(altro_b200_solver* s, const double* x0_dev,
                                     const double* U0_dev, const double* u_nominal,
                                     void* stream);
/* Z->State(k) for all k (only needed by step-wise tests that skip Rollout) */
int altro_b200_solver_set_states_host(altro_b200_solver* s, const double* X, void* stream);
/* SetPenalty(rho), al_solver.hpp:271-276 */
int altro_b200_solver_set_penalty(altro_b200_solver* s, double rho, void* stream);
/* stats.initial_cost = Cost() (ilqr.hpp:292 there): the cost the first inner iteration's decrease is measured
 * against.  A whole solve sets it itself; a caller who runs the iterations step by step sets it here. */
int altro_b200_solver_set_initial_cost(altro_b200_solver* s, double cost, void* stream);
/* GetALCost(k)->Get{Equality,Inequality}Constraints()[i]->GetDuals() = lambda, rows in
 * ALCost order (equalities then inequalities, al_cost.hpp:264-273); same duals for every instance */
int altro_b200_solver_set_duals_host(altro_b200_solver* s, int k, const double* lambda, int p,
                                     void* stream);

/* --- execution engine -----------------------------------------------------------------
 * Both engines run the same per-instance arithmetic and return the same results.
 *  FUSED : one persistent warp-per-tile kernel (k_solve) runs whole solves.
 *  PHASED: every inner iteration of the batch is UpdateExpansions -> BackwardPass (TMA-streamed
 *          over the materialised expansions, the reference's data flow ilqr.hpp:300-313) ->
 *          ForwardPass kernels, each with its own thread mapping; default for n <= 6.
 * The engine is fixed when a solver is created: altro_b200_set_default_engine() (process-wide),
 * else the environment variable ALTRO_B200_ENGINE=fused|phased, else PHASED. */
#define ALTRO_B200_ENGINE_FUSED 0
#define ALTRO_B200_ENGINE_PHASED 1
void altro_b200_set_default_engine(int engine);
int altro_b200_solver_engine(const altro_b200_solver* s);

/* --- whole solves, device resident, no host round trip -------------------------------
 * AugmentedLagrangianiLQR::Solve(), al_solver.hpp:304-334 */
int altro_b200_solve_al(altro_b200_solver* s, void* stream);
/* iLQR::Solve(), ilqr.hpp:284-316 */
int altro_b200_solve_ilqr(altro_b200_solver* s, void* stream);
/* Host-buffer convenience = the call a reference user makes: inputs from host memory,
 * Solve(), results back to host memory (all on `stream`, synchronised before return).
 * Any output pointer may be NULL. iters = [B][3] (inner, outer, total). */
int altro_b200_solve_al_host(altro_b200_solver* s, const double* x0, const double* U0,
                             const double* u_nominal, double* X, double* U, double* cost,
                             double* viol, int32_t* status, int32_t* iters, void* stream);

/* --- step-wise phases (public methods of iLQR<n,m>; used by parity tests and ncu) ---- */
int altro_b200_rollout(altro_b200_solver* s, void* stream);             /* ilqr.hpp:453-459 */
int altro_b200_cost(altro_b200_solver* s, void* stream);                /* ilqr.hpp:326-334 */
int altro_b200_update_expansions(altro_b200_solver* s, void* stream);   /* ilqr.hpp:350-366 */
int altro_b200_backward_pass(altro_b200_solver* s, void* stream);       /* ilqr.hpp:385-445 */
int altro_b200_forward_pass(altro_b200_solver* s, void* stream);        /* ilqr.hpp:512-558 */
int altro_b200_update_convergence_statistics(altro_b200_solver* s, void* stream); /* :568-587 */
int altro_b200_update_duals(altro_b200_solver* s, void* stream);        /* al_solver.hpp:336-345 */
/* AugmentedLagrangianiLQR::Init(): duals reset (reset_duals), penalties set to initial_penalty (> 0), outer and total
 * iteration counters cleared — al_solver.hpp:287-302.  A whole solve does this itself. */
int altro_b200_al_init(altro_b200_solver* s, void* stream);
int altro_b200_update_penalties(altro_b200_solver* s, void* stream);    /* al_solver.hpp:347-355 */

/* SolveSetup(), ilqr.hpp:629-645 (resets iterations_inner, status, regularisation, deltaV) */
int altro_b200_solve_setup(altro_b200_solver* s, void* stream);
/* Measurement / test variants of BackwardPass (not part of the reference surface):
 *   _stream_only : the materialised backward pass writing only K and d (the SURVEY.md 8d
 *                  contract bytes; altro_b200_backward_pass additionally stores P, p per knot
 *                  for GetCostToGo*());
 *   _fused       : UpdateExpansions fused into the backward sweep (what the solve kernel runs). */
int altro_b200_backward_pass_stream_only(altro_b200_solver* s, void* stream);
int altro_b200_backward_pass_fused(altro_b200_solver* s, void* stream);
/* the backward-pass kernel as a whole solve launches it (per-instance phase mask, regularisation
 * hand-off), every instance marked active; needs update_expansions first and scrambles the solve
 * state — measurement only (bench.py roofline) */
int altro_b200_backward_pass_insolve(altro_b200_solver* s, void* stream);

/* --- outputs (any pointer may be NULL) -----------------------------------------------
 * GetTrajectory() (ilqr.hpp:140) -> X [B][N+1][n], U [B][N][m] */
int altro_b200_get_trajectory_host(altro_b200_solver* s, double* X, double* U, void* stream);
int altro_b200_get_trajectory_dev(altro_b200_solver* s, double* X_dev, double* U_dev,
                                  void* stream);
/* GetKnotPointFunction(k).GetFeedbackGain()/GetFeedforwardGain()
 * (knot_point_function_type.hpp:265-268) -> K [B][N][m*n], d [B][N][m] */
int altro_b200_get_gains_host(altro_b200_solver* s, double* K, double* d, void* stream);
/* GetCostToGoHessian/Gradient of knot k (knot_point_function_type.hpp:254-255), valid after
 * altro_b200_backward_pass -> P [B][n*n], p [B][n] */
int altro_b200_get_ctg_host(altro_b200_solver* s, int k, double* P, double* p, void* stream);
/* GetCosts() (ilqr.hpp:163 there): costs[B][N+1], the per-knot costs written by the last Cost() or
 * UpdateExpansions() of the step-wise API (augmented-Lagrangian terms included) */
int altro_b200_get_costs_host(altro_b200_solver* s, double* costs, void* stream);
/* GetCostExpansion()/GetDynamicsExpansion() of knot k (knot_point_function_type.hpp:249-252),
 * valid after altro_b200_update_expansions. Each [B][...] col-major: A n*n, Bm n*m, lxx n*n,
 * lxu n*m, luu m*m, lx n, lu m */
int altro_b200_get_expansion_host(altro_b200_solver* s, int k, double* A, double* Bm, double* lxx,
                                  double* lxu, double* luu, double* lx, double* lu, void* stream);
/* duals of knot k, ALCost order -> lambda [B][p]; returns p via *p_out */
int altro_b200_get_duals_host(altro_b200_solver* s, int k, double* lambda, int* p_out,
                              void* stream);
/* Constraint values c(x_k, u_k) of the current trajectory at knot k, ALCost row order -> c [B][pmax]
 * (pmax = the row count returned by altro_b200_get_duals_host; rows past *p_out, the knot's own
 * count, are zero).  What GetALCost(k)->...->GetConstraintValue() holds after Cost(); feeds
 * GetConstraintInfo()/PrintViolations() (al_solver.hpp:68-104, constraint_values.hpp:216-221). */
int altro_b200_get_constraint_values_host(altro_b200_solver* s, int k, double* c, int* p_out,
                                          void* stream);
/* Per-instance results of the last solve / phase:
 *   cost   [B]   Cost() of the current trajectory under the current duals/penalties
 *   viol   [B]   GetMaxViolation() (al_solver.hpp:417-422)
 *   status [B]   GetStatus()
 *   iters  [B][3] iterations_inner, iterations_outer, iterations_total (solver_stats.hpp:50-52) */
int altro_b200_get_results_host(altro_b200_solver* s, double* cost, double* viol, int32_t* status,
                                int32_t* iters, void* stream);
/* iLQR<n,m>::GetStatus() of the inner solver (al_solver.hpp GetiLQRSolver().GetStatus()) */
int altro_b200_get_ilqr_status_host(altro_b200_solver* s, int32_t* status, void* stream);
/* More per-instance scalars, [B] each: regularisation rho_ (GetRegularization()), deltaV[0],
 * deltaV[1], last accepted alpha, last z, last dJ, last grad, penalty, initial_cost */
int altro_b200_get_scalars_host(altro_b200_solver* s, double* reg, double* dV0, double* dV1,
                                double* alpha, double* z, double* dJ, double* grad,
                                double* penalty, double* initial_cost, void* stream);

/* ------------------------------------------------------------------------------------
 * The same batch sharded over several GPUs of one node (replaces the reference's
 * SolverOptions::nthreads thread pool, altro/ilqr/ilqr.hpp:354-365, 707-738, one level up):
 * contiguous slices of the batch, one solver and one host thread per device, no collective.
 * Per-instance results do not depend on the placement (bit-identical to a one-GPU solve).
 * ---------------------------------------------------------------------------------- */
typedef struct altro_b200_multi altro_b200_multi;
int altro_b200_multi_create(const altro_b200_problem* p, int batch, int use_constraints, const int* devices,
                            int ndevices, altro_b200_multi** out);
void altro_b200_multi_destroy(altro_b200_multi* m);
int altro_b200_multi_num_devices(const altro_b200_multi* m);
int altro_b200_multi_set_options(altro_b200_multi* m, const altro_b200_options* o);
/* AugmentedLagrangianiLQR::Solve() for the whole batch with host buffers (layouts as
 * altro_b200_solve_al_host); blocks until every device is done */
int altro_b200_multi_solve_al_host(altro_b200_multi* m, const double* x0, const double* U0, const double* u_nominal,
                                   double* X, double* U, double* cost, double* viol, int32_t* status,
                                   int32_t* iters);
/* device time of the last solve per device [ndevices]: inputs H2D + packing, solve, results D2H */
int altro_b200_multi_last_timings(const altro_b200_multi* m, double* scatter_ms, double* solve_ms,
                                  double* gather_ms);

/* --- measurement helpers -------------------------------------------------------------
 * Algorithmic bytes of one batched backward pass over the materialised expansions
 * (8*[N*(n(n+m)+n^2+nm+m^2+n+m+mn+m)+n^2+n] per instance). */
size_t altro_b200_backward_pass_bytes(const altro_b200_solver* s);
/* number of kernels this library has launched on behalf of `s` since creation */
int64_t altro_b200_kernel_launches(const altro_b200_solver* s);

/* ------------------------------------------------------------------------------------
 * Per-iteration statistics  (replaces the vectors of SolverStats: cost, alpha, improvement_ratio,
 * gradient, cost_decrease, regularization, violations, max_penalty — altro/common/solver_stats.hpp:54-61,
 * filled by Log() in ilqr.hpp:442, 545-547, 578-584 and al_solver.hpp:298-299, 361-362)
 * ---------------------------------------------------------------------------------- */
/* Record one row per inner iteration for the first `instances` instances, at most `rows` rows each.
 * Call before solving.  A recording solver runs on the fused engine (same results). */
int altro_b200_solver_enable_history(altro_b200_solver* s, int instances, int rows);
#define ALTRO_B200_HISTORY_COLS 8 /* cost, alpha, z, gradient, cost_decrease, regularization, violations, max_penalty */
/* rows_out[r * 8 + c] = value column c holds in the SolverStats row that iteration r + 1 wrote
 * (carry-forward included); *nrows = iterations recorded (<= max_rows). */
int altro_b200_get_history_host(altro_b200_solver* s, int instance, double* rows_out, int max_rows,
                                int* nrows, void* stream);
/* latency probe: cycles per call of the per-knot device functions on one warp (cycles[32]) */
int altro_b200_microbench(altro_b200_solver* s, long long* cycles, int reps);
/* device bytes held by the solver */
size_t altro_b200_device_bytes(const altro_b200_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* ALTRO_B200_H_ */
