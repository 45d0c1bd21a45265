/*
 * altro_b200.h -- C ABI of the B200-native batched AL-iLQR solve path.
 *
 * This is the drop-in boundary for the hot path of bjack205/altro (reference paths relative to
 * /root/reference).  Plain pointers and sizes only; every call returns the reference's
 * ErrorCodes integer (src/altro/solver/exceptions.hpp:24-51) unless stated otherwise.
 *
 *   section A  tvlqr_*            : the reference's inner C-style API (src/tvlqr/tvlqr.h:15-33),
 *                                   same argument lists, now `extern "C"`, executed on the GPU.
 *   section B  altro_b200_tvlqr_* : batched variants over device-resident problem-fastest arrays.
 *   section C  altro_b200_*       : batched solver handle mirroring altro::ALTROSolver
 *                                   (src/altro/altro_solver.hpp:21-442); one handle = B problems.
 *
 * Host arrays are problem-major ("B separate solvers laid end to end"): [B][...]; matrices inside
 * a problem are column-major (altro_solver.hpp:185).  Device arrays are problem-fastest
 * (altro_b200/csrc/device_problem.h).  No CPU fallback exists: every compute entry point fails
 * with ALTRO_B200_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef ALTRO_B200_H
#define ALTRO_B200_H

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums shared with the reference (same integer values) -------------------------------- */
/* ErrorCodes, exceptions.hpp:24-51 */
enum {
  ALTRO_B200_NO_ERROR = 0,
  ALTRO_B200_STATE_DIM_UNKNOWN = 1,
  ALTRO_B200_INPUT_DIM_UNKNOWN = 2,
  ALTRO_B200_NEXT_STATE_DIM_UNKNOWN = 3,
  ALTRO_B200_DIMENSION_UNKNOWN = 4,
  ALTRO_B200_BAD_INDEX = 5,
  ALTRO_B200_DIMENSION_MISMATCH = 6,
  ALTRO_B200_SOLVER_NOT_INITIALIZED = 7,
  ALTRO_B200_SOLVER_ALREADY_INITIALIZED = 8,
  ALTRO_B200_NON_POSITIVE = 9,
  ALTRO_B200_TIMESTEP_NOT_POSITIVE = 10,
  ALTRO_B200_COST_FUN_NOT_SET = 11,
  ALTRO_B200_DYNAMICS_FUN_NOT_SET = 12,
  ALTRO_B200_INVALID_OPT_AT_TERMINAL = 13,
  ALTRO_B200_MAX_CONSTRAINTS_EXCEEDED = 14,
  ALTRO_B200_INVALID_CONSTRAINT_DIM = 15,
  ALTRO_B200_CHOLESKY_FAILED = 16,
  ALTRO_B200_OP_ONLY_VALID_AT_TERMINAL = 17,
  ALTRO_B200_INVALID_POINTER = 18,
  ALTRO_B200_BACKWARD_PASS_FAILED = 19,
  ALTRO_B200_LINESEARCH_FAILED = 20,
  ALTRO_B200_MERIT_GRADIENT_TOO_SMALL = 21,
  ALTRO_B200_INVALID_BOUND_CONSTRAINT = 22,
  ALTRO_B200_NON_POSITIVE_PENALTY = 23,
  ALTRO_B200_COST_NOT_QUADRATIC = 24,
  ALTRO_B200_FILE_ERROR = 25,
  /* extensions (not in the reference) */
  ALTRO_B200_ERR_NO_DEVICE = 100,   /* no CUDA device / CUDA runtime error                    */
  ALTRO_B200_ERR_UNSUPPORTED = 101, /* model/dimension combination not compiled in            */
};
/* SolveStatus, typedefs.hpp:19-27 */
enum { ALTRO_B200_SUCCESS = 0, ALTRO_B200_UNSOLVED = 1, ALTRO_B200_MAX_ITERATIONS = 2 };
/* ConstraintType, typedefs.hpp:53 */
enum {
  ALTRO_B200_EQUALITY = 0,
  ALTRO_B200_IDENTITY = 1,
  ALTRO_B200_INEQUALITY = 2,
  ALTRO_B200_SECOND_ORDER_CONE = 3,
};
/* index sentinels, typedefs.hpp:16-17 */
enum { ALTRO_B200_LAST_INDEX = -1, ALTRO_B200_ALL_INDICES = -2 };
/* device dynamics models replacing the std::function callbacks (typedefs.hpp:31-35) */
enum {
  ALTRO_B200_MODEL_LINEAR = 0,            /* KnotPointData::SetLinearDynamics                 */
  ALTRO_B200_MODEL_DOUBLE_INTEGRATOR = 1, /* test/test_utils.cpp:18-41, params[0] = dim       */
  ALTRO_B200_MODEL_PENDULUM = 2,          /* test/test_utils.cpp:43-82 + midpoint :84-132     */
  ALTRO_B200_MODEL_BICYCLE4 = 3,          /* test/test_utils.cpp:134-238, params = {L, lr}    */
  ALTRO_B200_MODEL_BICYCLE5 = 4,          /* [x,y,theta,delta,v], u=[a,delta_dot], {L, lr}    */
  ALTRO_B200_MODEL_CHAIN = 5,             /* coupled pendulum chain, params = {n, m}          */
};

/* ================================================================================ section A
 * Drop-in for src/tvlqr/tvlqr.h:15-33 (declared there without extern "C"; same arguments,
 * same return convention: TVLQR_SUCCESS = -1, or the knot index whose Cholesky failed,
 * tvlqr.cpp:162-164).  Host pointer tables, one pointer per knot; the caller owns all memory.
 * Dimensions must be uniform over the horizon (every reference call site is); otherwise -2.
 * The Q*_tmp scratch tables may be NULL (the GPU path keeps scratch in registers).  Qxx, Quu,
 * Qux, Qx, Qu -- outputs of the reference: the action-value expansion of every knot
 * (tvlqr.cpp:123-152) -- are written when all five tables are passed, and skipped when NULL. */
#ifndef TVLQR_SUCCESS
#define TVLQR_SUCCESS -1
#endif
typedef double lqr_float;

int tvlqr_TotalMemSize(const int *nx, const int *nu, int num_horizon, bool is_diag);

int tvlqr_BackwardPass(const int *nx, const int *nu, int num_horizon,
                       const lqr_float *const *A, const lqr_float *const *B, const lqr_float *const *f,
                       const lqr_float *const *Q, const lqr_float *const *R, const lqr_float *const *H,
                       const lqr_float *const *q, const lqr_float *const *r, lqr_float reg,
                       lqr_float **K, lqr_float **d,
                       lqr_float **P, lqr_float **p, lqr_float *delta_V,
                       lqr_float **Qxx, lqr_float **Quu, lqr_float **Qux,
                       lqr_float **Qx, lqr_float **Qu,
                       lqr_float **Qxx_tmp, lqr_float **Quu_tmp, lqr_float **Qux_tmp,
                       lqr_float **Qx_tmp, lqr_float **Qu_tmp,
                       bool linear_only_update, bool is_diag);

int tvlqr_ForwardPass(const int *nx, const int *nu, int num_horizon,
                      const lqr_float *const *A, const lqr_float *const *B, const lqr_float *const *f,
                      const lqr_float *const *K, const lqr_float *const *d,
                      const lqr_float *const *P, const lqr_float *const *p,
                      const lqr_float *x0, lqr_float **x, lqr_float **u, lqr_float **y);

/* ================================================================================ section B
 * Batched TVLQR over B independent LQ problems.  HOST arrays, problem-major:
 *   A [B][N][n*n]  B [B][N][n*m]  f [B][N][n]  Q [B][N+1][n*n or n]  R [B][N][m*m or m]
 *   H [B][N][m*n] (ignored when is_diag)  q [B][N+1][n]  r [B][N][m]
 * outputs K [B][N][m*n]  d [B][N][m]  P [B][N+1][n*n]  p [B][N+1][n]  delta_V [B][2]
 * status[B]: -1 or failing knot (same convention as section A).  Any output may be NULL.       */
int altro_b200_tvlqr_backward_batch(int batch, int n, int m, int num_horizon, const double *A,
                                    const double *B, const double *f, const double *Q,
                                    const double *R, const double *H, const double *q,
                                    const double *r, double reg, bool is_diag, double *K,
                                    double *d, double *P, double *p, double *delta_V,
                                    int *status);
/* x0 [B][n]; outputs x [B][N+1][n], u [B][N][m], y [B][N+1][n] (y may be NULL) */
int altro_b200_tvlqr_forward_batch(int batch, int n, int m, int num_horizon, const double *A,
                                   const double *B, const double *f, const double *K,
                                   const double *d, const double *P, const double *p,
                                   const double *x0, double *x, double *u, double *y);

/* Device-resident workspace behind the two calls above, for callers that run the sweep more than
 * once: upload the LQ data once, run backward()/forward() many times -- no allocation and no
 * synchronisation inside backward().  Compiled-in shapes (n, m): (2,1) (4,2) (4,4) (5,2) (6,2) (6,3)
 * (6,4); create returns NULL for any other shape (use the stateless calls, which fall back to a
 * run-time-dimension kernel) or without a CUDA device.  Arrays as in section B. */
typedef struct altro_b200_tvlqr_ws altro_b200_tvlqr_ws;
altro_b200_tvlqr_ws *altro_b200_tvlqr_ws_create(int batch, int n, int m, int num_horizon, bool is_diag,
                                                int device);
void altro_b200_tvlqr_ws_destroy(altro_b200_tvlqr_ws *w);
/* any array may be NULL (left as it is) */
int altro_b200_tvlqr_ws_upload(altro_b200_tvlqr_ws *w, const double *A, const double *B, const double *f,
                               const double *Q, const double *R, const double *H, const double *q,
                               const double *r);
int altro_b200_tvlqr_ws_backward(altro_b200_tvlqr_ws *w, double reg); /* asynchronous */
int altro_b200_tvlqr_ws_download(altro_b200_tvlqr_ws *w, double *K, double *d, double *P, double *p,
                                 double *delta_V, int *status);
/* gains / cost-to-go from the host instead of a previous backward() */
int altro_b200_tvlqr_ws_set_gains(altro_b200_tvlqr_ws *w, const double *K, const double *d, const double *P,
                                  const double *p);
int altro_b200_tvlqr_ws_forward(altro_b200_tvlqr_ws *w, const double *x0, double *x, double *u, double *y);
/* CUDA-event time of `reps` back-to-back backward() launches (after one warm-up), per launch */
int altro_b200_tvlqr_ws_time_backward(altro_b200_tvlqr_ws *w, double reg, int reps, float *ms_per_launch);
/* compulsory HBM bytes of one backward step of one problem (all input rows read, K d P p written) */
long altro_b200_tvlqr_ws_bytes_per_knot(const altro_b200_tvlqr_ws *w);

/* ================================================================================ section C */
typedef struct altro_b200_solver altro_b200_solver;

/* AltroOptions fields read by the solve path (solver_options.hpp:16-39) */
typedef struct {
  int iterations_max;
  double tol_primal_feasibility;
  double tol_stationarity;
  double tol_meritfun_gradient;
  double penalty_initial;
  double penalty_scaling;
  double penalty_max;
  int use_backtracking_linesearch;
  double linesearch_c1, linesearch_c2; /* CubicLineSearch::SetOptimalityTolerances */
} altro_b200_options;

int altro_b200_device_count(void);
const char *altro_b200_error_string(int code);          /* ErrorCodeToString, exceptions.cpp   */
void altro_b200_default_options(altro_b200_options *o); /* AltroOptions defaults               */

/* ALTROSolver(horizon_length) for `batch` independent problems on CUDA device `device`.
 * Returns NULL if no device is available (no CPU fallback). */
altro_b200_solver *altro_b200_create(int horizon_length, int batch, int device);
void altro_b200_destroy(altro_b200_solver *s);
int altro_b200_set_stream(altro_b200_solver *s, void *cuda_stream);

int altro_b200_set_dimension(altro_b200_solver *s, int num_states, int num_inputs); /* SetDimension */
int altro_b200_set_time_step(altro_b200_solver *s, float h);                        /* SetTimeStep  */
/* SetTimeStep(h, k_start, k_stop) (src/altro/altro_solver.cpp:49-63) for knots [k_start, k_stop) of
 * the horizon; index conventions of the reference (ALL_INDICES / LAST_INDEX, k_stop = 0: one knot) */
int altro_b200_set_time_step_range(altro_b200_solver *s, float h, int k_start, int k_stop);
/* SetExplicitDynamics with a device model id instead of callbacks */
int altro_b200_set_model(altro_b200_solver *s, int model_id, const double *params, int nparams);
/* KnotPointData::SetLinearDynamics for knots [k_start,k_stop): A n*n, B n*m, f n (f may be NULL);
 * shared by the whole batch; requires ALTRO_B200_MODEL_LINEAR */
int altro_b200_set_linear_dynamics(altro_b200_solver *s, const double *A, const double *B,
                                   const double *f, int k_start, int k_stop);

/* SetLQRCost (altro_solver.cpp:138-172) on knots [k_start,k_stop) (inclusive-terminal index
 * semantics of altro_solver.cpp:385-433).  Qd[n], Rd[m] shared; xref/uref are [n]/[m] when
 * per_problem == 0, else [B][n] / [B][m]. */
int altro_b200_set_lqr_cost(altro_b200_solver *s, const double *Qd, const double *Rd,
                            const double *xref, const double *uref, int per_problem, int k_start,
                            int k_stop);
/* Tracking form of SetLQRCost (test/bicycle_test.cpp:183-186): knot k of problem b tracks row
 * offsets[b] + k of the shared reference tables xtab [T][n], utab [T][m]; all knots 0..N. */
int altro_b200_set_lqr_cost_window(altro_b200_solver *s, const double *Qd, const double *Rd,
                                   const double *xtab, const double *utab, int T,
                                   const int *offsets);
/* SetDiagonalCost: q [n], r [m], c scalar per knot (per_problem == 0: one knot's worth applied to
 * every knot in range; per_problem == 1: [B][k_stop-k_start][n], [B][..][m], [B][..]) */
int altro_b200_set_diagonal_cost(altro_b200_solver *s, const double *Qd, const double *Rd,
                                 const double *q, const double *r, const double *c,
                                 int per_problem, int k_start, int k_stop);
/* SetQuadraticCost (altro_solver.cpp:118-136, knotpoint_data.cpp:64-85): dense Q [n*n], R [m*m],
 * H [m*n] (column-major, shared by the batch; H may be NULL = 0; R, r ignored on the terminal
 * knot); q, r, c as in altro_b200_set_diagonal_cost. */
int altro_b200_set_quadratic_cost(altro_b200_solver *s, const double *Q, const double *R,
                                  const double *H, const double *q, const double *r,
                                  const double *c, int per_problem, int k_start, int k_stop);
/* UpdateLinearCosts (altro_solver.cpp:266-281): q and/or r may be NULL */
int altro_b200_update_linear_costs(altro_b200_solver *s, const double *q, const double *r,
                                   const double *c, int per_problem, int k_start, int k_stop);
/* advance every problem's tracking window by `steps` rows (on-device UpdateLinearCosts for the
 * receding-horizon loop of test/bicycle_test.cpp:320-331) */
int altro_b200_advance_window(altro_b200_solver *s, int steps);
/* the reference's own MPC cost update on the moved window (test/bicycle_test.cpp:317-328):
 * UpdateLinearCosts(q, nullptr, c, k) for every knot with q_k = -(Qd .* xref[row]),
 * c_k = 1/2 xref' Qd xref (+ c_u for k < N); r_k is left as the original SetLQRCost set it. */
int altro_b200_advance_window_linear(altro_b200_solver *s, int steps, double c_u);
/* which of the two altro_b200_mpc_step applies: mode 1 (default) the reference's update above with
 * the frozen input-cost constant c_u = 1/2 u0' R u0 (test/bicycle_test.cpp:296), mode 0 the full
 * re-windowing of altro_b200_advance_window (q, r and c all follow the window). */
int altro_b200_set_mpc_cost_update(altro_b200_solver *s, int mode, double c_u);

/* SetConstraint with a built-in row family instead of callbacks:
 *   c_i = scale_i * [x;u][idx_i] + off_i   (idx_i = -1: c_i = off_i),   i < dim
 * off_b: optional per-problem offsets [B][dim] (NULL: shared `off`). */
int altro_b200_set_constraint(altro_b200_solver *s, int cone, int dim, const int *idx,
                              const double *scale, const double *off, const double *off_b,
                              int k_start, int k_stop);

/* SetConstraint, general affine rows: c = J [x;u] + e with a dense J [dim x (n+m)] (column-major,
 * the layout of the reference's constraint Jacobian, altro_solver.hpp:69-70) shared by the batch;
 * e [dim] shared, or e_b [B][dim] per problem (then e is ignored but must be non-NULL). */
int altro_b200_set_constraint_affine(altro_b200_solver *s, int cone, int dim, const double *J,
                                     const double *e, const double *e_b, int k_start, int k_stop);
/* SetConstraint, nonlinear family "keep-out disc" (INEQUALITY, dim 1):
 *   c = r^2 - ([x;u][idx_a] - cx)^2 - ([x;u][idx_b] - cy)^2 <= 0,  disc = {cx, cy, r},
 * disc_b: optional per-problem discs [B][3]. */
int altro_b200_set_constraint_disc(altro_b200_solver *s, int idx_a, int idx_b, const double *disc,
                                   const double *disc_b, int k_start, int k_stop);

int altro_b200_set_initial_state(altro_b200_solver *s, const double *x0, int per_problem);
int altro_b200_initialize(altro_b200_solver *s);
/* SetInput: layout 0: u[m] for every knot in range and every problem; 1: [k_stop-k_start][m]
 * shared by the batch; 2: [B][k_stop-k_start][m] */
int altro_b200_set_input(altro_b200_solver *s, const double *u, int layout, int k_start,
                         int k_stop);
int altro_b200_set_state(altro_b200_solver *s, const double *x, int layout, int k_start,
                         int k_stop);
int altro_b200_set_options(altro_b200_solver *s, const altro_b200_options *o);
/* reset duals (z = 0) and penalties (rho = 1) to their post-Initialize values */
int altro_b200_reset_duals(altro_b200_solver *s);
int altro_b200_shift_trajectory(altro_b200_solver *s); /* ShiftTrajectory, altro_solver.cpp:283 */
/* one receding-horizon (MPC) step entirely on the device (test/bicycle_test.cpp:302-337, whose
 * plant is the model): x0 <- x_[1] of the solved trajectory, ShiftTrajectory, the tracking window
 * moves one row with the cost update selected by altro_b200_set_mpc_cost_update; duals and
 * penalties carry over.  Call altro_b200_solve again afterwards. */
int altro_b200_mpc_step(altro_b200_solver *s);
/* restore the working inputs u_ to the last SetInput guess (device-to-device; lets a resident
 * batch be re-solved from the same starting point without a host round trip) */
int altro_b200_reset_trajectory(altro_b200_solver *s);

/* KnotPointData expansions at the working trajectory x_, u_ for every knot of every problem:
 * CalcDynamicsExpansion, CalcConstraints, CalcConstraintJacobians, CalcProjectedDuals,
 * CalcCostGradient (knotpoint_data.cpp:406-437, :473-487, :523-595) with the current duals and
 * penalty; read the results with altro_b200_get_field.  This is the entry point the reference's
 * knotpoint_data_test.cpp bodies map onto. */
int altro_b200_knot_eval(altro_b200_solver *s);
/* KnotPointData::SetPenalty for every constraint of every problem (rho > 0) */
int altro_b200_set_penalty(altro_b200_solver *s, double rho);
/* OpenLoopRollout (altro_solver.cpp:253): x_[k+1] = f(x_[k], u_[k]) from the initial state */
int altro_b200_open_loop_rollout(altro_b200_solver *s);
/* CalcCost (altro_solver.cpp:313): total cost incl. AL terms of the working trajectory, [B] */
int altro_b200_calc_cost(altro_b200_solver *s, double *cost);

/* Solve() for all B problems: launches the kernel sequence on the handle's stream and waits. */
int altro_b200_solve(altro_b200_solver *s);
int altro_b200_solve_async(altro_b200_solver *s); /* no wait; use altro_b200_synchronize */
int altro_b200_synchronize(altro_b200_solver *s);
/* 0 (default): two kernels per iLQR iteration (backward sweep; forward = whole line search +
 * criteria + AL update), enqueued back to back with no host decision in the loop -- the host only
 * looks at a stop counter two iterations late, so solve_async returns once the last iterations
 * are queued;
 * 1: one persistent kernel for the whole solve (thread per trajectory; the bit-exact twin the
 * tests compare mode 0 against) */
int altro_b200_set_solve_mode(altro_b200_solver *s, int mode);
/* Riccati sweep of the default mode: 0 one warp per group of 32 problems (thread = trajectory),
 * 1 the blocks of a problem spread by columns over the warps of a CTA (exchange through shared
 * memory; what lets n = 12 stay on chip), -1 (default) by block size: 1 for n > 6.  Bit-identical
 * results. */
int altro_b200_set_backward_mode(altro_b200_solver *s, int team);
/* pipelined sub-batches: the batch is cut into `nsplit` contiguous ranges (1..32; 0 = automatic),
 * each on its own stream (one host thread enqueues all) so the sweeps of one range overlap the
 * rollouts of another.  Results do not depend on nsplit. */
int altro_b200_set_pipeline_split(altro_b200_solver *s, int nsplit);
/* number of candidate step lengths rolled out concurrently per backtracking round (1..8,
 * default 6; before altro_b200_initialize).  1 reproduces the strictly sequential search. */
int altro_b200_set_speculation(altro_b200_solver *s, int nslots);
/* how many of the speculative candidates keep their trajectory in HBM (0..nslots-1, after
 * altro_b200_initialize; default all).  An accepted candidate that did not is rolled out once
 * more.  Results do not depend on it. */
int altro_b200_set_candidate_store(altro_b200_solver *s, int nstore);
/* per-phase instrumentation of the pipeline.  Entries: 0 init rollout, 1 prologue expansion,
 * 2 k_phase_backward (Riccati + alpha=0 scan), 3 k_phase_forward, and the forward kernel's time
 * split by its in-kernel sub-phase clocks: 4 rollout passes, 5 expansions, 6 d(phi) scan +
 * line-search machines, 7 criteria + AL update.  on=1 times every launch with CUDA events
 * (serialises the pipeline; not for throughput runs).  Arrays of 8 entries; *syncs = host waits
 * on the device (lagged stop-counter checks, at most one per iteration and sub-batch). */
int altro_b200_set_profiling(altro_b200_solver *s, int on);
int altro_b200_get_phase_stats(altro_b200_solver *s, double *ms, long *launches, double *units,
                               long *syncs);
/* accepted-step histogram of the line searches run so far (32 bins: 0 alpha0 accepted, 1..15
 * halving j accepted, 16 cubic-first probe, 17 zoom/other, 18 failed, 19 merit gradient too
 * small); reset != 0 clears it */
int altro_b200_get_linesearch_histogram(altro_b200_solver *s, long *hist32, int reset);
/* number of kernels this handle has launched since creation */
long altro_b200_kernel_launches(const altro_b200_solver *s);

/* getters: caller-owned host buffers, problem-major */
int altro_b200_get_states(altro_b200_solver *s, double *X);  /* [B][N+1][n]  GetState          */
int altro_b200_get_inputs(altro_b200_solver *s, double *U);  /* [B][N][m]    GetInput          */
int altro_b200_get_dual_dynamics(altro_b200_solver *s, double *Y); /* [B][N+1][n]             */
int altro_b200_get_feedback_gains(altro_b200_solver *s, double *K); /* [B][N][m*n]            */
int altro_b200_get_feedforward_gains(altro_b200_solver *s, double *d); /* [B][N][m]           */
/* any KnotPointData member by name (knotpoint_data.hpp:160-233).  Stored: "x" "u" "y" (x_, u_, y_)
 * "xbar" "ubar" (x, u) "A" "B" "lx" "lu" "K" "d" "P" "p" "q" "r" "c" "z" "z_est"; re-created on
 * demand at the working trajectory: "constraint_val" "z_proj" "lxx" "luu" "lux" "rho".  Constraint
 * members hold all slots of the knot one after the other (rows of slots that do not apply at a
 * knot are zero).  out: [B][N+1][rows] (column-major blocks); rows_out receives the rows per knot;
 * out may be NULL to query rows only. */
int altro_b200_get_field(altro_b200_solver *s, const char *name, double *out, int *rows_out);
/* GetDualGeneral / SetDualGeneric (altro_solver.hpp:359, :416): dual z of constraint `constraint`
 * (the order of the SetConstraint calls) at knot k, [B][dim]; set: z [dim] shared (per_problem 0)
 * or [B][dim]. */
int altro_b200_get_dual_general(altro_b200_solver *s, int constraint, int k, double *z);
int altro_b200_set_dual_general(altro_b200_solver *s, int constraint, int k, const double *z,
                                int per_problem);
int altro_b200_get_num_constraints(const altro_b200_solver *s);
int altro_b200_get_constraint_dim(const altro_b200_solver *s, int constraint); /* 0 if out of range */
int altro_b200_get_status(altro_b200_solver *s, int *status);       /* [B] SolveStatus        */
int altro_b200_get_iterations(altro_b200_solver *s, int *iters);    /* [B] GetIterations      */
int altro_b200_get_merit_evals(altro_b200_solver *s, int *evals);   /* [B]                    */
int altro_b200_get_final_objective(altro_b200_solver *s, double *phi);   /* [B]               */
int altro_b200_get_stationarity(altro_b200_solver *s, double *stat);     /* [B]               */
int altro_b200_get_primal_feasibility(altro_b200_solver *s, double *feas); /* [B]             */
int altro_b200_get_penalty(altro_b200_solver *s, double *rho);      /* [B]                    */
int altro_b200_get_horizon_length(const altro_b200_solver *s);
int altro_b200_get_batch(const altro_b200_solver *s);
int altro_b200_get_state_dim(const altro_b200_solver *s);
int altro_b200_get_input_dim(const altro_b200_solver *s);
/* bytes of HBM held by the handle */
long altro_b200_device_bytes(const altro_b200_solver *s);

/* ================================================================================ section D
 * Trajectory files with the reference's keys (host only, no device needed).  test/scotty.json
 * (read by ReadScottyTrajectory, test/test_utils.cpp:240-289): "N", "tf", "state_trajectory"
 * [knots][n], "input_trajectory" [knots][m]; test/scotty_mpc.json (written by
 * test/bicycle_test.cpp:344-359) adds "solve_iters" [steps] and "tracking_error" [steps].  A BATCH
 * file carries "batch": B and one more leading dimension on every array; a file without "batch" is
 * one problem, so the reference's own files load unchanged.                                     */
typedef struct altro_b200_traj_file altro_b200_traj_file;
/* parse `path`; NULL on failure with *err = ALTRO_B200_FILE_ERROR / DIMENSION_MISMATCH */
altro_b200_traj_file *altro_b200_traj_open(const char *path, int *err);
/* any output may be NULL.  batch = 1 for a single-problem file; steps = length of solve_iters /
 * tracking_error (0 when absent) */
int altro_b200_traj_dims(const altro_b200_traj_file *f, int *batch, int *N, float *tf, int *knots_x,
                         int *n, int *knots_u, int *m, int *steps);
/* X [batch][knots_x][n], U [batch][knots_u][m], solve_iters [batch][steps],
 * tracking_error [batch][steps]; any may be NULL */
int altro_b200_traj_read(const altro_b200_traj_file *f, double *X, double *U, int *solve_iters,
                         double *tracking_error);
void altro_b200_traj_close(altro_b200_traj_file *f);
/* batch = 0 writes the reference's single-problem layout (no "batch" key, 2-D arrays);
 * solve_iters / tracking_error may be NULL */
int altro_b200_traj_write(const char *path, int batch, int N, float tf, int knots_x, int n,
                          const double *X, int knots_u, int m, const double *U, int steps,
                          const int *solve_iters, const double *tracking_error);

#ifdef __cplusplus
}
#endif
#endif
