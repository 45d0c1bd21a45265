/*
 * altro_oracle.h -- CPU ORACLE for the AL-iLQR solve path of bjack205/altro.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (altro_b200/) never links, imports or calls anything in oracle/.
 *
 * It is a plain-C restatement (no Eigen) of the reference algorithm.  Every function
 * cites the reference file:line it follows (paths relative to /root/reference).
 * Parity status: PINNED -- checked against the reference's own golden vectors
 * (tests/test_oracle_*.py) and, for the line search,
 * against the reference's own linesearch.cpp + cubicspline.c compiled into
 * oracle/_ref/ (oracle/Makefile target `ref`).
 */
#ifndef ALTRO_ORACLE_H
#define ALTRO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_CON 8      /* constraints per knot point handled by the oracle      */
#define ORACLE_MAX_CON_DIM 16 /* rows per built-in (selector-affine) constraint         */

/* src/altro/solver/typedefs.hpp:53 */
enum { ORACLE_EQUALITY = 0, ORACLE_IDENTITY = 1, ORACLE_INEQUALITY = 2, ORACLE_SOC = 3 };
/* src/altro/solver/typedefs.hpp:19-27 */
enum {
  ORACLE_STATUS_SUCCESS = 0,
  ORACLE_STATUS_UNSOLVED = 1,
  ORACLE_STATUS_MAX_ITERATIONS = 2,
};
/* subset of src/altro/solver/exceptions.hpp:24-51 used on the path (same values) */
enum {
  ORACLE_NOERROR = 0,
  ORACLE_BACKWARD_PASS_FAILED = 19,
  ORACLE_LINESEARCH_FAILED = 20,
  ORACLE_MERIT_GRAD_TOO_SMALL = 21,
};
/* built-in dynamics models (test/test_utils.cpp + the two models SURVEY 8d defines) */
enum {
  ORACLE_MODEL_LINEAR = 0,           /* KnotPointData::SetLinearDynamics                */
  ORACLE_MODEL_DOUBLE_INTEGRATOR = 1,/* test_utils.cpp:18-41, params[0]=dim             */
  ORACLE_MODEL_PENDULUM = 2,         /* test_utils.cpp:43-82 + midpoint :84-132         */
  ORACLE_MODEL_BICYCLE4 = 3,         /* test_utils.cpp:134-238 CoG frame + midpoint     */
  ORACLE_MODEL_BICYCLE5 = 4,         /* [x,y,th,delta,v], u=[a,ddelta]    */
  ORACLE_MODEL_CHAIN = 5,            /* coupled pendulum chain            */
  ORACLE_MODEL_CALLBACK = 6,
};
enum { ORACLE_COST_QUADRATIC = 1, ORACLE_COST_DIAGONAL = 2 };

typedef void (*oracle_dyn_fn)(void *ud, double *xnext, const double *x, const double *u, float h);
typedef void (*oracle_con_fn)(void *ud, double *c, const double *x, const double *u);
typedef void (*oracle_merit_fn)(void *ctx, double alpha, double *phi, double *dphi);

/* src/altro/solver/solver_options.hpp:16-39 (fields the solve path reads) */
typedef struct {
  int iterations_max;
  double tol_primal_feasibility;
  double tol_stationarity;
  double tol_meritfun_gradient;
  double penalty_initial;
  double penalty_scaling;
  double penalty_max;
  int use_backtracking_linesearch;
  double ls_c1, ls_c2; /* CubicLineSearch::SetOptimalityTolerances (linesearch.hpp:55-56) */
} oracle_options;

void oracle_default_options(oracle_options *o);

/* ---------------------------------------------------------------- cubic spline + line search
 * restatement of src/linesearch/cubicspline.c and src/linesearch/linesearch.cpp            */
typedef struct {
  double x0, a, b, c, d;
} oracle_spline;
int oracle_spline_from2points(oracle_spline *p, double x1, double y1, double d1, double x2,
                              double y2, double d2);
double oracle_spline_argmin(const oracle_spline *p, int *err);

typedef struct {
  int max_iters;
  double alpha_max, beta_increase, beta_decrease, min_interval_size;
  int try_cubic_first, use_backtracking_linesearch;
  double c1, c2;
  int return_code, n_iters;
  double phi0, phi, phi_lo, phi_hi, dphi0, dphi, dphi_lo, dphi_hi;
  int sufficient_decrease, curvature;
} oracle_linesearch;
void oracle_ls_init(oracle_linesearch *ls);
double oracle_ls_run(oracle_linesearch *ls, oracle_merit_fn f, void *ctx, double alpha0,
                     double phi0, double dphi0);

/* ---------------------------------------------------------------- cones (cones.cpp) */
int oracle_dual_cone(int cone);
void oracle_conic_projection(int cone, int dim, const double *x, double *px);
void oracle_conic_projection_jacobian(int cone, int dim, const double *x, double *jac);
void oracle_conic_projection_hessian(int cone, int dim, const double *x, const double *b,
                                     double *hess);

/* ---------------------------------------------------------------- tvlqr (tvlqr.cpp) */
int oracle_tvlqr_total_mem_size(const int *nx, const int *nu, int num_horizon, int is_diag);
int oracle_tvlqr_backward_pass(const int *nx, const int *nu, int num_horizon,
                               const double *const *A, const double *const *B,
                               const double *const *f, const double *const *Q,
                               const double *const *R, const double *const *H,
                               const double *const *q, const double *const *r, double reg,
                               double **K, double **d, double **P, double **p, double *delta_V,
                               double **Qxx, double **Quu, double **Qux, double **Qx, double **Qu,
                               double **Qxx_tmp, double **Quu_tmp, double **Qux_tmp,
                               double **Qx_tmp, double **Qu_tmp, int linear_only_update,
                               int is_diag);
int oracle_tvlqr_forward_pass(const int *nx, const int *nu, int num_horizon,
                              const double *const *A, const double *const *B,
                              const double *const *f, const double *const *K,
                              const double *const *d, const double *const *P,
                              const double *const *p, const double *x0, double **x, double **u,
                              double **y);

/* ---------------------------------------------------------------- models */
void oracle_model_dims(int model_id, const double *params, int *n, int *m);
void oracle_model_dynamics(int model_id, const double *params, double *xn, const double *x,
                           const double *u, float h);
void oracle_model_jacobian(int model_id, const double *params, double *jac, const double *x,
                           const double *u, float h);
void oracle_model_continuous(int model_id, const double *params, double *xdot, const double *x,
                             const double *u);
void oracle_model_continuous_jacobian(int model_id, const double *params, double *jac,
                                      const double *x, const double *u);

/* ---------------------------------------------------------------- solver (solver.cpp) */
typedef struct oracle_solver oracle_solver;

oracle_solver *oracle_create(int N, int n, int m);
void oracle_destroy(oracle_solver *s);
void oracle_set_options(oracle_solver *s, const oracle_options *o);
void oracle_set_time_step(oracle_solver *s, float h);
void oracle_set_time_step_range(oracle_solver *s, float h, int k_start, int k_stop);
void oracle_set_model(oracle_solver *s, int model_id, const double *params, int nparams);
void oracle_set_dynamics_callback(oracle_solver *s, oracle_dyn_fn dyn, oracle_dyn_fn jac,
                                  void *ud);
void oracle_set_linear_dynamics(oracle_solver *s, int k, const double *A, const double *B,
                                const double *f);
void oracle_set_diagonal_cost(oracle_solver *s, int k, const double *Qd, const double *Rd,
                              const double *q, const double *r, double c);
void oracle_set_quadratic_cost(oracle_solver *s, int k, const double *Q, const double *R,
                               const double *H, const double *q, const double *r, double c);
void oracle_set_lqr_cost(oracle_solver *s, int k, const double *Qd, const double *Rd,
                         const double *xref, const double *uref);
int oracle_add_constraint_selector(oracle_solver *s, int k, int cone, int dim, const int *idx,
                                   const double *scale, const double *off);
int oracle_add_constraint_callback(oracle_solver *s, int k, int cone, int dim, oracle_con_fn con,
                                   oracle_con_fn jac, void *ud);
void oracle_set_initial_state(oracle_solver *s, const double *x0);
int oracle_initialize(oracle_solver *s);
void oracle_set_state(oracle_solver *s, int k, const double *x);
void oracle_set_input(oracle_solver *s, int k, const double *u);
void oracle_set_dual(oracle_solver *s, int k, int j, const double *z);
void oracle_set_penalty(oracle_solver *s, double rho);
void oracle_update_linear_costs(oracle_solver *s, int k, const double *q, const double *r,
                                double c);
void oracle_shift_trajectory(oracle_solver *s);

int oracle_solve(oracle_solver *s);
int oracle_get_status(const oracle_solver *s);
int oracle_get_iterations(const oracle_solver *s);
long oracle_get_merit_evals(const oracle_solver *s);
double oracle_get_final_phi(const oracle_solver *s);

/* stage-level entry points (reference tests drive these directly) */
void oracle_open_loop_rollout(oracle_solver *s);
void oracle_linear_rollout(oracle_solver *s);
void oracle_copy_trajectory(oracle_solver *s);
double oracle_calc_cost(oracle_solver *s);
void oracle_calc_cost_gradient(oracle_solver *s);
void oracle_calc_expansions(oracle_solver *s);
int oracle_backward_pass(oracle_solver *s);
void oracle_merit_function(oracle_solver *s, double alpha, double *phi, double *dphi);
int oracle_forward_pass(oracle_solver *s, double *alpha);
double oracle_stationarity(oracle_solver *s);
double oracle_feasibility(oracle_solver *s);
void oracle_dual_update(oracle_solver *s);
void oracle_penalty_update(oracle_solver *s);
int oracle_ls_iters(const oracle_solver *s);
void oracle_get_merit_values(const oracle_solver *s, double *phi0, double *dphi0, double *phi,
                             double *dphi);

/* knot-level ops (KnotPointData::Calc*), op names as in the reference */
enum {
  ORACLE_OP_CALC_CONSTRAINTS = 0,
  ORACLE_OP_CALC_CONSTRAINT_JACOBIANS,
  ORACLE_OP_CALC_PROJECTED_DUALS,
  ORACLE_OP_CALC_CONIC_JACOBIANS,
  ORACLE_OP_CALC_CONIC_HESSIANS,
  ORACLE_OP_CALC_COST_GRADIENT,
  ORACLE_OP_CALC_COST_HESSIAN,
  ORACLE_OP_CALC_DYNAMICS_EXPANSION,
  ORACLE_OP_CALC_CONSTRAINT_COST_GRADIENTS,
  ORACLE_OP_CALC_CONSTRAINT_COST_HESSIANS,
  ORACLE_OP_CALC_ORIGINAL_COST_GRADIENT,
  ORACLE_OP_CALC_ORIGINAL_COST_HESSIAN,
};
void oracle_knot_op(oracle_solver *s, int k, int op);
double oracle_knot_calc_cost(oracle_solver *s, int k);
double oracle_knot_calc_constraint_costs(oracle_solver *s, int k);
double oracle_knot_calc_violations(oracle_solver *s, int k);

/* field access by KnotPointData member name ("x_", "K_", "lxx_", "constraint_hess_0", ...).
 * Returns the number of doubles copied, or -1 if the name is unknown. */
int oracle_get_field(const oracle_solver *s, int k, const char *name, double *out);
int oracle_set_field(oracle_solver *s, int k, const char *name, const double *in);

/* ---------------------------------------------------------------- batch driver
 * (CPU baseline + GPU parity oracle).  Host layout is problem-major ([B][...]).        */
typedef struct {
  int k_start, k_stop; /* half-open knot range                                             */
  int cone, dim;
  int idx[ORACLE_MAX_CON_DIM];      /* index into [x;u], or -1                             */
  double scale[ORACLE_MAX_CON_DIM]; /* c_i = scale_i * v[idx_i] + off_i                    */
  double off[ORACLE_MAX_CON_DIM];
  const double *off_b;              /* optional per-problem offsets [B][dim] (else NULL)   */
} oracle_con_spec;

typedef struct {
  int N, n, m, B;
  float h;
  int model_id;
  double model_params[8];
  const double *Qd; /* [(N+1)*n] shared diagonal state weights per knot                    */
  const double *Rd; /* [N*m]                                                                */
  int ref_mode;     /* 0: shared q,r,c  1: per-problem q,r,c  2: per-problem goal (xref,uref)
                       3: window into a shared reference table                              */
  const double *q, *r, *c;
  const double *xref, *uref;
  const int *offsets;
  int T;
  const double *x0; /* [B*n] */
  const double *U0; /* [N*m] or [B*N*m] */
  int U0_per_problem;
  int ncon;
  oracle_con_spec con[ORACLE_MAX_CON];
  oracle_options opts;
} oracle_batch_spec;

typedef struct {
  double *X;      /* [B][(N+1)][n] */
  double *U;      /* [B][N][m]     */
  double *Y;      /* [B][(N+1)][n] or NULL */
  int *status;    /* [B] */
  int *iters;     /* [B] */
  long *merit_evals; /* [B] or NULL */
  double *cost;   /* [B] final merit value phi */
  double *stat;   /* [B] final stationarity or NULL */
  double *feas;   /* [B] final feasibility or NULL */
} oracle_batch_result;

/* builds the single-problem solver for problem b of the spec (caller destroys) */
oracle_solver *oracle_batch_make_solver(const oracle_batch_spec *spec, int b);
/* solves problems [b0,b1) with nthreads OpenMP threads; returns wall seconds */
double oracle_batch_solve(const oracle_batch_spec *spec, int b0, int b1, int nthreads,
                          oracle_batch_result *res);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
