// device_problem.h -- the batched problem as the kernels see it (plain pointers into HBM).
//
// HBM layout (DESIGN.md "Data layout"): structure-of-arrays, PROBLEM INDEX FASTEST.  A field with
// E doubles per knot point stores element e of knot k of problem b at
//     field[(k*E + e) * Bp + b]            (Bp = batch padded to a multiple of 32)
// with blocks column-major inside a knot, i.e. the reference's KnotPointData members
// (knotpoint_data.hpp:160-233) transposed so that the 32 lanes of a warp -- 32 consecutive
// problems -- read one contiguous 256-byte row per matrix element.
#pragma once

namespace altro_b200 {

constexpr int kMaxCon = 4;       // constraint slots per problem on the device
constexpr int kMaxConDim = 16;   // rows per constraint
constexpr int kMaxSocDim = 6;    // rows of a second-order-cone constraint

// ConstraintType, typedefs.hpp:53
enum Cone { CONE_EQUALITY = 0, CONE_IDENTITY = 1, CONE_INEQUALITY = 2, CONE_SOC = 3 };

// SolveStatus, typedefs.hpp:19-27
enum DevSolveStatus { SOLVE_SUCCESS = 0, SOLVE_UNSOLVED = 1, SOLVE_MAX_ITERATIONS = 2 };

// One constraint slot: rows c_i = scale_i * [x;u][idx_i] + off_i on knots [k_start, k_stop).
// (idx_i = -1: constant row.)  This "selector-affine" family covers every constraint the
// reference's end-to-end tests use: goal, control box, control-norm SOC, steering bound.
struct ConSlot {
  int k_start, k_stop;
  int cone;
  int dim;
  int row0;             // first row of this slot in the packed per-knot dual arrays
  int off_per_problem;  // 1: offsets come from off_b (per problem) instead of off[]
  int idx[kMaxConDim];
  double scale[kMaxConDim];
  double off[kMaxConDim];
  const double* off_b;  // [dim][Bp]
};

struct ConTable {
  int ncon;
  int rows;  // total rows over all slots (stride of the per-knot dual arrays)
  ConSlot slot[kMaxCon];
};

// AltroOptions fields the path reads (solver_options.hpp:16-39) + line-search tolerances
struct DevOptions {
  int iterations_max;
  double tol_primal_feasibility;
  double tol_stationarity;
  double tol_meritfun_gradient;
  double penalty_initial;
  double penalty_scaling;
  double penalty_max;
  int use_backtracking_linesearch;
  double ls_c1, ls_c2;
};

struct DeviceProblem {
  int N, B;
  long Bp;  // padded batch = stride between consecutive elements
  float h;
  double model_params[8];
  const double* lin;  // MODEL_LINEAR: per knot [A (n*n) | B (n*m) | f (n)], shared by the batch

  // cost (KnotPointData Q_, R_, q_, r_, c_): diagonal weights shared per knot, linear terms per problem
  const double* Qd;  // [(N+1)*n]
  const double* Rd;  // [N*m]
  const double* q;   // [(N+1)*n][Bp]
  const double* r;   // [N*m][Bp]
  const double* c;   // [(N+1)][Bp]

  const double* x0;  // [n][Bp]            (SolverImpl::initial_state_)
  double *xbar, *ubar;      // accepted trajectory  (KnotPointData x, u)
  double *x, *u, *y;        // working trajectory   (x_, u_, y_)
  double *A, *Bm;           // dynamics expansion   (A_, B_)
  double *lx, *lu;          // cost gradient        (lx_, lu_)
  double *K, *d, *P, *p;    // gains / cost-to-go   (K_, d_, P_, p_)

  const ConTable* con;      // device pointer; ncon == 0 when unconstrained
  double *z, *zest;         // duals z_ and estimates z_est_: [(N+1)*rows][Bp]
  double* rho;              // penalty rho_ (uniform over knots and constraints): [Bp]

  // per-problem results
  int *status, *iters, *merit_evals, *ls_fail;
  double *phi, *stat, *feas;

  DevOptions opts;
};

}  // namespace altro_b200
