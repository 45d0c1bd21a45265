// device_problem.h -- the batched problem as the kernels see it (plain pointers into HBM).
//
// HBM layout (DESIGN.md "Data layout"): KNOT RECORDS.  The batch is cut into groups of 32
// consecutive problems (one warp); everything a group owns at knot k -- the reference's
// KnotPointData members (knotpoint_data.hpp:160-233), each block column-major -- is ONE contiguous
// record of `rows` 256-byte lines, line = one block element for the 32 problems of the group:
//     element e of field F at knot k of problem b:
//         F[(b / 32) * GS + k * R + e * 32 + (b % 32)],   R = rows * 32,  GS = (N + 1) * R
// (F = record buffer + 32 * the field's first row, so all fields share R and GS).  Row order:
//     xbar ubar | q r c | K d | x u | A B lx lu | y | P p | u_init
// chosen so that what each sequential sweep reads per knot is one contiguous multi-KB range: the
// rollout [xbar..d], the phi0 scan [q..B], the d(phi) scan [K..lu] minus [x u], the Riccati sweep
// [A..lu], the residual kernel [x..y].  A warp therefore streams whole records with one TMA bulk
// copy (cp.async.bulk) per knot into a shared-memory ring instead of E separate 256-byte pieces
// Bp*8 bytes apart: measured 6.8 TB/s vs 3.2 TB/s at the 3.5 warps/SM of B = 16384
// (profiles/r01_microbench_layout.txt).
#pragma once
#include "linesearch.cuh"

namespace altro_b200 {

constexpr int kMaxHalvings = 26;  // SimpleBacktracking tries at most max_iters - 1 = 24 halvings
constexpr int kMaxCon = 4;       // constraint slots per problem on the device
constexpr int kMaxConDim = 16;   // rows per constraint
constexpr int kMaxSocDim = 6;    // rows of a second-order-cone constraint

// ConstraintType, typedefs.hpp:53
enum Cone { CONE_EQUALITY = 0, CONE_IDENTITY = 1, CONE_INEQUALITY = 2, CONE_SOC = 3 };

// SolveStatus, typedefs.hpp:19-27
enum DevSolveStatus { SOLVE_SUCCESS = 0, SOLVE_UNSOLVED = 1, SOLVE_MAX_ITERATIONS = 2 };

// Constraint FAMILIES: what replaces the reference's constraint callbacks c(x,u), dc/d[x;u]
// (typedefs.hpp:48-52) on the device, the way ModelId replaces the dynamics callbacks.
enum ConFamily {
  // rows c_i = scale_i * [x;u][idx_i] + off_i  (idx_i = -1: constant row): one variable per row.
  // Covers every constraint of the reference's end-to-end tests (goal, control box, control-norm
  // SOC, steering bound) and has a diagonal fast path for the linear cones (TrajSolver CON = 1).
  CON_FAMILY_SELECTOR = 0,
  // general affine rows c = J [x;u] + e with a dense dim x (n+m) Jacobian shared by the batch
  // (column-major, in HBM) and e from off[] or per problem from off_b
  CON_FAMILY_AFFINE = 1,
  // nonlinear: keep-out disc, c = r^2 - (s_a - cx)^2 - (s_b - cy)^2 (dim 1, INEQUALITY: c <= 0)
  // with s_a = [x;u][idx[0]], s_b = [x;u][idx[1]], (cx, cy, r) = off[0..2] or per problem off_b
  CON_FAMILY_DISC = 2,
};

// One constraint slot on knots [k_start, k_stop).
struct ConSlot {
  int k_start, k_stop;
  int cone;
  int dim;
  int row0;             // first row of this slot in the packed per-knot dual arrays
  int off_per_problem;  // 1: offsets come from off_b (per problem) instead of off[]
  int family;           // ConFamily
  int idx[kMaxConDim];
  double scale[kMaxConDim];
  double off[kMaxConDim];
  const double* off_b;  // [G][nparam][32]  (nparam = dim, or 3 for CON_FAMILY_DISC)
  const double* Jd;     // CON_FAMILY_AFFINE: [dim x (n+m)] column-major, device memory
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
  long Bp;  // batch padded to a multiple of 32 (per-problem scalar arrays)
  int G;    // groups this launch sequence covers (a sub-batch of the handle when pipelined)
  int g0;   // first group of the sub-batch
  int Gtot; // groups of the whole handle = Bp / 32
  long R, GS;    // main record stream: knot stride, group stride (doubles)
  long Rz, GSz;  // dual record stream [group][knot][z rows | z_est rows][32]
  int zrows;     // constraint rows per knot (0 when unconstrained)
  long Rs;       // candidate-slot record stream [slot][group][knot][x rows | u rows][32]
  float h;
  const float* hk;  // [N] per-knot time steps, or null when every knot uses h
  double model_params[8];
  const double* lin;  // MODEL_LINEAR: per knot [A (n*n) | B (n*m) | f (n)], shared by the batch

  // cost (KnotPointData Q_, R_, q_, r_, c_): diagonal weights shared per knot, linear terms per problem
  const double* Qd;  // [(N+1)*n]
  const double* Rd;  // [N*m]
  // dense quadratic cost (KnotPointData::SetQuadraticCost, knotpoint_data.cpp:64-85), shared per
  // knot, column-major; null = diagonal.  Only the general instantiation (CON = 2) reads them.
  const double* Qf;  // [(N+1)][n*n]
  const double* Rf;  // [N][m*m]
  const double* Hf;  // [N][m*n]
  const double* q;   // record field, n rows
  const double* r;   // record field, m rows
  const double* c;   // record field, 1 row

  const double* x0;  // [G][n][32]         (SolverImpl::initial_state_)
  double *xbar, *ubar;      // accepted trajectory  (KnotPointData x, u)
  double *x, *u, *y;        // working trajectory   (x_, u_, y_)
  double *A, *Bm;           // dynamics expansion   (A_, B_)
  double *lx, *lu;          // cost gradient        (lx_, lu_)
  double *K, *d, *P, *p;    // gains / cost-to-go   (K_, d_, P_, p_)

  // constraint slots, by value: the struct is a __grid_constant__ kernel parameter, so the
  // per-knot loops over slots and rows read them through the constant cache instead of HBM/L2
  ConTable contab;
  double *z, *zest;         // duals z_ and estimates z_est_ (dual record stream)
  double* rho;              // penalty rho_ (uniform over knots and constraints): [Bp]

  // per-problem results
  int *status, *iters, *merit_evals, *ls_fail;
  double *phi, *stat, *feas;

  // ---- phase-kernel pipeline (solver_phases.cuh): per-trajectory scalar state
  LsMachine* ls;        // [B] line-search state machines
  double *alpha_eval;   // [Bp] step length of the pending merit evaluation
  double *alpha_bt;     // [Bp] first backtracking step after alpha_eval (speculative rounds)
  double *phi_eval;     // [Bp] merit value produced by the last rollout
  double *phi0, *dphi0; // [Bp]
  int* flags;           // [Bp] bit mask, see TrajFlags
  int* iter_count;      // [Bp] iLQR iterations done so far
  // [FS_COUNT + 1] profile mode only (else null): nanoseconds per sub-phase of k_phase_forward
  unsigned long long* prof;
  // [32] accepted-step histogram of the line search: 0 alpha0 accepted, 1..15 halving j accepted,
  // 16 cubic-first probe accepted, 17 zoom/other, 18 failed, 19 merit gradient too small
  unsigned long long* ls_hist;
  // speculative line-search slots: candidate step lengths of one trajectory are rolled out
  // concurrently, each into its own copy of the working trajectory
  int nslots;           // candidates per speculative round (>= 1): slot 0 = the requested step
  int nstore;           // speculative slots 1..nstore also keep their trajectory (slot buffers)
  // 1: a follower warp of k_phase_forward does the derivative half of a merit evaluation behind the
  // rollout warp (no separate expansion / d(phi) scan, no re-read of x, u, [J], lx, lu); 0: separate
  // knot-parallel expansion + staged scan
  int follow_deriv;
  // 1: the linear cost terms q_k, r_k, c_k are the same for every k < N (one SetLQRCost call over the
  // stage knots with a goal-type reference): the sweeps read them once from knot 0 instead of
  // streaming [q r c] with every knot
  int qrc_uniform;
  // 1: the halvings of a backtracking search are rolled out speculatively from the first round on
  // (next to alpha0); 0: from the second round on, next to the cubic-first probe -- the same number
  // of rounds, but no speculative work for the searches that accept alpha0
  int prof_tid;  // profile mode: the thread whose clocks are recorded (0: rollout warp, 32: follower, 64..: speculating)
  int spec_round1;
  double *xs, *us;      // slot record stream; us = xs + n * 32
  double* phi_s;        // [kMaxHalvings + 1][Bp] merit value of halving j (alpha0 * 2^-j), j >= 1
  int* spec_base;       // [Bp] halving index rolled out by slot 1 of the pending / last round
  int* spec_known;      // [Bp] halvings 1..spec_known have their merit value in phi_s (this iteration)
  int* sel;             // [Bp] slot holding the working trajectory (-1: the main x, u arrays)
  unsigned long long *stat_acc, *feas_acc;  // [Bp] max-reductions over knots (bit patterns of doubles >= 0)

  DevOptions opts;
};

enum TrajFlags {
  TF_NEED_EVAL = 1,         // the line search asked for a merit evaluation at alpha_eval
  TF_WANT_DERIV = 2,        // ... with derivative
  TF_REFRESH_DYN = 4,       // backtracking accepted alpha != 1: recompute A,B,lx,lu (solver.cpp:256-262)
  TF_REFRESH_GRAD = 8,      // duals/penalty changed: recompute projected duals and lx,lu (:483-486)
  TF_ACTIVE = 16,           // still iterating
  TF_LS_FAILED = 32,
  TF_CONVERGED = 64,
  TF_SPECULATE = 128,       // the pending round also rolls out the halvings alpha_bt * 2^-j (merit only)
  TF_REROLL = 256,          // accepted step came from a merit-only candidate: roll it out again, storing
};

}  // namespace altro_b200
