// altro/altro_solver.hpp -- drop-in C++ facade for the reference's public API
// (src/altro/altro_solver.hpp:21-442, solver/typedefs.hpp, solver/solver_options.hpp,
// solver/exceptions.hpp), implemented over the C ABI of include/altro_b200.h.
//
// Same class name, method names, argument order and defaults, enum values, copy-in /
// caller-owned-out pointer rules and print-or-throw error macro as the reference.  Differences,
// all additive:
//   * ALTROSolver(horizon_length, batch = 1): one object can hold a batch of B problems;
//     per-problem data are passed problem-major ([B][...]) through the *Batch methods, and every
//     reference method keeps its single-problem meaning (applied to all problems of the batch).
//   * dynamics and constraints run on the device, so the std::function arguments must wrap one of
//     the device models / constraint families of namespace altro::b200 (DeviceDynamics,
//     DeviceConstraint).  A plain host lambda is rejected with DynamicsFunNotSet /
//     InvalidConstraintDim-style errors instead of silently running on the CPU.
//   * methods the reference declares but never defines (GetStatus, GetFeedbackGain, ...) work.
#pragma once

#include <cstdio>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace altro {

using a_float = double;  // typedefs.hpp:12

constexpr int LastIndex = -1;   // typedefs.hpp:16
constexpr int AllIndices = -2;  // typedefs.hpp:17

enum class SolveStatus {  // typedefs.hpp:19-27
  Success,
  Unsolved,
  MaxIterations,
  MaxObjectiveExceeded,
  StateOutOfBounds,
  InputOutOfBounds,
  MeritFunGradientTooSmall,
};

enum class ErrorCodes {  // exceptions.hpp:24-51
  NoError,
  StateDimUnknown,
  InputDimUnknown,
  NextStateDimUnknown,
  DimensionUnknown,
  BadIndex,
  DimensionMismatch,
  SolverNotInitialized,
  SolverAlreadyInitialized,
  NonPositive,
  TimestepNotPositive,
  CostFunNotSet,
  DynamicsFunNotSet,
  InvalidOptAtTerminalKnotPoint,
  MaxConstraintsExceeded,
  InvalidConstraintDim,
  CholeskyFailed,
  OpOnlyValidAtTerminalKnotPoint,
  InvalidPointer,
  BackwardPassFailed,
  LineSearchFailed,
  MeritFunctionGradientTooSmall,
  InvalidBoundConstraint,
  NonPositivePenalty,
  CostNotQuadratic,
  FileError,
  // extensions
  NoDevice = 100,
  Unsupported = 101,
};

const char* ErrorCodeToString(ErrorCodes err);
void PrintErrorCode(ErrorCodes err);

class AltroErrorException : public std::runtime_error {  // exceptions.hpp:57-68
 public:
  AltroErrorException(std::string msg, ErrorCodes code)
      : std::runtime_error(msg.c_str()), code_(code) {}
  virtual ErrorCodes Errno() { return code_; }
  virtual ~AltroErrorException() {}

 private:
  ErrorCodes code_;
};

// exceptions.hpp:13-20: prints in red and evaluates to `code`, or throws
#undef ALTRO_THROW
#ifdef ALTRO_ENABLE_RUNTIME_EXCEPTIONS
#define ALTRO_THROW(msg, code) (throw(::altro::AltroErrorException((msg), code)), code)
#else
#define ALTRO_THROW(msg, code)                                                                   \
  (std::fprintf(stderr, "\033[31mALTRO ERROR Code %d: %s %s:%d\n  Message: %s\033[0m\n",         \
                static_cast<int>(code), ::altro::ErrorCodeToString(code), __FILE__, __LINE__,    \
                std::string(msg).c_str()),                                                       \
   code)
#endif

enum class ConstraintType { EQUALITY, IDENTITY, INEQUALITY, SECOND_ORDER_CONE };  // typedefs.hpp:53

enum class Verbosity { Silent, Outer, Inner, LineSearch };  // solver_options.hpp:14

struct AltroOptions {  // solver_options.hpp:16-39 (same fields, same defaults)
  AltroOptions() = default;
  int iterations_max = 200;
  double tol_cost = 1e-4;
  double tol_cost_intermediate = 1e-4;
  double tol_primal_feasibility = 1e-4;
  double tol_stationarity = 1e-4;
  double tol_meritfun_gradient = 1e-8;
  double max_state_value = std::numeric_limits<double>::infinity();
  double max_input_value = std::numeric_limits<double>::infinity();
  double penalty_initial = 1.0;
  double penalty_scaling = 10.0;
  double penalty_max = 1e8;
  Verbosity verbose = Verbosity::Silent;
  double max_solve_time = std::numeric_limits<a_float>::infinity();
  double use_backtracking_linesearch = false;  // a double in the reference too
  bool throw_errors = true;
};

class ALTROSolver;

// typedefs.hpp:29-52
using CallbackFunction = std::function<void(const ALTROSolver*)>;
using ExplicitDynamicsFunction =
    std::function<void(double* xnext, const double* x, const double* u, float h)>;
using ExplicitDynamicsJacobian =
    std::function<void(double* jac, const double* x, const double* u, float h)>;
using CostFunction = std::function<a_float(const a_float* x, const a_float* u)>;
using CostGradient =
    std::function<void(a_float* dx, a_float* du, const a_float* x, const a_float* u)>;
using CostHessian = std::function<void(a_float* ddx, a_float* ddu, a_float* dxdu, const a_float* x,
                                       const a_float* u)>;
using ConstraintFunction = std::function<void(a_float* val, const a_float* x, const a_float* u)>;
using ConstraintJacobian = std::function<void(a_float* jac, const a_float* x, const a_float* u)>;

class ConstraintIndex {  // typedefs.hpp:55-66
 public:
  int KnotPointIndex() const { return k; }
  friend ALTROSolver;

 private:
  ConstraintIndex(int k, int i) : k(k), i(i) {}
  int k;
  int i;
};

namespace b200 {

// Device dynamics models (ids of include/altro_b200.h).  Use as
//   DeviceDynamics model(DeviceDynamics::Pendulum);
//   solver.SetExplicitDynamics(model.Function(), model.Jacobian());
struct DeviceDynamics {
  enum Model { Linear = 0, DoubleIntegrator = 1, Pendulum = 2, Bicycle4 = 3, Bicycle5 = 4, Chain = 5 };
  int model;
  double params[8];
  bool is_jacobian = false;
  explicit DeviceDynamics(Model m, std::vector<double> p = {});
  // The functors exist so the object can travel inside the reference's std::function types; they
  // cannot be evaluated on the host (the product has no CPU path) and say so when called.
  void operator()(double* out, const double* x, const double* u, float h) const;
  ExplicitDynamicsFunction Function() const;
  ExplicitDynamicsJacobian Jacobian() const;
};

// Device constraint families replacing the c(x,u) / Jacobian callbacks (typedefs.hpp:48-52):
//   Selector  rows c_i = scale_i * [x;u][idx_i] + off_i (idx_i = -1: c_i = off_i)
//   Affine    c = J [x;u] + e with a dense dim x (n+m) column-major J
//   Disc      c = r^2 - ([x;u][idx_0] - cx)^2 - ([x;u][idx_1] - cy)^2 (nonlinear, dim 1, INEQUALITY)
struct DeviceConstraint {
  enum Kind { Selector = 0, AffineRows = 1, Disc = 2 };
  int kind = Selector;
  std::vector<int> idx;
  std::vector<double> scale, off;  // Affine: off = e;  Disc: off = {cx, cy, r}
  std::vector<double> jac;         // Affine: J, column-major dim x (n+m)
  std::vector<double> off_batch;   // optional per-problem offsets / e / discs [B][dim or 3]
  bool is_jacobian = false;
  DeviceConstraint(std::vector<int> idx, std::vector<double> scale, std::vector<double> off);
  void operator()(a_float* out, const a_float* x, const a_float* u) const;
  ConstraintFunction Function() const;
  ConstraintJacobian Jacobian() const;
  // common families
  static DeviceConstraint Goal(const std::vector<double>& xf, bool x_minus_xf = true);
  static DeviceConstraint InputBox(int n, const std::vector<double>& u_max);
  static DeviceConstraint InputNormBound(int n, int m, double u_max);  // SOC rows [u; u_max]
  static DeviceConstraint StateBound(int index, double lo, double hi);
  static DeviceConstraint Affine(int dim, std::vector<double> J_colmajor, std::vector<double> e);
  static DeviceConstraint KeepOutDisc(int idx_a, int idx_b, double cx, double cy, double r);
};

}  // namespace b200

struct AltroStats {  // solver_stats.hpp:14-25
  SolveStatus status = SolveStatus::Unsolved;
  double solve_time_ms = 0.0;
  int iterations = 0;
  int outer_iterations = 0;
  double objective_value = 0.0;
  double stationarity = 0.0;
  double primal_feasibility = 0.0;
  double complimentarity = 0.0;
};

class SolverImpl;  // owns the altro_b200_solver handle and the host-side mirrors

class ALTROSolver {
 public:
  explicit ALTROSolver(int horizon_length, int batch = 1, int device = 0);
  ALTROSolver(const ALTROSolver& other) = delete;
  ALTROSolver(ALTROSolver&& other);
  ALTROSolver& operator=(const ALTROSolver& other) = delete;
  ALTROSolver& operator=(ALTROSolver&& other);
  ~ALTROSolver();

  // ---- problem definition (altro_solver.hpp:37-330)
  ErrorCodes SetDimension(int num_states, int num_inputs, int k_start = AllIndices, int k_stop = 0);
  ErrorCodes SetTimeStep(float h, int k_start = AllIndices, int k_stop = 0);
  ErrorCodes SetExplicitDynamics(ExplicitDynamicsFunction dynamics_function,
                                 ExplicitDynamicsJacobian dynamics_jacobian,
                                 int k_start = AllIndices, int k_stop = 0);
  ErrorCodes SetCostFunction(CostFunction cost_function, CostGradient cost_gradient,
                             CostHessian cost_hessian, int k_start = AllIndices, int k_stop = 0);
  ErrorCodes SetDiagonalCost(int num_states, int num_inputs, const a_float* Q_diag,
                             const a_float* R_diag, const a_float* q, const a_float* r, a_float c,
                             int k_start = AllIndices, int k_stop = 0);
  ErrorCodes SetQuadraticCost(int num_states, int num_inputs, const a_float* Q, const a_float* R,
                              const a_float* H, const a_float* q, const a_float* r, a_float c,
                              int k_start = AllIndices, int k_stop = 0);
  ErrorCodes SetLQRCost(int num_states, int num_inputs, const a_float* Q_diag,
                        const a_float* R_diag, const a_float* x_ref, const a_float* u_ref,
                        int k_start, int k_stop = 0);
  ErrorCodes SetConstraint(ConstraintFunction constraint_function,
                           ConstraintJacobian constraint_jacobian, int dim,
                           ConstraintType constraint_type, std::string label, int k_start,
                           int k_stop = 0, std::vector<ConstraintIndex>* con_inds = nullptr);
  // declared-but-undefined in the reference (altro_solver.hpp:257-290); here: INEQUALITY rows
  ErrorCodes SetStateUpperBound(a_float* x_max, int k_start, int k_stop = 0);
  ErrorCodes SetStateLowerBound(a_float* x_min, int k_start, int k_stop = 0);
  ErrorCodes SetInputUpperBound(a_float* u_max, int k_start, int k_stop = 0);
  ErrorCodes SetInputLowerBound(a_float* u_min, int k_start, int k_stop = 0);

  bool IsInitialized() const;
  ErrorCodes Initialize();
  ErrorCodes SetInitialState(const double* x0, int n);
  ErrorCodes SetState(const a_float* x, int n, int k_start = AllIndices, int k_stop = 0);
  ErrorCodes SetInput(const a_float* u, int m, int k_start = AllIndices, int k_stop = 0);

  // ---- batch extensions: problem-major host arrays
  ErrorCodes SetInitialStateBatch(const double* x0 /* [B][n] */);
  ErrorCodes SetLQRCostBatch(const a_float* Q_diag, const a_float* R_diag,
                             const a_float* x_ref /* [B][n] */, const a_float* u_ref /* [B][m] */,
                             int k_start, int k_stop = 0);
  ErrorCodes SetInputBatch(const a_float* u /* [B][k_stop-k_start][m] */, int k_start = AllIndices,
                           int k_stop = 0);
  ErrorCodes GetStatesBatch(a_float* X /* [B][N+1][n] */) const;
  ErrorCodes GetInputsBatch(a_float* U /* [B][N][m] */) const;
  ErrorCodes GetStatusBatch(SolveStatus* status /* [B] */) const;
  ErrorCodes GetIterationsBatch(int* iters /* [B] */) const;
  int GetBatchSize() const;

  ErrorCodes OpenLoopRollout();
  ErrorCodes UpdateLinearCosts(const a_float* q, const a_float* r, a_float c,
                               int k_start = AllIndices, int k_stop = 0);
  ErrorCodes ShiftTrajectory();
  // One receding-horizon step on the device (plant = model): x0 <- x_[1], ShiftTrajectory and, for
  // a tracking-window cost, the window moves one row (test/bicycle_test.cpp:302-337).
  ErrorCodes MpcStep();

  void SetOptions(const AltroOptions& opts);
  AltroOptions& GetOptions();
  const AltroOptions& GetOptions() const;

  SolveStatus Solve();

  // ---- getters (problem `b` of the batch, default the first)
  SolveStatus GetStatus() const;
  int GetIterations() const;
  a_float GetSolveTimeMs() const;
  a_float GetPrimalFeasibility() const;
  a_float GetFinalObjective() const;
  a_float CalcCost();

  int GetHorizonLength() const;
  int GetStateDim(int k) const;
  int GetInputDim(int k) const;
  float GetFinalTime() const;
  float GetTimeStep(int k) const;
  ErrorCodes GetState(a_float* x, int k) const;
  ErrorCodes GetInput(a_float* u, int k) const;
  ErrorCodes GetDualDynamics(a_float* y, int k) const;
  ErrorCodes GetFeedbackGain(a_float* K, int k) const;
  ErrorCodes GetFeedforwardGain(a_float* d, int k) const;
  // duals of a general constraint (declared, never defined in the reference: altro_solver.hpp:359,
  // :416); constraint_index comes from SetConstraint's con_inds
  ErrorCodes SetDualGeneric(const a_float* z, const ConstraintIndex& constraint_index);
  ErrorCodes GetDualGeneral(a_float* z, const ConstraintIndex& constraint_index) const;
  // KnotPointData member `name` of knot k of the first problem (the views the reference's tests
  // reach through solver_->data_[k]): "x" "u" "y" (x_, u_, y_) "xbar" "ubar" (x, u) "A" "B" "lx" "lu"
  // "K" "d" "P" "p" "q" "r" "c" "z" "z_est" "z_proj" "constraint_val" "lxx" "luu" "lux" "rho" -- the
  // reference's spelling with a trailing underscore ("K_", "lxx_", "z_est_" ...) is accepted too;
  // `out` holds the column-major block.
  ErrorCodes GetKnotPointField(const char* name, a_float* out, int k) const;

  void PrintStateTrajectory() const;
  void PrintInputTrajectory() const;

  std::unique_ptr<SolverImpl> solver_;  // public in the reference too (altro_solver.hpp:430)

 private:
  enum class LastIndexMode { Inclusive, Exclusive };
  ErrorCodes CheckKnotPointIndices(int& k_start, int& k_stop, LastIndexMode last_index) const;
  ErrorCodes AssertInitialized() const;
};

}  // namespace altro
