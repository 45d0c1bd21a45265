// altro_solver.cpp -- host-side C++ facade `altro::ALTROSolver` over the C ABI
// (include/altro_b200.h).  Mirrors src/altro/altro_solver.cpp of the reference: same index-range
// semantics (CheckKnotPointIndices, altro_solver.cpp:385-433), same error codes and
// print-or-throw convention, raw pointers copied in and caller-owned buffers filled on the way
// out.  All numerical work happens in the CUDA library; nothing here computes on the CPU.
#include "../../include/altro/altro_solver.hpp"

#include <chrono>
#include <cmath>
#include <cstring>

#include "../../include/altro_b200.h"

namespace altro {

const char* ErrorCodeToString(ErrorCodes err) { return altro_b200_error_string(static_cast<int>(err)); }

void PrintErrorCode(ErrorCodes err) {
  std::printf("Got error code %d: %s\n", static_cast<int>(err), ErrorCodeToString(err));
}

// ------------------------------------------------------------------ device functors
namespace b200 {

DeviceDynamics::DeviceDynamics(Model m, std::vector<double> p) : model(static_cast<int>(m)) {
  for (int i = 0; i < 8; ++i) params[i] = i < static_cast<int>(p.size()) ? p[i] : 0.0;
}
void DeviceDynamics::operator()(double*, const double*, const double*, float) const {
  ALTRO_THROW("device dynamics models are evaluated on the GPU only; there is no host path",
              ErrorCodes::Unsupported);
}
ExplicitDynamicsFunction DeviceDynamics::Function() const { return ExplicitDynamicsFunction(*this); }
ExplicitDynamicsJacobian DeviceDynamics::Jacobian() const {
  DeviceDynamics j = *this;
  j.is_jacobian = true;
  return ExplicitDynamicsJacobian(j);
}

DeviceConstraint::DeviceConstraint(std::vector<int> idx_, std::vector<double> scale_,
                                   std::vector<double> off_)
    : idx(std::move(idx_)), scale(std::move(scale_)), off(std::move(off_)) {}
void DeviceConstraint::operator()(a_float*, const a_float*, const a_float*) const {
  ALTRO_THROW("device constraints are evaluated on the GPU only; there is no host path",
              ErrorCodes::Unsupported);
}
ConstraintFunction DeviceConstraint::Function() const { return ConstraintFunction(*this); }
ConstraintJacobian DeviceConstraint::Jacobian() const {
  DeviceConstraint j = *this;
  j.is_jacobian = true;
  return ConstraintJacobian(j);
}
DeviceConstraint DeviceConstraint::Goal(const std::vector<double>& xf, bool x_minus_xf) {
  const int n = static_cast<int>(xf.size());
  std::vector<int> idx(n);
  std::vector<double> scale(n), off(n);
  for (int i = 0; i < n; ++i) {
    idx[i] = i;
    scale[i] = x_minus_xf ? 1.0 : -1.0;  // c = x - xf (double_integrator_test.cpp:174) or xf - x
    off[i] = x_minus_xf ? -xf[i] : xf[i];
  }
  return DeviceConstraint(idx, scale, off);
}
DeviceConstraint DeviceConstraint::InputBox(int n, const std::vector<double>& u_max) {
  const int m = static_cast<int>(u_max.size());  // double_integrator_test.cpp:283-304
  std::vector<int> idx(2 * m);
  std::vector<double> scale(2 * m), off(2 * m);
  for (int i = 0; i < m; ++i) {
    idx[i] = n + i;
    scale[i] = 1.0;
    off[i] = -u_max[i];
    idx[i + m] = n + i;
    scale[i + m] = -1.0;
    off[i + m] = -u_max[i];
  }
  return DeviceConstraint(idx, scale, off);
}
DeviceConstraint DeviceConstraint::InputNormBound(int n, int m, double u_max) {
  std::vector<int> idx(m + 1);  // double_integrator_test.cpp:405-424: c = [u; u_max]
  std::vector<double> scale(m + 1), off(m + 1, 0.0);
  for (int i = 0; i < m; ++i) {
    idx[i] = n + i;
    scale[i] = 1.0;
  }
  idx[m] = -1;
  scale[m] = 0.0;
  off[m] = u_max;
  return DeviceConstraint(idx, scale, off);
}
DeviceConstraint DeviceConstraint::StateBound(int index, double lo, double hi) {
  return DeviceConstraint({index, index}, {1.0, -1.0}, {-hi, lo});  // bicycle_test.cpp:189-196
}
DeviceConstraint DeviceConstraint::Affine(int dim, std::vector<double> J_colmajor, std::vector<double> e) {
  DeviceConstraint c(std::vector<int>(dim, -1), std::vector<double>(dim, 0.0), std::move(e));
  c.kind = AffineRows;
  c.jac = std::move(J_colmajor);
  return c;
}
DeviceConstraint DeviceConstraint::KeepOutDisc(int idx_a, int idx_b, double cx, double cy, double r) {
  DeviceConstraint c({idx_a, idx_b}, {1.0, 1.0}, {cx, cy, r});
  c.kind = Disc;
  return c;
}

}  // namespace b200

// ------------------------------------------------------------------ SolverImpl
class SolverImpl {
 public:
  SolverImpl(int N, int B, int device) : horizon_length_(N), batch_(B) {
    handle = altro_b200_create(N, B, device);
    h_.assign(N, 0.0f);
  }
  ~SolverImpl() {
    if (handle) altro_b200_destroy(handle);
  }
  altro_b200_solver* handle = nullptr;
  int horizon_length_;
  int batch_;
  int n = 0, m = 0;
  std::vector<float> h_;
  bool is_initialized_ = false;
  AltroOptions opts;
  AltroStats stats;
  int num_constraints = 0;
  // lazily refreshed host copies for the per-knot getters
  mutable bool cache_valid = false;
  mutable std::vector<double> X, U, Y, K, d;
  void Invalidate() const { cache_valid = false; }
  ErrorCodes Refresh() const {
    if (cache_valid) return ErrorCodes::NoError;
    const size_t N = horizon_length_, B = batch_;
    X.resize(B * (N + 1) * n);
    U.resize(B * N * m);
    Y.resize(B * (N + 1) * n);
    int e = altro_b200_get_states(handle, X.data());
    if (!e) e = altro_b200_get_inputs(handle, U.data());
    if (!e) e = altro_b200_get_dual_dynamics(handle, Y.data());
    if (e) return static_cast<ErrorCodes>(e);
    K.clear();
    d.clear();
    cache_valid = true;
    return ErrorCodes::NoError;
  }
};

static ErrorCodes EC(int code) { return static_cast<ErrorCodes>(code); }

ALTROSolver::ALTROSolver(int horizon_length, int batch, int device)
    : solver_(std::make_unique<SolverImpl>(horizon_length, batch, device)) {}
ALTROSolver::ALTROSolver(ALTROSolver&& other) = default;
ALTROSolver& ALTROSolver::operator=(ALTROSolver&& other) = default;
ALTROSolver::~ALTROSolver() = default;

#define REQUIRE_HANDLE()                                                                   \
  if (!solver_->handle)                                                                    \
  return ALTRO_THROW("No usable CUDA device: the solve path has no CPU fallback.", ErrorCodes::NoDevice)

// altro_solver.cpp:385-433
ErrorCodes ALTROSolver::CheckKnotPointIndices(int& k_start, int& k_stop,
                                              LastIndexMode last_index) const {
  const int terminal_index =
      last_index == LastIndexMode::Inclusive ? GetHorizonLength() : GetHorizonLength() - 1;
  if (k_start == AllIndices && k_stop == 0) {
    k_start = 0;
    k_stop = LastIndex;
  }
  if (k_start == 0 && k_stop == LastIndex) {
    k_start = 0;
    k_stop = terminal_index + 1;
  }
  if (k_stop <= 0) k_stop = k_start + 1;
  if (k_start < 0 || k_start > terminal_index) {
    return ALTRO_THROW("Knot point index out of range. Should be in range [0 - " +
                           std::to_string(terminal_index) + "], got " + std::to_string(k_start) + ".",
                       ErrorCodes::BadIndex);
  }
  if (k_stop < 0 || k_start > terminal_index + 1) {
    return ALTRO_THROW("Terminal knot point index out of range.", ErrorCodes::BadIndex);
  }
  if (k_stop > 0 && k_stop <= k_start) {
    std::printf("WARNING [ALTRO]: Stopping index %d not greater than starting index %d. Index range is empty.\n",
                k_stop, k_start);
  }
  return ErrorCodes::NoError;
}

ErrorCodes ALTROSolver::AssertInitialized() const {
  if (!IsInitialized()) return ALTRO_THROW("Solver must be initialized.", ErrorCodes::SolverNotInitialized);
  return ErrorCodes::NoError;
}

// altro_solver.cpp:26-47.  Dimensions must be uniform over the horizon on the device.
ErrorCodes ALTROSolver::SetDimension(int num_states, int num_inputs, int k_start, int k_stop) {
  REQUIRE_HANDLE();
  if (IsInitialized()) {
    return ALTRO_THROW("Cannot change the dimension once the solver has been initialized.",
                       ErrorCodes::SolverAlreadyInitialized);
  }
  ErrorCodes err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return err;
  if (num_states <= 0) return ErrorCodes::StateDimUnknown;
  if (num_inputs <= 0) return ErrorCodes::InputDimUnknown;
  if (solver_->n > 0) {
    if (solver_->n != num_states || solver_->m != num_inputs) {
      return ALTRO_THROW("The device path needs the same state/input dimension at every knot point.",
                         ErrorCodes::DimensionMismatch);
    }
    return ErrorCodes::NoError;
  }
  int e = altro_b200_set_dimension(solver_->handle, num_states, num_inputs);
  if (e) return EC(e);
  solver_->n = num_states;
  solver_->m = num_inputs;
  return ErrorCodes::NoError;
}

// altro_solver.cpp:49-63
ErrorCodes ALTROSolver::SetTimeStep(float h, int k_start, int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Exclusive);
  if (err != ErrorCodes::NoError) return err;
  if (h <= 0.0f) return ErrorCodes::TimestepNotPositive;
  for (int k = k_start; k < k_stop; ++k) solver_->h_[k] = h;
  return EC(altro_b200_set_time_step_range(solver_->handle, h, k_start, k_stop));
}

// altro_solver.cpp:68-81
ErrorCodes ALTROSolver::SetExplicitDynamics(ExplicitDynamicsFunction dynamics_function,
                                            ExplicitDynamicsJacobian dynamics_jacobian, int k_start,
                                            int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Exclusive);
  if (err != ErrorCodes::NoError) return err;
  if (solver_->n <= 0) {
    return ALTRO_THROW("Cannot set the dynamics. Dimensions haven't been set.", ErrorCodes::DimensionUnknown);
  }
  const b200::DeviceDynamics* f = dynamics_function.target<b200::DeviceDynamics>();
  const b200::DeviceDynamics* j = dynamics_jacobian.target<b200::DeviceDynamics>();
  if (!f || !j || f->model != j->model) {
    return ALTRO_THROW(
        "Host callbacks cannot run on the device: pass altro::b200::DeviceDynamics::Function()/Jacobian().",
        ErrorCodes::DynamicsFunNotSet);
  }
  return EC(altro_b200_set_model(solver_->handle, f->model, f->params, 8));
}

ErrorCodes ALTROSolver::SetCostFunction(CostFunction, CostGradient, CostHessian, int k_start,
                                        int k_stop) {
  // a no-op in the reference as well (altro_solver.cpp:86-96, quirk Q9): generic callback costs
  // are not live there; only quadratic / diagonal costs reach the hot path
  ErrorCodes err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return err;
  return ErrorCodes::NoError;
}

// KnotPointData::SetDiagonalCost (knotpoint_data.cpp:87-110); the reference's public wrapper
// never runs its loop (altro_solver.cpp:105), this one does (superset).
ErrorCodes ALTROSolver::SetDiagonalCost(int num_states, int num_inputs, const a_float* Q_diag,
                                        const a_float* R_diag, const a_float* q, const a_float* r,
                                        a_float c, int k_start, int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return err;
  if (num_states != solver_->n || num_inputs != solver_->m) return ErrorCodes::DimensionMismatch;
  solver_->Invalidate();
  return EC(altro_b200_set_diagonal_cost(solver_->handle, Q_diag, R_diag, q, r, &c, 0, k_start, k_stop));
}

// altro_solver.cpp:118-136 -> KnotPointData::SetQuadraticCost (knotpoint_data.cpp:64-85): dense Q,
// R and the cross term H.  Diagonal Q, R with H = 0 keep the diagonal device path (fewer bytes per
// knot); anything else selects the general instantiation of the kernels.
ErrorCodes ALTROSolver::SetQuadraticCost(int num_states, int num_inputs, const a_float* Q,
                                         const a_float* R, const a_float* H, const a_float* q,
                                         const a_float* r, a_float c, int k_start, int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return err;
  const int n = solver_->n, m = solver_->m;
  if (num_states != n || num_inputs != m) return ErrorCodes::DimensionMismatch;
  bool diagonal = true;
  std::vector<double> Qd(n), Rd(m);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      if (i == j) Qd[i] = Q[i + n * j];
      else if (Q[i + n * j] != 0.0) diagonal = false;
    }
  for (int j = 0; j < m && R; ++j)
    for (int i = 0; i < m; ++i) {
      if (i == j) Rd[i] = R[i + m * j];
      else if (R[i + m * j] != 0.0) diagonal = false;
    }
  for (int i = 0; i < m * n && H; ++i)
    if (H[i] != 0.0) diagonal = false;
  solver_->Invalidate();
  if (diagonal)
    return EC(altro_b200_set_diagonal_cost(solver_->handle, Qd.data(), Rd.data(), q, r, &c, 0, k_start, k_stop));
  return EC(altro_b200_set_quadratic_cost(solver_->handle, Q, R, H, q, r, &c, 0, k_start, k_stop));
}

// altro_solver.cpp:138-172
ErrorCodes ALTROSolver::SetLQRCost(int num_states, int num_inputs, const a_float* Q_diag,
                                   const a_float* R_diag, const a_float* x_ref,
                                   const a_float* u_ref, int k_start, int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return err;
  if (num_states != solver_->n) {
    return ALTRO_THROW("State dimension mismatch. Expected " + std::to_string(solver_->n) + ", got " +
                           std::to_string(num_states),
                       ErrorCodes::DimensionMismatch);
  }
  if (num_inputs != solver_->m) {
    return ALTRO_THROW("Input dimension mismatch. Expected " + std::to_string(solver_->m) + ", got " +
                           std::to_string(num_inputs),
                       ErrorCodes::DimensionMismatch);
  }
  solver_->Invalidate();
  return EC(altro_b200_set_lqr_cost(solver_->handle, Q_diag, R_diag, x_ref, u_ref, 0, k_start, k_stop));
}

ErrorCodes ALTROSolver::SetLQRCostBatch(const a_float* Q_diag, const a_float* R_diag,
                                        const a_float* x_ref, const a_float* u_ref, int k_start,
                                        int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return err;
  solver_->Invalidate();
  return EC(altro_b200_set_lqr_cost(solver_->handle, Q_diag, R_diag, x_ref, u_ref, 1, k_start, k_stop));
}

// altro_solver.cpp:177-190
ErrorCodes ALTROSolver::SetInitialState(const double* x0, int n) {
  REQUIRE_HANDLE();
  if (solver_->n > 0 && n != solver_->n) {
    return ALTRO_THROW("Dimension mismatch: The provided state dimension was " + std::to_string(n) +
                           ", but was previously set to " + std::to_string(solver_->n) + ".",
                       ErrorCodes::DimensionMismatch);
  }
  solver_->Invalidate();
  return EC(altro_b200_set_initial_state(solver_->handle, x0, 0));
}

ErrorCodes ALTROSolver::SetInitialStateBatch(const double* x0) {
  REQUIRE_HANDLE();
  solver_->Invalidate();
  return EC(altro_b200_set_initial_state(solver_->handle, x0, 1));
}

// altro_solver.cpp:192-223
ErrorCodes ALTROSolver::SetConstraint(ConstraintFunction constraint_function,
                                      ConstraintJacobian constraint_jacobian, int dim,
                                      ConstraintType constraint_type, std::string label,
                                      int k_start, int k_stop, std::vector<ConstraintIndex>* con_inds) {
  REQUIRE_HANDLE();
  (void)label;
  ErrorCodes err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return err;
  if (IsInitialized()) {
    return ALTRO_THROW("Cannot Set Constraints: Solver Already Initialized.", ErrorCodes::SolverAlreadyInitialized);
  }
  if (dim <= 0) {
    return ALTRO_THROW("Got a non-positive constraint dimension of " + std::to_string(dim), ErrorCodes::InvalidConstraintDim);
  }
  const b200::DeviceConstraint* c = constraint_function.target<b200::DeviceConstraint>();
  const b200::DeviceConstraint* j = constraint_jacobian.target<b200::DeviceConstraint>();
  if (!c || !j) {
    return ALTRO_THROW(
        "Host callbacks cannot run on the device: pass altro::b200::DeviceConstraint::Function()/Jacobian().",
        ErrorCodes::Unsupported);
  }
  const double* off_b = c->off_batch.empty() ? nullptr : c->off_batch.data();
  int e = 0;
  if (c->kind == b200::DeviceConstraint::AffineRows) {
    if (static_cast<int>(c->off.size()) != dim ||
        static_cast<int>(c->jac.size()) != dim * (solver_->n + solver_->m))
      return ErrorCodes::InvalidConstraintDim;
    e = altro_b200_set_constraint_affine(solver_->handle, static_cast<int>(constraint_type), dim,
                                         c->jac.data(), c->off.data(), off_b, k_start, k_stop);
  } else if (c->kind == b200::DeviceConstraint::Disc) {
    if (dim != 1 || constraint_type != ConstraintType::INEQUALITY) return ErrorCodes::InvalidConstraintDim;
    e = altro_b200_set_constraint_disc(solver_->handle, c->idx[0], c->idx[1], c->off.data(), off_b,
                                       k_start, k_stop);
  } else {
    if (static_cast<int>(c->idx.size()) != dim) return ErrorCodes::InvalidConstraintDim;
    e = altro_b200_set_constraint(solver_->handle, static_cast<int>(constraint_type), dim,
                                  c->idx.data(), c->scale.data(), c->off.data(), off_b, k_start, k_stop);
  }
  if (e) return EC(e);
  if (con_inds) {
    con_inds->reserve(k_stop - k_start);
    for (int k = k_start; k < k_stop; ++k) con_inds->emplace_back(ConstraintIndex(k, solver_->num_constraints));
  }
  solver_->num_constraints += 1;
  return ErrorCodes::NoError;
}

static ErrorCodes BoundRows(ALTROSolver* s, const a_float* v, int count, int base, bool upper,
                            int k_start, int k_stop) {
  std::vector<int> idx;
  std::vector<double> scale, off;
  for (int i = 0; i < count; ++i) {
    if (!std::isfinite(v[i])) continue;
    idx.push_back(base + i);
    scale.push_back(upper ? 1.0 : -1.0);
    off.push_back(upper ? -v[i] : v[i]);
  }
  if (idx.empty()) return ErrorCodes::NoError;
  b200::DeviceConstraint c(idx, scale, off);
  return s->SetConstraint(c.Function(), c.Jacobian(), static_cast<int>(idx.size()),
                          ConstraintType::INEQUALITY, "bound", k_start, k_stop);
}
ErrorCodes ALTROSolver::SetStateUpperBound(a_float* x_max, int k_start, int k_stop) {
  return BoundRows(this, x_max, solver_->n, 0, true, k_start, k_stop);
}
ErrorCodes ALTROSolver::SetStateLowerBound(a_float* x_min, int k_start, int k_stop) {
  return BoundRows(this, x_min, solver_->n, 0, false, k_start, k_stop);
}
ErrorCodes ALTROSolver::SetInputUpperBound(a_float* u_max, int k_start, int k_stop) {
  if (k_stop <= 0 && k_start == GetHorizonLength()) return ErrorCodes::InvalidOptAtTerminalKnotPoint;
  return BoundRows(this, u_max, solver_->m, solver_->n, true, k_start, k_stop);
}
ErrorCodes ALTROSolver::SetInputLowerBound(a_float* u_min, int k_start, int k_stop) {
  if (k_stop <= 0 && k_start == GetHorizonLength()) return ErrorCodes::InvalidOptAtTerminalKnotPoint;
  return BoundRows(this, u_min, solver_->m, solver_->n, false, k_start, k_stop);
}

bool ALTROSolver::IsInitialized() const { return solver_->is_initialized_; }

// altro_solver.cpp:225-229 -> SolverImpl::Initialize -> KnotPointData::Initialize checks
ErrorCodes ALTROSolver::Initialize() {
  REQUIRE_HANDLE();
  for (int k = 0; k < GetHorizonLength(); ++k) {
    if (!(solver_->h_[k] > 0.0f)) {
      return ALTRO_THROW("Cannot initialize solver. Timestep is nonpositive at timestep " + std::to_string(k),
                         ErrorCodes::TimestepNotPositive);
    }
  }
  int e = altro_b200_initialize(solver_->handle);
  if (e) return ALTRO_THROW("Failed to initialize the solver", EC(e));
  solver_->is_initialized_ = true;
  return ErrorCodes::NoError;
}

// altro_solver.cpp:231-251
ErrorCodes ALTROSolver::SetState(const a_float* x, int n, int k_start, int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = AssertInitialized();
  if (err != ErrorCodes::NoError) return err;
  err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return err;
  if (n != solver_->n) return ALTRO_THROW("State dimension mismatch.", ErrorCodes::DimensionMismatch);
  solver_->Invalidate();
  return EC(altro_b200_set_state(solver_->handle, x, 0, k_start, k_stop));
}
ErrorCodes ALTROSolver::SetInput(const a_float* u, int m, int k_start, int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = AssertInitialized();
  if (err != ErrorCodes::NoError) return err;
  err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Exclusive);
  if (err != ErrorCodes::NoError) return err;
  if (m != solver_->m) return ALTRO_THROW("Input dimension mismatch.", ErrorCodes::DimensionMismatch);
  solver_->Invalidate();
  return EC(altro_b200_set_input(solver_->handle, u, 0, k_start, k_stop));
}
ErrorCodes ALTROSolver::SetInputBatch(const a_float* u, int k_start, int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = AssertInitialized();
  if (err != ErrorCodes::NoError) return err;
  err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Exclusive);
  if (err != ErrorCodes::NoError) return err;
  solver_->Invalidate();
  return EC(altro_b200_set_input(solver_->handle, u, 2, k_start, k_stop));
}

ErrorCodes ALTROSolver::OpenLoopRollout() {
  REQUIRE_HANDLE();
  ErrorCodes err = AssertInitialized();
  if (err != ErrorCodes::NoError) return err;
  solver_->Invalidate();
  return EC(altro_b200_open_loop_rollout(solver_->handle));
}

void ALTROSolver::SetOptions(const AltroOptions& opts) { solver_->opts = opts; }
AltroOptions& ALTROSolver::GetOptions() { return solver_->opts; }
const AltroOptions& ALTROSolver::GetOptions() const { return solver_->opts; }

// altro_solver.cpp:257-260
SolveStatus ALTROSolver::Solve() {
  if (!solver_->handle || !IsInitialized()) {
    ALTRO_THROW("Solver must be initialized.", ErrorCodes::SolverNotInitialized);
    return SolveStatus::Unsolved;
  }
  const AltroOptions& o = solver_->opts;
  altro_b200_options co;
  altro_b200_default_options(&co);
  co.iterations_max = o.iterations_max;
  co.tol_primal_feasibility = o.tol_primal_feasibility;
  co.tol_stationarity = o.tol_stationarity;
  co.tol_meritfun_gradient = o.tol_meritfun_gradient;
  co.penalty_initial = o.penalty_initial;
  co.penalty_scaling = o.penalty_scaling;
  co.penalty_max = o.penalty_max;
  co.use_backtracking_linesearch = o.use_backtracking_linesearch != 0.0;
  altro_b200_set_options(solver_->handle, &co);
  const auto t0 = std::chrono::steady_clock::now();
  int e = altro_b200_solve(solver_->handle);
  const auto t1 = std::chrono::steady_clock::now();
  solver_->Invalidate();
  AltroStats& st = solver_->stats;
  st.solve_time_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  if (e) {
    ALTRO_THROW("Solve failed on the device", EC(e));
    st.status = SolveStatus::Unsolved;
    return st.status;
  }
  const int B = solver_->batch_;
  std::vector<int> status(B), iters(B);
  std::vector<double> phi(B), feas(B), stat(B);
  altro_b200_get_status(solver_->handle, status.data());
  altro_b200_get_iterations(solver_->handle, iters.data());
  altro_b200_get_final_objective(solver_->handle, phi.data());
  altro_b200_get_primal_feasibility(solver_->handle, feas.data());
  altro_b200_get_stationarity(solver_->handle, stat.data());
  st.status = static_cast<SolveStatus>(status[0]);
  st.iterations = iters[0];
  st.objective_value = phi[0];
  st.primal_feasibility = feas[0];
  st.stationarity = stat[0];
  return st.status;
}

// altro_solver.cpp:266-293
ErrorCodes ALTROSolver::UpdateLinearCosts(const a_float* q, const a_float* r, a_float c,
                                          int k_start, int k_stop) {
  REQUIRE_HANDLE();
  ErrorCodes err = AssertInitialized();
  if (err != ErrorCodes::NoError) return err;
  err = CheckKnotPointIndices(k_start, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return ALTRO_THROW("Error in UpdateLinearCosts", err);
  if (r != nullptr && k_stop - k_start == 1 && k_start == GetHorizonLength()) {
    return ALTRO_THROW("Cannot update linear input costs at terminal index", ErrorCodes::InvalidOptAtTerminalKnotPoint);
  }
  return EC(altro_b200_update_linear_costs(solver_->handle, q, r, &c, 0, k_start, k_stop));
}
ErrorCodes ALTROSolver::ShiftTrajectory() {
  REQUIRE_HANDLE();
  solver_->Invalidate();
  return EC(altro_b200_shift_trajectory(solver_->handle));
}

// ---- getters (altro_solver.cpp:299-360)
int ALTROSolver::GetHorizonLength() const { return solver_->horizon_length_; }
int ALTROSolver::GetBatchSize() const { return solver_->batch_; }
int ALTROSolver::GetStateDim(int) const { return solver_->n; }
int ALTROSolver::GetInputDim(int) const { return solver_->m; }
float ALTROSolver::GetTimeStep(int k) const { return solver_->h_[k]; }
float ALTROSolver::GetFinalTime() const {
  float t = 0.0f;
  for (float h : solver_->h_) t += h;
  return t;
}
SolveStatus ALTROSolver::GetStatus() const { return solver_->stats.status; }
int ALTROSolver::GetIterations() const { return solver_->stats.iterations; }
a_float ALTROSolver::GetSolveTimeMs() const { return solver_->stats.solve_time_ms; }
a_float ALTROSolver::GetPrimalFeasibility() const { return solver_->stats.primal_feasibility; }
a_float ALTROSolver::GetFinalObjective() const { return solver_->stats.objective_value; }
a_float ALTROSolver::CalcCost() {
  if (!solver_->handle || !IsInitialized()) return std::numeric_limits<double>::quiet_NaN();
  std::vector<double> cost(solver_->batch_);
  if (altro_b200_calc_cost(solver_->handle, cost.data())) return std::numeric_limits<double>::quiet_NaN();
  return cost[0];
}

ErrorCodes ALTROSolver::GetState(a_float* x, int k) const {
  REQUIRE_HANDLE();
  int k_stop = k + 1;
  ErrorCodes err = CheckKnotPointIndices(k, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return ALTRO_THROW("Error in GetState with k = " + std::to_string(k), err);
  err = solver_->Refresh();
  if (err != ErrorCodes::NoError) return err;
  std::memcpy(x, &solver_->X[static_cast<size_t>(k) * solver_->n], sizeof(double) * solver_->n);
  return ErrorCodes::NoError;
}
ErrorCodes ALTROSolver::GetInput(a_float* u, int k) const {
  REQUIRE_HANDLE();
  int k_stop = k + 1;
  ErrorCodes err = CheckKnotPointIndices(k, k_stop, LastIndexMode::Exclusive);
  if (err != ErrorCodes::NoError) return ALTRO_THROW("Error in GetInput with k = " + std::to_string(k), err);
  err = solver_->Refresh();
  if (err != ErrorCodes::NoError) return err;
  std::memcpy(u, &solver_->U[static_cast<size_t>(k) * solver_->m], sizeof(double) * solver_->m);
  return ErrorCodes::NoError;
}
ErrorCodes ALTROSolver::GetDualDynamics(a_float* y, int k) const {
  REQUIRE_HANDLE();
  int k_stop = k + 1;
  ErrorCodes err = CheckKnotPointIndices(k, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return ALTRO_THROW("Error in GetDualDynamics with k = " + std::to_string(k), err);
  err = solver_->Refresh();
  if (err != ErrorCodes::NoError) return err;
  std::memcpy(y, &solver_->Y[static_cast<size_t>(k) * solver_->n], sizeof(double) * solver_->n);
  return ErrorCodes::NoError;
}
ErrorCodes ALTROSolver::GetFeedbackGain(a_float* K, int k) const {
  REQUIRE_HANDLE();
  int k_stop = k + 1;
  ErrorCodes err = CheckKnotPointIndices(k, k_stop, LastIndexMode::Exclusive);
  if (err != ErrorCodes::NoError) return err;
  const size_t N = solver_->horizon_length_, B = solver_->batch_, mn = solver_->m * solver_->n;
  if (solver_->K.empty()) {
    solver_->K.resize(B * N * mn);
    int e = altro_b200_get_feedback_gains(solver_->handle, solver_->K.data());
    if (e) return EC(e);
  }
  std::memcpy(K, &solver_->K[static_cast<size_t>(k) * mn], sizeof(double) * mn);
  return ErrorCodes::NoError;
}
ErrorCodes ALTROSolver::MpcStep() {
  REQUIRE_HANDLE();
  solver_->Invalidate();
  return EC(altro_b200_mpc_step(solver_->handle));
}
ErrorCodes ALTROSolver::SetDualGeneric(const a_float* z, const ConstraintIndex& ci) {
  REQUIRE_HANDLE();
  if (!z) return ErrorCodes::InvalidPointer;
  if (!IsInitialized()) return ErrorCodes::SolverNotInitialized;
  return EC(altro_b200_set_dual_general(solver_->handle, ci.i, ci.k, z, 0));
}
ErrorCodes ALTROSolver::GetDualGeneral(a_float* z, const ConstraintIndex& ci) const {
  REQUIRE_HANDLE();
  if (!z) return ErrorCodes::InvalidPointer;
  if (!solver_->is_initialized_) return ErrorCodes::SolverNotInitialized;
  const int dim = altro_b200_get_constraint_dim(solver_->handle, ci.i);
  if (dim <= 0) return ErrorCodes::BadIndex;
  std::vector<double> all(static_cast<size_t>(solver_->batch_) * dim);
  int e = altro_b200_get_dual_general(solver_->handle, ci.i, ci.k, all.data());
  if (e) return EC(e);
  std::memcpy(z, all.data(), sizeof(double) * static_cast<size_t>(dim));  // the first problem
  return ErrorCodes::NoError;
}

ErrorCodes ALTROSolver::GetKnotPointField(const char* name, a_float* out, int k) const {
  REQUIRE_HANDLE();
  if (!name || !out) return ErrorCodes::InvalidPointer;
  int k_stop = k + 1;
  ErrorCodes err = CheckKnotPointIndices(k, k_stop, LastIndexMode::Inclusive);
  if (err != ErrorCodes::NoError) return err;
  // the reference spells members with a trailing underscore (knotpoint_data.hpp:160-233)
  std::string nm(name);
  if (nm.size() > 1 && nm.back() == '_') nm.pop_back();
  name = nm.c_str();
  int rows = 0;
  int e = altro_b200_get_field(solver_->handle, name, nullptr, &rows);
  if (e) return EC(e);
  const size_t N = solver_->horizon_length_, B = solver_->batch_;
  std::vector<double> all(B * (N + 1) * static_cast<size_t>(rows));
  e = altro_b200_get_field(solver_->handle, name, all.data(), nullptr);
  if (e) return EC(e);
  std::memcpy(out, &all[static_cast<size_t>(k) * rows], sizeof(double) * rows);
  return ErrorCodes::NoError;
}
ErrorCodes ALTROSolver::GetFeedforwardGain(a_float* d, int k) const {
  REQUIRE_HANDLE();
  int k_stop = k + 1;
  ErrorCodes err = CheckKnotPointIndices(k, k_stop, LastIndexMode::Exclusive);
  if (err != ErrorCodes::NoError) return err;
  const size_t N = solver_->horizon_length_, B = solver_->batch_, m = solver_->m;
  if (solver_->d.empty()) {
    solver_->d.resize(B * N * m);
    int e = altro_b200_get_feedforward_gains(solver_->handle, solver_->d.data());
    if (e) return EC(e);
  }
  std::memcpy(d, &solver_->d[static_cast<size_t>(k) * m], sizeof(double) * m);
  return ErrorCodes::NoError;
}

ErrorCodes ALTROSolver::GetStatesBatch(a_float* X) const {
  REQUIRE_HANDLE();
  return EC(altro_b200_get_states(solver_->handle, X));
}
ErrorCodes ALTROSolver::GetInputsBatch(a_float* U) const {
  REQUIRE_HANDLE();
  return EC(altro_b200_get_inputs(solver_->handle, U));
}
ErrorCodes ALTROSolver::GetStatusBatch(SolveStatus* status) const {
  REQUIRE_HANDLE();
  std::vector<int> st(solver_->batch_);
  int e = altro_b200_get_status(solver_->handle, st.data());
  for (int b = 0; b < solver_->batch_; ++b) status[b] = static_cast<SolveStatus>(st[b]);
  return EC(e);
}
ErrorCodes ALTROSolver::GetIterationsBatch(int* iters) const {
  REQUIRE_HANDLE();
  return EC(altro_b200_get_iterations(solver_->handle, iters));
}

// altro_solver.cpp:464-476
void ALTROSolver::PrintStateTrajectory() const {
  std::printf("STATE TRAJECTORY:\n");
  if (solver_->Refresh() != ErrorCodes::NoError) return;
  for (int k = 0; k <= GetHorizonLength(); ++k) {
    std::printf(" x[%03d]: [", k);
    for (int i = 0; i < solver_->n; ++i) std::printf("%s%g", i ? " " : "", solver_->X[static_cast<size_t>(k) * solver_->n + i]);
    std::printf("]\n");
  }
}
void ALTROSolver::PrintInputTrajectory() const {
  std::printf("INPUT TRAJECTORY:\n");
  if (solver_->Refresh() != ErrorCodes::NoError) return;
  for (int k = 0; k < GetHorizonLength(); ++k) {
    std::printf(" u[%03d]: [", k);
    for (int i = 0; i < solver_->m; ++i) std::printf("%s%g", i ? " " : "", solver_->U[static_cast<size_t>(k) * solver_->m + i]);
    std::printf("]\n");
  }
}

}  // namespace altro
