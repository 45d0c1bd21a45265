// capi.cu -- the extern "C" boundary (include/altro_b200.h, section C) and the host side of the
// batched solver handle: HBM allocation in the problem-fastest layout, host<->device staging
// with on-device transposition, model dispatch and kernel launches.
//
// Mirrors altro::ALTROSolver (src/altro/altro_solver.cpp) call for call; error codes are the
// reference's ErrorCodes integers (exceptions.hpp:24-51).  No CPU fallback: without a CUDA device
// altro_b200_create returns NULL and compute calls return ALTRO_B200_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/altro_b200.h"
#include "device_problem.h"
#include "launchers.h"
#include "models.cuh"

using namespace altro_b200;

#define CUDA_OK(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      fprintf(stderr, "altro_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e__), \
              __FILE__, __LINE__);                                                      \
      return ALTRO_B200_ERR_NO_DEVICE;                                                  \
    }                                                                                   \
  } while (0)

// ------------------------------------------------------------------ layout kernels
// One field of a record stream (device_problem.h) as the layout kernels see it: element (k, e) of
// problem b lives at p[(b / 32) * GS + k * R + e * 32 + b % 32]; a flat per-problem index
// j = k * E + e enumerates the elements the way the host's problem-major arrays do.
struct FieldView {
  double* p;
  int E;
  long R, GS;
};
__device__ __forceinline__ long fv_index(const FieldView& f, long j, int b) {
  const long k = j / f.E;
  const int e = (int)(j - k * f.E);
  return (long)(b >> 5) * f.GS + k * f.R + (long)e * 32 + (b & 31);
}

// dst(j, b) = src[b * W + j]   (host problem-major -> knot records), 32x32 smem tiles
__global__ void k_pm_to_pf(const double* __restrict__ src, int B, long W, FieldView dst) {
  __shared__ double tile[32][33];
  const long j0 = (long)blockIdx.x * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int bb = b0 + r;
    const long j = j0 + threadIdx.x;
    if (bb < B && j < W) tile[r][threadIdx.x] = src[(long)bb * W + j];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long j = j0 + r;
    const int bb = b0 + threadIdx.x;
    if (bb < B && j < W) dst.p[fv_index(dst, j, bb)] = tile[threadIdx.x][r];
  }
}

// dst[b * W + j] = src(j, b)   (knot records -> host problem-major)
__global__ void k_pf_to_pm(FieldView src, int B, long W, double* __restrict__ dst) {
  __shared__ double tile[32][33];
  const long j0 = (long)blockIdx.x * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long j = j0 + r;
    const int bb = b0 + threadIdx.x;
    if (bb < B && j < W) tile[r][threadIdx.x] = src.p[fv_index(src, j, bb)];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int bb = b0 + r;
    const long j = j0 + threadIdx.x;
    if (bb < B && j < W) dst[(long)bb * W + j] = tile[threadIdx.x][r];
  }
}

// dst(j, b) = src[j]  for all b   (broadcast a shared row set to every problem)
__global__ void k_broadcast(const double* __restrict__ src, int B, long W, FieldView dst) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (long j = blockIdx.y; j < W; j += gridDim.y) dst.p[fv_index(dst, j, b)] = src[j];
}

// dst(j, b) = src(j, b)  (device-to-device copy between two fields of the same width)
__global__ void k_copy_field(FieldView src, FieldView dst, int B, long W) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (long j = blockIdx.y; j < W; j += gridDim.y) dst.p[fv_index(dst, j, b)] = src.p[fv_index(src, j, b)];
}

__global__ void k_fill_field(FieldView dst, int B, long W, double v) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (long j = blockIdx.y; j < W; j += gridDim.y) dst.p[fv_index(dst, j, b)] = v;
}

__global__ void k_fill(double* dst, long count, double v) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long)gridDim.x * blockDim.x)
    dst[i] = v;
}

// SetLQRCost (altro_solver.cpp:159-169) for knots [k0,k1): q = -(Qd .* xref), r = -(Rd .* uref),
// c = 1/2 xref' Qd xref (+ 1/2 uref' Rd uref for k < N).
// ref_mode 0: xref [n], uref [m] shared; 1: per problem, problem-major [B][n]/[B][m];
// 2: window tables xtab [T][n], utab [T][m] with row = offsets[b] + k.
__global__ void k_lqr_cost(int n, int m, int N, int B, int k0, int k1,
                           const double* __restrict__ Qd, const double* __restrict__ Rd,
                           const double* __restrict__ xref, const double* __restrict__ uref,
                           int ref_mode, const int* __restrict__ offsets, FieldView q, FieldView r,
                           FieldView c) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int k = k0 + blockIdx.y; k < k1; k += gridDim.y) {
    const double* xr = xref;
    const double* ur = uref;
    if (ref_mode == 1) {
      xr = xref + (long)b * n;
      ur = uref + (long)b * m;
    } else if (ref_mode == 2) {
      const long row = offsets[b] + k;
      xr = xref + row * n;
      ur = uref + row * m;
    }
    double cc = 0.0;
    for (int i = 0; i < n; ++i) {
      const double w = Qd[k * n + i];
      q.p[fv_index(q, (long)k * n + i, b)] = -(w * xr[i]);
      cc += (0.5 * xr[i]) * w * xr[i];
    }
    if (k < N) {
      double cu = 0.0;
      for (int i = 0; i < m; ++i) {
        const double w = Rd[k * m + i];
        r.p[fv_index(r, (long)k * m + i, b)] = -(w * ur[i]);
        cu += (0.5 * ur[i]) * w * ur[i];
      }
      cc += cu;
    }
    c.p[fv_index(c, k, b)] = cc;
  }
}

// The reference's receding-horizon cost update (test/bicycle_test.cpp:317-328): for every knot k
// UpdateLinearCosts(q, nullptr, c, k) with q = -(Qd .* xref[row]), c = -(1/2 q . xref[row]) and
// the FROZEN input part c_u added for k < N; r_k keeps the value of the original SetLQRCost
// (altro_solver.cpp:266-281 leaves r alone when the pointer is null).  row = offsets[b] + k.
__global__ void k_window_linear_update(int n, int N, int B, const double* __restrict__ Qd,
                                       const double* __restrict__ xtab,
                                       const int* __restrict__ offsets, FieldView q, FieldView c,
                                       double c_u) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int k = blockIdx.y; k <= N; k += gridDim.y) {
    const double* xr = xtab + ((long)offsets[b] + k) * n;
    double dot = 0.0;
    for (int i = 0; i < n; ++i) {
      const double qi = -(Qd[k * n + i] * xr[i]);
      q.p[fv_index(q, (long)k * n + i, b)] = qi;
      dot += qi * xr[i];
    }
    double cc = -(0.5 * dot);
    if (k < N) cc += c_u;
    c.p[fv_index(c, k, b)] = cc;
  }
}

__global__ void k_add_int(int* v, int B, int add) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) v[b] += add;
}

// ShiftTrajectory (altro_solver.cpp:283-293): x_[k] = x_[k+1] for k < N, u_[k] = u_[k+1] for k < N-1
__global__ void k_shift(int n, int m, int N, int B, FieldView x, FieldView u) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int k = 0; k < N; ++k) {
    for (int i = 0; i < n; ++i)
      x.p[fv_index(x, (long)k * n + i, b)] = x.p[fv_index(x, (long)(k + 1) * n + i, b)];
    if (k < N - 1)
      for (int i = 0; i < m; ++i)
        u.p[fv_index(u, (long)k * m + i, b)] = u.p[fv_index(u, (long)(k + 1) * m + i, b)];
  }
}

// ------------------------------------------------------------------ handle
struct altro_b200_solver {
  int N = 0, B = 0, device = 0;
  long Bp = 0;
  int n = 0, m = 0;
  float h = 0.0f;
  int model = -1;
  double params[8] = {0};
  bool dims_set = false, cost_set = false, initialized = false;
  cudaStream_t stream = nullptr;
  long launches = 0;
  long bytes = 0;
  altro_b200_options opts;

  // device arrays: knot records (device_problem.h); every per-knot field below points into `rec`
  double* rec = nullptr;
  long R = 0, GS = 0;    // knot / group stride of the main record stream (doubles)
  double* zrec = nullptr;
  long Rz = 0, GSz = 0;  // dual record stream
  long Rs = 0;           // candidate-slot record stream
  int G = 0;             // groups of 32 problems
  double *Qd = nullptr, *Rd = nullptr, *lin = nullptr;
  double *q = nullptr, *r = nullptr, *c = nullptr, *x0 = nullptr;
  double *xbar = nullptr, *ubar = nullptr, *x = nullptr, *u = nullptr, *y = nullptr;
  double* u_init = nullptr;  // last SetInput guess, restored by altro_b200_reset_trajectory
  double *A = nullptr, *Bm = nullptr, *lx = nullptr, *lu = nullptr;
  double *K = nullptr, *d = nullptr, *P = nullptr, *p = nullptr;
  double *z = nullptr, *zest = nullptr, *rho = nullptr;
  int *status = nullptr, *iters = nullptr, *merit_evals = nullptr, *ls_fail = nullptr;
  double *phi = nullptr, *stat = nullptr, *feas = nullptr;
  // phase pipeline state
  LsMachine* ls = nullptr;
  double *alpha_eval = nullptr, *alpha_bt = nullptr, *phi_eval = nullptr, *phi0 = nullptr, *dphi0 = nullptr;
  double *xs = nullptr, *us = nullptr, *phi_s = nullptr;
  int* sel = nullptr;
  int *spec_base = nullptr, *spec_known = nullptr;
  unsigned long long *stat_acc = nullptr, *feas_acc = nullptr;
  int nslots = 6;   // candidate steps per speculative line-search round (slot 0 = requested step)
  int nstore = 15;  // speculative slots 1..nstore keep their trajectory (candidate slot buffers)
  int nstore_alloc = 0;  // slot buffers allocated by Initialize
  int *flags = nullptr, *iter_count = nullptr;
  int* d_done = nullptr;                 // [kMaxSplit] stopped problems per sub-batch
  unsigned long long* d_prof = nullptr;  // [kMaxSplit][16] sub-phase clocks of k_phase_forward
  unsigned long long* ls_hist = nullptr;
  PhaseHost ph;        // accumulated statistics + device limits
  int solve_mode = 0;  // 0: phase pipeline (default), 1: single persistent kernel
  // Riccati sweep: 0 one warp per group, 1 the warps of a CTA (solver_team.cuh), -1 by block size:
  // the team form where the blocks do not fit the registers of one thread (n > 6)
  int backward_team = -1;
  // forward kernel variant: 1 the derivative half of a merit evaluation is done by a follower warp
  // behind the rollout warp (default), 0 separate knot-parallel expansion + d(phi) scan
  int follow_deriv = 1;
  int qrc_uniform_enable = 1;  // ALTRO_B200_QRC_UNIFORM=0 streams [q r c] with every knot regardless
  int spec_round1 = 1;
  int fwd_depth = 8;  // staging depth cap of k_phase_forward (BulkPipe holds up to 8 stages)
  // pipelined sub-batches: the groups are cut into `nsplit` contiguous ranges, each on its own
  // stream, so that the sweeps of one range (one busy warp per group) overlap the rollouts and
  // knot-parallel phases of another.  One host thread enqueues all of them.
  static constexpr int kMaxSplit = 32;
  int nsplit = 0;  // 0: choose from the batch size
  cudaStream_t sub_stream[kMaxSplit] = {nullptr};
  PhaseHost sub_ph[kMaxSplit];
  bool sub_ready[kMaxSplit] = {false};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxSplit] = {nullptr};
  // tracking-window cost
  double *xtab = nullptr, *utab = nullptr;
  int* offsets = nullptr;
  int T = 0, T_alloc = 0;
  int off_min = 0, off_max = 0;  // host-side range of the window offsets (bounds checks)
  int mpc_mode = 1;     // cost update of altro_b200_mpc_step: 1 reference (q, c; r frozen), 0 re-window q, r, c
  double mpc_cu = 0.0;  // frozen input part of c_k, k < N (mode 1)
  // constraints
  ConTable con_h;
  ConTable* con_d = nullptr;
  std::vector<double*> off_b;
  // host mirrors of the shared weights
  std::vector<double> Qd_h, Rd_h, lin_h;
  std::vector<float> h_h;   // per-knot time steps (empty until SetTimeStep is used with a range)
  bool h_uniform = true;
  float* hk = nullptr;      // device copy of h_h while the steps differ
  // which setter call wrote the linear cost terms q_k, r_k, c_k of knot k: knots written by ONE call
  // with k-independent values share an id (> 0), anything else is -1 (DeviceProblem::qrc_uniform)
  std::vector<int> qrc_id;
  int qrc_next = 0;
  // dense quadratic cost (SetQuadraticCost): host mirrors of every knot's Q, R, H (diagonal-cost
  // knots hold diag(Qd), diag(Rd), 0) and their device copies, uploaded once a dense knot exists
  std::vector<double> Qf_h, Rf_h, Hf_h;
  double *Qf = nullptr, *Rf = nullptr, *Hf = nullptr;
  bool dense_cost = false;
  // staging
  double* stage = nullptr;
  long stage_count = 0;
  double* view_buf = nullptr;  // scratch stream of the derived KnotPointData views
  long view_count = 0;
  std::vector<void*> allocs;
};

static int dalloc(altro_b200_solver* s, void** p, size_t bytes, bool zero = true) {
  CUDA_OK(cudaMalloc(p, bytes > 0 ? bytes : 8));
  if (zero) CUDA_OK(cudaMemsetAsync(*p, 0, bytes > 0 ? bytes : 8, s->stream));
  s->allocs.push_back(*p);
  s->bytes += (long)bytes;
  return 0;
}
// frees one tracked allocation
static void release(altro_b200_solver* s, void* p) {
  for (size_t i = 0; i < s->allocs.size(); ++i)
    if (s->allocs[i] == p) {
      cudaFree(p);
      s->allocs.erase(s->allocs.begin() + (long)i);
      return;
    }
}
#define DALLOC(s, ptr, count)                                                      \
  do {                                                                             \
    int e__ = dalloc((s), (void**)&(ptr), sizeof(*(ptr)) * (size_t)(count));       \
    if (e__) return e__;                                                           \
  } while (0)

static int ensure_stage(altro_b200_solver* s, long count) {
  if (count <= s->stage_count) return 0;
  if (s->stage) {
    CUDA_OK(cudaStreamSynchronize(s->stream));
    CUDA_OK(cudaFree(s->stage));
    s->bytes -= s->stage_count * 8;
  }
  CUDA_OK(cudaMalloc((void**)&s->stage, (size_t)count * 8));
  s->stage_count = count;
  s->bytes += count * 8;
  return 0;
}

// view of knots [k0, ...) of a per-knot field with E rows
static FieldView fview(const altro_b200_solver* s, double* field, int E, int k0 = 0) {
  return FieldView{field + (long)k0 * s->R, E, s->R, s->GS};
}
// view of a per-group block that is not per knot: [G][E][32]
static FieldView gview(double* base, int E) { return FieldView{base, E, 0, (long)E * 32}; }

// host problem-major [B][W] -> device field
static int upload_pm(altro_b200_solver* s, const double* host, long W, FieldView dst) {
  int e = ensure_stage(s, (long)s->B * W);
  if (e) return e;
  CUDA_OK(cudaMemcpyAsync(s->stage, host, (size_t)s->B * W * 8, cudaMemcpyHostToDevice, s->stream));
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((s->B + 31) / 32));
  k_pm_to_pf<<<grid, dim3(32, 8), 0, s->stream>>>(s->stage, s->B, W, dst);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}
// host shared [W] -> broadcast to every problem
static int upload_shared(altro_b200_solver* s, const double* host, long W, FieldView dst) {
  int e = ensure_stage(s, W);
  if (e) return e;
  CUDA_OK(cudaMemcpyAsync(s->stage, host, (size_t)W * 8, cudaMemcpyHostToDevice, s->stream));
  dim3 grid((unsigned)((s->B + 127) / 128), (unsigned)(W < 1024 ? W : 1024));
  k_broadcast<<<grid, 128, 0, s->stream>>>(s->stage, s->B, W, dst);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}
// device field -> host problem-major [B][W]
static int download_pm(altro_b200_solver* s, FieldView src, long W, double* host) {
  int e = ensure_stage(s, (long)s->B * W);
  if (e) return e;
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((s->B + 31) / 32));
  k_pf_to_pm<<<grid, dim3(32, 8), 0, s->stream>>>(src, s->B, W, s->stage);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(host, s->stage, (size_t)s->B * W * 8, cudaMemcpyDeviceToHost, s->stream));
  CUDA_OK(cudaStreamSynchronize(s->stream));
  return 0;
}

// Index-range resolution of ALTROSolver::CheckKnotPointIndices (altro_solver.cpp:385-433)
static int resolve_range(const altro_b200_solver* s, int& k_start, int& k_stop, bool inclusive) {
  const int terminal_index = inclusive ? s->N : s->N - 1;
  if (k_start == ALTRO_B200_ALL_INDICES && k_stop == 0) {
    k_start = 0;
    k_stop = ALTRO_B200_LAST_INDEX;
  }
  if (k_start == 0 && k_stop == ALTRO_B200_LAST_INDEX) {
    k_start = 0;
    k_stop = terminal_index + 1;
  }
  if (k_stop <= 0) k_stop = k_start + 1;
  if (k_start < 0 || k_start > terminal_index) return ALTRO_B200_BAD_INDEX;
  if (k_stop > terminal_index + 1) return ALTRO_B200_BAD_INDEX;
  return ALTRO_B200_NO_ERROR;
}
// 0 < k_stop <= k_start: the reference only warns and loops over nothing (altro_solver.cpp:423-427);
// callers return NoError without touching anything
static bool empty_range(int k_start, int k_stop) { return k_stop <= k_start; }

// ------------------------------------------------------------------ model dispatch
// The per-model kernels are instantiated in solve_inst.cu (one translation unit per group so the
// build parallelises); launchers.h declares one launcher per compiled-in model.
static solve_launcher find_launcher(int model, int n, int m, const double* prm) {
  switch (model) {
    case MODEL_LINEAR:
      if (n == 4 && m == 2) return launch_solve_linear_4_2;
      if (n == 2 && m == 1) return launch_solve_linear_2_1;
      if (n == 6 && m == 3) return launch_solve_linear_6_3;
      return nullptr;
    case MODEL_DOUBLE_INTEGRATOR: {
      const int dim = (int)prm[0];
      if (dim == 1 && n == 2 && m == 1) return launch_solve_di_1;
      if (dim == 2 && n == 4 && m == 2) return launch_solve_di_2;
      if (dim == 3 && n == 6 && m == 3) return launch_solve_di_3;
      return nullptr;
    }
    case MODEL_PENDULUM:
      return (n == 2 && m == 1) ? launch_solve_pendulum : nullptr;
    case MODEL_BICYCLE4:
      return (n == 4 && m == 2) ? launch_solve_bicycle4 : nullptr;
    case MODEL_BICYCLE5:
      return (n == 5 && m == 2) ? launch_solve_bicycle5 : nullptr;
    case MODEL_CHAIN:
      if (n == 4 && m == 2) return launch_solve_chain_4_2;
      if (n == 4 && m == 4) return launch_solve_chain_4_4;
      if (n == 6 && m == 2) return launch_solve_chain_6_2;
      if (n == 6 && m == 4) return launch_solve_chain_6_4;
      if (n == 12 && m == 2) return launch_solve_chain_12_2;
      if (n == 12 && m == 4) return launch_solve_chain_12_4;
      return nullptr;
  }
  return nullptr;
}

// ------------------------------------------------------------------ C ABI
extern "C" {

int altro_b200_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) return 0;
  return cnt;
}

const char* altro_b200_error_string(int code) {
  switch (code) {  // exceptions.cpp:12-94
    case ALTRO_B200_NO_ERROR: return "no error";
    case ALTRO_B200_STATE_DIM_UNKNOWN: return "state dimension unknown";
    case ALTRO_B200_INPUT_DIM_UNKNOWN: return "input dimension unknown";
    case ALTRO_B200_NEXT_STATE_DIM_UNKNOWN: return "next state dimension unknown";
    case ALTRO_B200_DIMENSION_UNKNOWN: return "dimension unknown";
    case ALTRO_B200_BAD_INDEX: return "bad index";
    case ALTRO_B200_DIMENSION_MISMATCH: return "dimension mismatch";
    case ALTRO_B200_SOLVER_NOT_INITIALIZED: return "solver not initialized";
    case ALTRO_B200_SOLVER_ALREADY_INITIALIZED: return "solver already initialized";
    case ALTRO_B200_NON_POSITIVE: return "expected a positive value";
    case ALTRO_B200_TIMESTEP_NOT_POSITIVE: return "timestep not positive";
    case ALTRO_B200_COST_FUN_NOT_SET: return "cost function not set";
    case ALTRO_B200_DYNAMICS_FUN_NOT_SET: return "dynamics function not set";
    case ALTRO_B200_INVALID_OPT_AT_TERMINAL: return "invalid operation at terminal knot point";
    case ALTRO_B200_MAX_CONSTRAINTS_EXCEEDED: return "max number of constraints exceeded";
    case ALTRO_B200_INVALID_CONSTRAINT_DIM: return "invalid constraint dimension";
    case ALTRO_B200_CHOLESKY_FAILED: return "cholesky factorization failed";
    case ALTRO_B200_OP_ONLY_VALID_AT_TERMINAL: return "operation only valid at terminal knot point";
    case ALTRO_B200_INVALID_POINTER: return "invalid pointer";
    case ALTRO_B200_BACKWARD_PASS_FAILED: return "backward pass failed";
    case ALTRO_B200_LINESEARCH_FAILED: return "line search failed";
    case ALTRO_B200_MERIT_GRADIENT_TOO_SMALL: return "merit function gradient too small";
    case ALTRO_B200_INVALID_BOUND_CONSTRAINT: return "invalid bound constraint";
    case ALTRO_B200_NON_POSITIVE_PENALTY: return "penalty must be positive";
    case ALTRO_B200_COST_NOT_QUADRATIC: return "cost function not quadratic";
    case ALTRO_B200_FILE_ERROR: return "file error";
    case ALTRO_B200_ERR_NO_DEVICE: return "no usable CUDA device (no CPU fallback exists)";
    case ALTRO_B200_ERR_UNSUPPORTED: return "model/dimension combination not compiled in";
  }
  return "unknown error";
}

void altro_b200_default_options(altro_b200_options* o) {  // solver_options.hpp:18-37
  o->iterations_max = 200;
  o->tol_primal_feasibility = 1e-4;
  o->tol_stationarity = 1e-4;
  o->tol_meritfun_gradient = 1e-8;
  o->penalty_initial = 1.0;
  o->penalty_scaling = 10.0;
  o->penalty_max = 1e8;
  o->use_backtracking_linesearch = 0;
  o->linesearch_c1 = 1e-4;
  o->linesearch_c2 = 0.9;
}

altro_b200_solver* altro_b200_create(int horizon_length, int batch, int device) {
  if (horizon_length <= 0 || batch <= 0) return nullptr;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0 || device >= cnt) {
    fprintf(stderr, "altro_b200: no usable CUDA device; the solve path has no CPU fallback\n");
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  altro_b200_solver* s = new altro_b200_solver();
  s->N = horizon_length;
  s->B = batch;
  s->Bp = ((long)batch + 31) / 32 * 32;
  s->device = device;
  altro_b200_default_options(&s->opts);
  memset(&s->con_h, 0, sizeof(s->con_h));
  // test hook: run a whole test suite with the other Riccati schedule (results are bit-identical)
  if (const char* env = getenv("ALTRO_B200_BACKWARD_TEAM")) s->backward_team = atoi(env) != 0;
  if (const char* env = getenv("ALTRO_B200_FOLLOWER")) s->follow_deriv = atoi(env) != 0;
  if (const char* env = getenv("ALTRO_B200_QRC_UNIFORM")) s->qrc_uniform_enable = atoi(env) != 0;
  if (const char* env = getenv("ALTRO_B200_SPEC_ROUND1")) s->spec_round1 = atoi(env) != 0;
  if (const char* env = getenv("ALTRO_B200_FWD_DEPTH")) s->fwd_depth = std::max(2, std::min(8, atoi(env)));
  return s;
}

void altro_b200_destroy(altro_b200_solver* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  for (void* p : s->allocs) cudaFree(p);
  if (s->stage) cudaFree(s->stage);
  if (s->view_buf) cudaFree(s->view_buf);
  for (int i = 0; i < altro_b200_solver::kMaxSplit; ++i) {
    if (s->sub_ready[i]) {
      cudaStreamDestroy(s->sub_stream[i]);
      cudaFreeHost(s->sub_ph[i].h_done);
      cudaEventDestroy(s->sub_ph[i].ev0);
      cudaEventDestroy(s->sub_ph[i].ev1);
      for (int j = 0; j < PhaseHost::kDoneRing; ++j) cudaEventDestroy(s->sub_ph[i].ev_done[j]);
      cudaEventDestroy(s->ev_join[i]);
    }
  }
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  delete s;
}

int altro_b200_set_stream(altro_b200_solver* s, void* cuda_stream) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  s->stream = (cudaStream_t)cuda_stream;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_dimension(altro_b200_solver* s, int n, int m) {  // altro_solver.cpp:26-47
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (s->initialized || s->dims_set) return ALTRO_B200_SOLVER_ALREADY_INITIALIZED;
  if (n <= 0) return ALTRO_B200_STATE_DIM_UNKNOWN;
  if (m <= 0) return ALTRO_B200_INPUT_DIM_UNKNOWN;
  CUDA_OK(cudaSetDevice(s->device));
  s->n = n;
  s->m = m;
  const long N = s->N, S = s->Bp;
  s->Qd_h.assign((size_t)(N + 1) * n, 0.0);
  s->qrc_id.assign((size_t)N + 1, -1);
  s->Rd_h.assign((size_t)N * m, 0.0);
  s->Qf_h.assign((size_t)(N + 1) * n * n, 0.0);
  s->Rf_h.assign((size_t)N * m * m, 0.0);
  s->Hf_h.assign((size_t)N * m * n, 0.0);
  s->lin_h.assign((size_t)N * (n * n + n * m + n), 0.0);
  DALLOC(s, s->Qd, (N + 1) * n);
  DALLOC(s, s->Rd, N * m);
  {
    // main record stream; row order documented in device_problem.h
    s->G = (int)(S / 32);
    const long rows = 2L * (n + m) + (n + m + 1) + (m * n + m) + (n * n + n * m + n + m) + n +
                      (n * n + n) + m;
    s->R = rows * 32;
    s->GS = (N + 1) * s->R;
    DALLOC(s, s->rec, (long)s->G * s->GS);
    long row = 0;
    auto take = [&](int nrows) {
      double* ptr = s->rec + row * 32;
      row += nrows;
      return ptr;
    };
    s->xbar = take(n);
    s->ubar = take(m);
    s->q = take(n);
    s->r = take(m);
    s->c = take(1);
    s->K = take(m * n);
    s->d = take(m);
    s->x = take(n);
    s->u = take(m);
    s->A = take(n * n);
    s->Bm = take(n * m);
    s->lx = take(n);
    s->lu = take(m);
    s->y = take(n);
    s->P = take(n * n);
    s->p = take(n);
    s->u_init = take(m);
    DALLOC(s, s->x0, (long)s->G * n * 32);
  }
  DALLOC(s, s->rho, S);
  DALLOC(s, s->status, S);
  DALLOC(s, s->iters, S);
  DALLOC(s, s->merit_evals, S);
  DALLOC(s, s->ls_fail, S);
  DALLOC(s, s->phi, S);
  DALLOC(s, s->stat, S);
  DALLOC(s, s->feas, S);
  DALLOC(s, s->ls, S);
  DALLOC(s, s->alpha_eval, S);
  DALLOC(s, s->alpha_bt, S);
  DALLOC(s, s->sel, S);
  DALLOC(s, s->stat_acc, S);
  DALLOC(s, s->feas_acc, S);
  DALLOC(s, s->phi_eval, S);
  DALLOC(s, s->phi0, S);
  DALLOC(s, s->dphi0, S);
  DALLOC(s, s->flags, S);
  DALLOC(s, s->iter_count, S);
  DALLOC(s, s->d_done, altro_b200_solver::kMaxSplit);
  DALLOC(s, s->d_prof, 16 * altro_b200_solver::kMaxSplit);
  DALLOC(s, s->ls_hist, 32);
  memset(&s->ph, 0, sizeof(s->ph));
  {  // device limits that size the staging rings of the sequential sweeps (solve_inst.cu)
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, s->device);
    s->ph.num_sms = v > 0 ? v : 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, s->device);
    s->ph.smem_per_sm = v > 0 ? (size_t)v : (size_t)228 * 1024;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device);
    s->ph.smem_per_cta = v > 0 ? (size_t)v : (size_t)227 * 1024;
  }
  s->dims_set = true;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_time_step(altro_b200_solver* s, float h) {  // altro_solver.cpp:49-63
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (h <= 0.0f) return ALTRO_B200_TIMESTEP_NOT_POSITIVE;
  s->h = h;
  if (!s->h_h.empty()) std::fill(s->h_h.begin(), s->h_h.end(), h);
  s->h_uniform = true;
  return ALTRO_B200_NO_ERROR;
}

// SetTimeStep(h, k_start, k_stop) (altro_solver.cpp:49-63): knots [k_start, k_stop) of the horizon.
// While all N steps are equal the kernels use the scalar; otherwise a per-knot table on the device.
int altro_b200_set_time_step_range(altro_b200_solver* s, float h, int k_start, int k_stop) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  int e = resolve_range(s, k_start, k_stop, false);
  if (e) return e;
  if (h <= 0.0f) return ALTRO_B200_TIMESTEP_NOT_POSITIVE;
  if (empty_range(k_start, k_stop)) return ALTRO_B200_NO_ERROR;
  if (s->h_h.empty()) s->h_h.assign((size_t)s->N, s->h);  // s->h is 0 until a step has been set
  for (int k = k_start; k < k_stop && k < s->N; ++k) s->h_h[k] = h;
  s->h_uniform = true;
  for (int k = 1; k < s->N; ++k) s->h_uniform = s->h_uniform && s->h_h[k] == s->h_h[0];
  s->h = s->h_uniform ? s->h_h[0] : 0.0f;
  if (!s->h_uniform) {
    CUDA_OK(cudaSetDevice(s->device));
    if (!s->hk) DALLOC(s, s->hk, (long)s->N);
    CUDA_OK(cudaMemcpyAsync(s->hk, s->h_h.data(), sizeof(float) * s->N, cudaMemcpyHostToDevice, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
  }
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_model(altro_b200_solver* s, int model_id, const double* params, int nparams) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  double prm[8] = {0};
  for (int i = 0; i < 8 && i < nparams; ++i) prm[i] = params[i];
  if (!find_launcher(model_id, s->n, s->m, prm)) return ALTRO_B200_ERR_UNSUPPORTED;
  s->model = model_id;
  memcpy(s->params, prm, sizeof(prm));
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_linear_dynamics(altro_b200_solver* s, const double* A, const double* B,
                                   const double* f, int k_start, int k_stop) {
  if (!s || !A || !B) return ALTRO_B200_INVALID_POINTER;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  int e = resolve_range(s, k_start, k_stop, false);
  if (e) return e;
  if (empty_range(k_start, k_stop)) return ALTRO_B200_NO_ERROR;
  const int n = s->n, m = s->m;
  const int W = n * n + n * m + n;
  for (int k = k_start; k < k_stop; ++k) {
    double* T = &s->lin_h[(size_t)k * W];
    memcpy(T, A, sizeof(double) * n * n);
    memcpy(T + n * n, B, sizeof(double) * n * m);
    if (f) memcpy(T + n * n + n * m, f, sizeof(double) * n);
  }
  if (s->model < 0) s->model = MODEL_LINEAR;
  return ALTRO_B200_NO_ERROR;
}

// device copies of the dense cost blocks (only once a dense knot exists)
static int upload_dense_cost(altro_b200_solver* s) {
  if (!s->dense_cost) return 0;
  if (!s->Qf) {
    DALLOC(s, s->Qf, (long)s->Qf_h.size());
    DALLOC(s, s->Rf, (long)s->Rf_h.size());
    DALLOC(s, s->Hf, (long)s->Hf_h.size());
  }
  CUDA_OK(cudaMemcpyAsync(s->Qf, s->Qf_h.data(), s->Qf_h.size() * 8, cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaMemcpyAsync(s->Rf, s->Rf_h.data(), s->Rf_h.size() * 8, cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaMemcpyAsync(s->Hf, s->Hf_h.data(), s->Hf_h.size() * 8, cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaStreamSynchronize(s->stream));
  return 0;
}

// bookkeeping for DeviceProblem::qrc_uniform; same_for_all_knots: the call wrote identical q, r, c
// to every knot of [k0, k1) (per problem)
static void mark_qrc(altro_b200_solver* s, int k0, int k1, bool same_for_all_knots) {
  const int id = same_for_all_knots ? ++s->qrc_next : -1;
  for (int k = std::max(k0, 0); k < k1 && k < (int)s->qrc_id.size(); ++k) s->qrc_id[k] = id;
}
static int qrc_uniform(const altro_b200_solver* s) {
  if (s->qrc_id.empty() || s->qrc_id[0] <= 0) return 0;
  for (int k = 1; k < s->N; ++k)
    if (s->qrc_id[k] != s->qrc_id[0]) return 0;
  return 1;
}

static int store_weights(altro_b200_solver* s, const double* Qd, const double* Rd, int k0, int k1) {
  const int n = s->n, m = s->m;
  for (int k = k0; k < k1; ++k) {
    memcpy(&s->Qd_h[(size_t)k * n], Qd, sizeof(double) * n);
    double* Qk = &s->Qf_h[(size_t)k * n * n];
    memset(Qk, 0, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) Qk[i + n * i] = Qd[i];
    if (k < s->N && Rd) {
      memcpy(&s->Rd_h[(size_t)k * m], Rd, sizeof(double) * m);
      double* Rk = &s->Rf_h[(size_t)k * m * m];
      memset(Rk, 0, sizeof(double) * m * m);
      for (int i = 0; i < m; ++i) Rk[i + m * i] = Rd[i];
      memset(&s->Hf_h[(size_t)k * m * n], 0, sizeof(double) * m * n);
    }
  }
  if (int e = upload_dense_cost(s)) return e;
  CUDA_OK(cudaMemcpyAsync(s->Qd, s->Qd_h.data(), s->Qd_h.size() * 8, cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaMemcpyAsync(s->Rd, s->Rd_h.data(), s->Rd_h.size() * 8, cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaStreamSynchronize(s->stream));  // the host mirrors may change right after
  return 0;
}

int altro_b200_set_lqr_cost(altro_b200_solver* s, const double* Qd, const double* Rd,
                            const double* xref, const double* uref, int per_problem, int k_start,
                            int k_stop) {
  if (!s || !Qd || !Rd || !xref || !uref) return ALTRO_B200_INVALID_POINTER;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  CUDA_OK(cudaSetDevice(s->device));
  int e = resolve_range(s, k_start, k_stop, true);
  if (e) return e;
  if (empty_range(k_start, k_stop)) return ALTRO_B200_NO_ERROR;
  e = store_weights(s, Qd, Rd, k_start, k_stop);
  if (e) return e;
  const int n = s->n, m = s->m;
  const long cnt = per_problem ? (long)s->B * (n + m) : (n + m);
  e = ensure_stage(s, cnt);
  if (e) return e;
  double* xr = s->stage;
  double* ur = s->stage + (per_problem ? (long)s->B * n : n);
  CUDA_OK(cudaMemcpyAsync(xr, xref, sizeof(double) * (per_problem ? (size_t)s->B * n : n),
                          cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaMemcpyAsync(ur, uref, sizeof(double) * (per_problem ? (size_t)s->B * m : m),
                          cudaMemcpyHostToDevice, s->stream));
  dim3 grid((s->B + 127) / 128, (unsigned)(k_stop - k_start));
  k_lqr_cost<<<grid, 128, 0, s->stream>>>(n, m, s->N, s->B, k_start, k_stop, s->Qd, s->Rd, xr, ur,
                                          per_problem ? 1 : 0, nullptr, fview(s, s->q, n),
                                          fview(s, s->r, m), fview(s, s->c, 1));
  s->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(s->stream));
  // one weight set and a knot-independent reference: q, r, c do not depend on k inside the range
  // (c differs between a stage knot and the terminal knot, which is never part of the flag)
  mark_qrc(s, k_start, k_stop, true);
  s->cost_set = true;
  return ALTRO_B200_NO_ERROR;
}

static int apply_window(altro_b200_solver* s) {
  mark_qrc(s, 0, s->N + 1, false);
  dim3 grid((s->B + 127) / 128, (unsigned)(s->N + 1));
  k_lqr_cost<<<grid, 128, 0, s->stream>>>(s->n, s->m, s->N, s->B, 0, s->N + 1, s->Qd, s->Rd, s->xtab,
                                          s->utab, 2, s->offsets, fview(s, s->q, s->n),
                                          fview(s, s->r, s->m), fview(s, s->c, 1));
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int altro_b200_set_lqr_cost_window(altro_b200_solver* s, const double* Qd, const double* Rd,
                                   const double* xtab, const double* utab, int T,
                                   const int* offsets) {
  if (!s || !Qd || !Rd || !xtab || !utab || !offsets) return ALTRO_B200_INVALID_POINTER;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  CUDA_OK(cudaSetDevice(s->device));
  int e = store_weights(s, Qd, Rd, 0, s->N + 1);
  if (e) return e;
  if (T < s->N + 1) return ALTRO_B200_BAD_INDEX;
  int omin = offsets[0], omax = offsets[0];
  for (int b = 1; b < s->B; ++b) {
    omin = std::min(omin, offsets[b]);
    omax = std::max(omax, offsets[b]);
  }
  // every knot 0..N of every problem must read a row of the tables
  if (omin < 0 || omax + s->N >= T) return ALTRO_B200_BAD_INDEX;
  if (!s->xtab || T > s->T_alloc) {  // (re)allocate only when the tables grow; old ones are released
    if (s->xtab) {
      CUDA_OK(cudaStreamSynchronize(s->stream));
      release(s, s->xtab);
      release(s, s->utab);
    }
    DALLOC(s, s->xtab, (long)T * s->n);
    DALLOC(s, s->utab, (long)T * s->m);
    if (!s->offsets) DALLOC(s, s->offsets, s->Bp);
    s->T_alloc = T;
  }
  s->T = T;
  s->off_min = omin;
  s->off_max = omax;
  CUDA_OK(cudaMemcpyAsync(s->xtab, xtab, sizeof(double) * (size_t)T * s->n, cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaMemcpyAsync(s->utab, utab, sizeof(double) * (size_t)T * s->m, cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaMemcpyAsync(s->offsets, offsets, sizeof(int) * (size_t)s->B, cudaMemcpyHostToDevice, s->stream));
  e = apply_window(s);
  if (e) return e;
  CUDA_OK(cudaStreamSynchronize(s->stream));
  s->cost_set = true;
  return ALTRO_B200_NO_ERROR;
}

static int move_window(altro_b200_solver* s, int steps) {
  if (!s->xtab) return ALTRO_B200_COST_FUN_NOT_SET;
  // the moved window must stay inside the reference tables for every problem
  if (s->off_min + steps < 0 || s->off_max + steps + s->N >= s->T) return ALTRO_B200_BAD_INDEX;
  s->off_min += steps;
  s->off_max += steps;
  CUDA_OK(cudaSetDevice(s->device));
  k_add_int<<<(s->B + 127) / 128, 128, 0, s->stream>>>(s->offsets, s->B, steps);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int altro_b200_advance_window(altro_b200_solver* s, int steps) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  int e = move_window(s, steps);
  if (e) return e;
  return apply_window(s);
}

int altro_b200_advance_window_linear(altro_b200_solver* s, int steps, double c_u) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;  // UpdateLinearCosts, altro_solver.cpp:268
  int e = move_window(s, steps);
  if (e) return e;
  mark_qrc(s, 0, s->N + 1, false);
  dim3 grid((s->B + 127) / 128, (unsigned)(s->N + 1));
  k_window_linear_update<<<grid, 128, 0, s->stream>>>(s->n, s->N, s->B, s->Qd, s->xtab, s->offsets,
                                                      fview(s, s->q, s->n), fview(s, s->c, 1), c_u);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_mpc_cost_update(altro_b200_solver* s, int mode, double c_u) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (mode != 0 && mode != 1) return ALTRO_B200_BAD_INDEX;
  s->mpc_mode = mode;
  s->mpc_cu = c_u;
  return ALTRO_B200_NO_ERROR;
}

static int set_linear_terms(altro_b200_solver* s, const double* q, const double* r, const double* c,
                            int per_problem, int k0, int k1) {
  const int n = s->n, m = s->m;
  const int nk = k1 - k0;
  const int nku = (k1 > s->N ? s->N : k1) - k0;  // knots that carry an input
  int e = 0;
  mark_qrc(s, k0, k1, !per_problem && q && c && (r || nku <= 0));
  if (per_problem) {
    if (q) e = upload_pm(s, q, (long)nk * n, fview(s, s->q, n, k0));
    if (!e && r && nku > 0) {
      if (nku == nk) {
        e = upload_pm(s, r, (long)nk * m, fview(s, s->r, m, k0));
      } else {  // host rows are [B][nk][m] but only nku of them exist on the device
        std::vector<double> tmp((size_t)s->B * nku * m);
        for (int b = 0; b < s->B; ++b)
          memcpy(&tmp[(size_t)b * nku * m], r + (size_t)b * nk * m, sizeof(double) * nku * m);
        e = upload_pm(s, tmp.data(), (long)nku * m, fview(s, s->r, m, k0));
        if (!e) CUDA_OK(cudaStreamSynchronize(s->stream));
      }
    }
    if (!e && c) e = upload_pm(s, c, nk, fview(s, s->c, 1, k0));
  } else {
    std::vector<double> tmp;
    if (q) {
      tmp.resize((size_t)nk * n);
      for (int k = 0; k < nk; ++k) memcpy(&tmp[(size_t)k * n], q, sizeof(double) * n);
      e = upload_shared(s, tmp.data(), (long)nk * n, fview(s, s->q, n, k0));
      if (!e) CUDA_OK(cudaStreamSynchronize(s->stream));
    }
    if (!e && r && nku > 0) {
      tmp.resize((size_t)nku * m);
      for (int k = 0; k < nku; ++k) memcpy(&tmp[(size_t)k * m], r, sizeof(double) * m);
      e = upload_shared(s, tmp.data(), (long)nku * m, fview(s, s->r, m, k0));
      if (!e) CUDA_OK(cudaStreamSynchronize(s->stream));
    }
    if (!e && c) {
      tmp.assign((size_t)nk, c[0]);
      e = upload_shared(s, tmp.data(), nk, fview(s, s->c, 1, k0));
      if (!e) CUDA_OK(cudaStreamSynchronize(s->stream));
    }
  }
  return e;
}

int altro_b200_set_diagonal_cost(altro_b200_solver* s, const double* Qd, const double* Rd,
                                 const double* q, const double* r, const double* c,
                                 int per_problem, int k_start, int k_stop) {
  if (!s || !Qd || !q || !c) return ALTRO_B200_INVALID_POINTER;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  CUDA_OK(cudaSetDevice(s->device));
  int e = resolve_range(s, k_start, k_stop, true);
  if (e) return e;
  if (empty_range(k_start, k_stop)) return ALTRO_B200_NO_ERROR;
  e = store_weights(s, Qd, Rd, k_start, k_stop);
  if (e) return e;
  e = set_linear_terms(s, q, r, c, per_problem, k_start, k_stop);
  if (e) return e;
  s->cost_set = true;
  return ALTRO_B200_NO_ERROR;
}

// SetQuadraticCost (altro_solver.cpp:118-136, knotpoint_data.cpp:64-85): dense Q [n*n], R [m*m],
// H [m*n] (column-major, shared by the batch); q, r, c as in SetDiagonalCost.
int altro_b200_set_quadratic_cost(altro_b200_solver* s, const double* Q, const double* R,
                                  const double* H, const double* q, const double* r,
                                  const double* c, int per_problem, int k_start, int k_stop) {
  if (!s || !Q || !q || !c) return ALTRO_B200_INVALID_POINTER;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  CUDA_OK(cudaSetDevice(s->device));
  int e = resolve_range(s, k_start, k_stop, true);
  if (e) return e;
  if (empty_range(k_start, k_stop)) return ALTRO_B200_NO_ERROR;
  if (k_start < s->N && (!R || !r)) return ALTRO_B200_INVALID_POINTER;
  const int n = s->n, m = s->m;
  for (int k = k_start; k < k_stop; ++k) {
    memcpy(&s->Qf_h[(size_t)k * n * n], Q, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) s->Qd_h[(size_t)k * n + i] = Q[i + n * i];
    if (k < s->N) {
      memcpy(&s->Rf_h[(size_t)k * m * m], R, sizeof(double) * m * m);
      for (int i = 0; i < m; ++i) s->Rd_h[(size_t)k * m + i] = R[i + m * i];
      if (H)
        memcpy(&s->Hf_h[(size_t)k * m * n], H, sizeof(double) * m * n);
      else
        memset(&s->Hf_h[(size_t)k * m * n], 0, sizeof(double) * m * n);
    }
  }
  s->dense_cost = true;
  CUDA_OK(cudaMemcpyAsync(s->Qd, s->Qd_h.data(), s->Qd_h.size() * 8, cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaMemcpyAsync(s->Rd, s->Rd_h.data(), s->Rd_h.size() * 8, cudaMemcpyHostToDevice, s->stream));
  e = upload_dense_cost(s);
  if (e) return e;
  e = set_linear_terms(s, q, r, c, per_problem, k_start, k_stop);
  if (e) return e;
  s->cost_set = true;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_update_linear_costs(altro_b200_solver* s, const double* q, const double* r,
                                   const double* c, int per_problem, int k_start, int k_stop) {
  if (!s || !c) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;  // altro_solver.cpp:268
  CUDA_OK(cudaSetDevice(s->device));
  int e = resolve_range(s, k_start, k_stop, true);
  if (e) return e;
  if (empty_range(k_start, k_stop)) return ALTRO_B200_NO_ERROR;
  if (r && k_stop > s->N && k_stop - k_start == 1) return ALTRO_B200_INVALID_OPT_AT_TERMINAL;
  return set_linear_terms(s, q, r, c, per_problem, k_start, k_stop);
}

int altro_b200_set_constraint(altro_b200_solver* s, int cone, int dim, const int* idx,
                              const double* scale, const double* off, const double* off_b,
                              int k_start, int k_stop) {  // altro_solver.cpp:192-223
  if (!s || !idx || !scale || !off) return ALTRO_B200_INVALID_POINTER;
  if (s->initialized) return ALTRO_B200_SOLVER_ALREADY_INITIALIZED;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  if (dim <= 0 || dim > kMaxConDim) return ALTRO_B200_INVALID_CONSTRAINT_DIM;
  if (cone == CONE_SOC && dim > kMaxSocDim) return ALTRO_B200_INVALID_CONSTRAINT_DIM;
  if (s->con_h.ncon >= kMaxCon) return ALTRO_B200_MAX_CONSTRAINTS_EXCEEDED;
  CUDA_OK(cudaSetDevice(s->device));
  int e = resolve_range(s, k_start, k_stop, true);
  if (e) return e;
  if (empty_range(k_start, k_stop)) return ALTRO_B200_NO_ERROR;
  ConSlot& c = s->con_h.slot[s->con_h.ncon];
  memset(&c, 0, sizeof(c));
  c.k_start = k_start;
  c.k_stop = k_stop;
  c.cone = cone;
  c.dim = dim;
  c.row0 = s->con_h.rows;
  for (int i = 0; i < dim; ++i) {
    if (idx[i] < -1 || idx[i] >= s->n + s->m) return ALTRO_B200_BAD_INDEX;
    // a row that reads an input cannot live on the terminal knot (it has no input)
    if (idx[i] >= s->n && k_stop > s->N) return ALTRO_B200_INVALID_OPT_AT_TERMINAL;
    c.idx[i] = idx[i];
    c.scale[i] = scale[i];
    c.off[i] = off[i];
  }
  if (off_b) {
    double* dptr = nullptr;
    DALLOC(s, dptr, (long)dim * s->Bp);
    e = upload_pm(s, off_b, dim, gview(dptr, dim));
    if (e) return e;
    CUDA_OK(cudaStreamSynchronize(s->stream));
    c.off_per_problem = 1;
    c.off_b = dptr;
  }
  s->con_h.rows += dim;
  s->con_h.ncon += 1;
  return ALTRO_B200_NO_ERROR;
}

// shared part of the family setters: slot bookkeeping + optional per-problem parameters
static int add_slot(altro_b200_solver* s, int cone, int dim, int family, int nparam,
                    const double* param_b, int k_start, int k_stop, ConSlot** out) {
  if (s->initialized) return ALTRO_B200_SOLVER_ALREADY_INITIALIZED;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  if (dim <= 0 || dim > kMaxConDim) return ALTRO_B200_INVALID_CONSTRAINT_DIM;
  if (cone == CONE_SOC && dim > kMaxSocDim) return ALTRO_B200_INVALID_CONSTRAINT_DIM;
  if (cone < CONE_EQUALITY || cone > CONE_SOC) return ALTRO_B200_BAD_INDEX;
  if (s->con_h.ncon >= kMaxCon) return ALTRO_B200_MAX_CONSTRAINTS_EXCEEDED;
  CUDA_OK(cudaSetDevice(s->device));
  int e = resolve_range(s, k_start, k_stop, true);
  if (e) return e;
  if (empty_range(k_start, k_stop)) {
    *out = nullptr;
    return ALTRO_B200_NO_ERROR;
  }
  ConSlot& c = s->con_h.slot[s->con_h.ncon];
  memset(&c, 0, sizeof(c));
  c.k_start = k_start;
  c.k_stop = k_stop;
  c.cone = cone;
  c.dim = dim;
  c.family = family;
  c.row0 = s->con_h.rows;
  if (param_b) {
    double* dptr = nullptr;
    DALLOC(s, dptr, (long)nparam * s->Bp);
    e = upload_pm(s, param_b, nparam, gview(dptr, nparam));
    if (e) return e;
    CUDA_OK(cudaStreamSynchronize(s->stream));
    c.off_per_problem = 1;
    c.off_b = dptr;
  }
  *out = &c;
  return ALTRO_B200_NO_ERROR;
}
static void commit_slot(altro_b200_solver* s, const ConSlot& c) {
  s->con_h.rows += c.dim;
  s->con_h.ncon += 1;
}

// SetConstraint with general affine rows c = J [x;u] + e (dense J [dim x (n+m)], column-major,
// shared by the batch; e [dim] shared or e_b [B][dim] per problem).
int altro_b200_set_constraint_affine(altro_b200_solver* s, int cone, int dim, const double* J,
                                     const double* e, const double* e_b, int k_start,
                                     int k_stop) {
  if (!s || !J || !e) return ALTRO_B200_INVALID_POINTER;
  ConSlot* c = nullptr;
  int err = add_slot(s, cone, dim, CON_FAMILY_AFFINE, dim, e_b, k_start, k_stop, &c);
  if (err || !c) return err;
  const int nm = s->n + s->m;
  // rows that read an input cannot live on the terminal knot (it has no input)
  if (c->k_stop > s->N)
    for (int j = s->n; j < nm; ++j)
      for (int i = 0; i < dim; ++i)
        if (J[i + dim * j] != 0.0) return ALTRO_B200_INVALID_OPT_AT_TERMINAL;
  double* Jd = nullptr;
  DALLOC(s, Jd, (long)dim * nm);
  CUDA_OK(cudaMemcpyAsync(Jd, J, sizeof(double) * (size_t)dim * nm, cudaMemcpyHostToDevice, s->stream));
  CUDA_OK(cudaStreamSynchronize(s->stream));
  c->Jd = Jd;
  for (int i = 0; i < dim; ++i) c->off[i] = e[i];
  commit_slot(s, *c);
  return ALTRO_B200_NO_ERROR;
}

// SetConstraint with the nonlinear keep-out disc c = r^2 - (v_a - cx)^2 - (v_b - cy)^2 <= 0 on
// the variables v_a = [x;u][idx_a], v_b = [x;u][idx_b]; (cx, cy, r) shared or per problem [B][3].
int altro_b200_set_constraint_disc(altro_b200_solver* s, int idx_a, int idx_b, const double* disc,
                                   const double* disc_b, int k_start, int k_stop) {
  if (!s || !disc) return ALTRO_B200_INVALID_POINTER;
  if (s->dims_set && (idx_a < 0 || idx_b < 0 || idx_a >= s->n + s->m || idx_b >= s->n + s->m))
    return ALTRO_B200_BAD_INDEX;
  ConSlot* c = nullptr;
  int err = add_slot(s, CONE_INEQUALITY, 1, CON_FAMILY_DISC, 3, disc_b, k_start, k_stop, &c);
  if (err || !c) return err;
  if (c->k_stop > s->N && (idx_a >= s->n || idx_b >= s->n)) return ALTRO_B200_INVALID_OPT_AT_TERMINAL;
  c->idx[0] = idx_a;
  c->idx[1] = idx_b;
  for (int i = 0; i < 3; ++i) c->off[i] = disc[i];
  commit_slot(s, *c);
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_initial_state(altro_b200_solver* s, const double* x0, int per_problem) {
  if (!s || !x0) return ALTRO_B200_INVALID_POINTER;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  CUDA_OK(cudaSetDevice(s->device));
  int e = per_problem ? upload_pm(s, x0, s->n, gview(s->x0, s->n)) : upload_shared(s, x0, s->n, gview(s->x0, s->n));
  if (e) return e;
  CUDA_OK(cudaStreamSynchronize(s->stream));
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_initialize(altro_b200_solver* s) {  // altro_solver.cpp:225-229
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (s->initialized) return ALTRO_B200_SOLVER_ALREADY_INITIALIZED;
  if (!s->dims_set) return ALTRO_B200_STATE_DIM_UNKNOWN;          // knotpoint_data.cpp:242-246
  if (s->h_uniform) {
    if (!(s->h > 0.0f)) return ALTRO_B200_TIMESTEP_NOT_POSITIVE;    // :259-263
  } else {
    for (int k = 0; k < s->N; ++k)
      if (!(s->h_h[k] > 0.0f)) return ALTRO_B200_TIMESTEP_NOT_POSITIVE;
  }
  if (s->model < 0) return ALTRO_B200_DYNAMICS_FUN_NOT_SET;       // :264-269
  if (!s->cost_set) return ALTRO_B200_COST_FUN_NOT_SET;           // :271-275
  if (!find_launcher(s->model, s->n, s->m, s->params)) return ALTRO_B200_ERR_UNSUPPORTED;
  CUDA_OK(cudaSetDevice(s->device));
  const long rows = s->con_h.rows;
  if (rows > 0) {
    s->Rz = 2 * rows * 32;
    s->GSz = (long)(s->N + 1) * s->Rz;
    DALLOC(s, s->zrec, (long)s->G * s->GSz);
    s->z = s->zrec;
    s->zest = s->zrec + rows * 32;
    DALLOC(s, s->con_d, 1);
    CUDA_OK(cudaMemcpyAsync(s->con_d, &s->con_h, sizeof(ConTable), cudaMemcpyHostToDevice, s->stream));
  }
  if (s->model == MODEL_LINEAR) {
    DALLOC(s, s->lin, (long)s->lin_h.size());
    CUDA_OK(cudaMemcpyAsync(s->lin, s->lin_h.data(), s->lin_h.size() * 8, cudaMemcpyHostToDevice, s->stream));
  }
  k_fill<<<64, 256, 0, s->stream>>>(s->rho, s->Bp, 1.0);  // rho_ = 1.0, knotpoint_data.cpp:343
  s->launches++;
  CUDA_OK(cudaGetLastError());
  // candidate slots of the speculative line search
  s->Rs = (long)(s->n + s->m) * 32;
  {
    // candidate slot buffers: as many as fit a budget of 8 GB (at least 3)
    const long per_slot = (long)s->G * (s->N + 1) * s->Rs;
    const long fit = ((long)8 << 30) / 8 / per_slot;
    if (s->nstore > s->nslots - 1) s->nstore = s->nslots - 1;
    if (s->nstore > fit) s->nstore = (int)(fit < 3 ? 3 : fit);
    if (s->nstore < 1) s->nstore = 1;
  }
  s->nstore_alloc = s->nstore;
  DALLOC(s, s->xs, (long)s->nstore * s->G * (s->N + 1) * s->Rs);
  s->us = s->xs + (long)s->n * 32;
  DALLOC(s, s->phi_s, (long)(kMaxHalvings + 1) * s->Bp);
  DALLOC(s, s->spec_base, s->Bp);
  DALLOC(s, s->spec_known, s->Bp);
  CUDA_OK(cudaStreamSynchronize(s->stream));
  s->initialized = true;
  return ALTRO_B200_NO_ERROR;
}

static int set_traj(altro_b200_solver* s, const double* v, int layout, int k0, int k1, int E,
                    double* dst) {
  const int nk = k1 - k0;
  const FieldView base = fview(s, dst, E, k0);
  int e = 0;
  if (layout == 2) {
    e = upload_pm(s, v, (long)nk * E, base);
  } else {
    std::vector<double> tmp((size_t)nk * E);
    for (int k = 0; k < nk; ++k)
      memcpy(&tmp[(size_t)k * E], layout == 1 ? v + (size_t)k * E : v, sizeof(double) * E);
    e = upload_shared(s, tmp.data(), (long)nk * E, base);
  }
  if (e) return e;
  CUDA_OK(cudaStreamSynchronize(s->stream));
  return 0;
}

int altro_b200_set_input(altro_b200_solver* s, const double* u, int layout, int k_start,
                         int k_stop) {  // altro_solver.cpp:242-251
  if (!s || !u) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  CUDA_OK(cudaSetDevice(s->device));
  int e = resolve_range(s, k_start, k_stop, false);
  if (e) return e;
  if (empty_range(k_start, k_stop)) return ALTRO_B200_NO_ERROR;
  e = set_traj(s, u, layout, k_start, k_stop, s->m, s->u);
  if (e) return e;
  const long W = (long)(k_stop - k_start) * s->m;
  dim3 grid((unsigned)((s->B + 127) / 128), (unsigned)(W < 1024 ? W : 1024));
  k_copy_field<<<grid, 128, 0, s->stream>>>(fview(s, s->u, s->m, k_start), fview(s, s->u_init, s->m, k_start),
                                            s->B, W);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_reset_trajectory(altro_b200_solver* s) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  CUDA_OK(cudaSetDevice(s->device));
  const long W = (long)s->N * s->m;
  dim3 grid((unsigned)((s->B + 127) / 128), (unsigned)(W < 1024 ? W : 1024));
  k_copy_field<<<grid, 128, 0, s->stream>>>(fview(s, s->u_init, s->m), fview(s, s->u, s->m), s->B, W);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_state(altro_b200_solver* s, const double* x, int layout, int k_start,
                         int k_stop) {  // altro_solver.cpp:231-240
  if (!s || !x) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  CUDA_OK(cudaSetDevice(s->device));
  int e = resolve_range(s, k_start, k_stop, true);
  if (e) return e;
  if (empty_range(k_start, k_stop)) return ALTRO_B200_NO_ERROR;
  return set_traj(s, x, layout, k_start, k_stop, s->n, s->x);
}

int altro_b200_set_options(altro_b200_solver* s, const altro_b200_options* o) {
  if (!s || !o) return ALTRO_B200_INVALID_POINTER;
  s->opts = *o;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_reset_duals(altro_b200_solver* s) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  CUDA_OK(cudaSetDevice(s->device));
  if (s->zrec) CUDA_OK(cudaMemsetAsync(s->zrec, 0, (size_t)s->G * s->GSz * 8, s->stream));
  k_fill<<<64, 256, 0, s->stream>>>(s->rho, s->Bp, 1.0);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_shift_trajectory(altro_b200_solver* s) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  CUDA_OK(cudaSetDevice(s->device));
  k_shift<<<(s->B + 127) / 128, 128, 0, s->stream>>>(s->n, s->m, s->N, s->B, fview(s, s->x, s->n),
                                                     fview(s, s->u, s->m));
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return ALTRO_B200_NO_ERROR;
}

// TrajSolver's constraint level: 0 none, 1 linear cones only, 2 with second-order cones
static int con_level(const altro_b200_solver* s) {
  int level = s->con_h.ncon > 0 ? 1 : 0;
  for (int j = 0; j < s->con_h.ncon; ++j)
    if (s->con_h.slot[j].cone == CONE_SOC || s->con_h.slot[j].family != CON_FAMILY_SELECTOR) level = 2;
  if (s->dense_cost) level = 2;  // dense Q, R, H live in the general instantiation
  return level;
}

// One receding-horizon step without a host round trip (test/bicycle_test.cpp:302-337; the plant
// there is the model itself, :312): the tracking window moves one row with the reference's cost
// update -- UpdateLinearCosts(q, nullptr, c): q and c follow the window, r stays, c's input part
// is the frozen c_u (:317-328; altro_b200_set_mpc_cost_update selects the full re-windowing
// instead) --, the next initial state is x_[1] = f(x0, u_[0]) of the solved trajectory (:312,
// :331), then ShiftTrajectory (:334).  Duals and penalties carry over (quirk Q14).
int altro_b200_mpc_step(altro_b200_solver* s) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  // refuse BEFORE touching anything when the moved window would leave the reference tables
  if (s->xtab && (s->off_min + 1 < 0 || s->off_max + 1 + s->N >= s->T)) return ALTRO_B200_BAD_INDEX;
  CUDA_OK(cudaSetDevice(s->device));
  dim3 grid((unsigned)((s->B + 127) / 128), (unsigned)s->n);
  k_copy_field<<<grid, 128, 0, s->stream>>>(fview(s, s->x, s->n, 1), gview(s->x0, s->n), s->B, s->n);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  int e = altro_b200_shift_trajectory(s);
  if (e) return e;
  if (s->xtab)
    e = s->mpc_mode == 1 ? altro_b200_advance_window_linear(s, 1, s->mpc_cu) : altro_b200_advance_window(s, 1);
  return e;
}

static int run_host_op(altro_b200_solver* s, int op, double* cost_out);

static void fill_device_problem(const altro_b200_solver* s, DeviceProblem& P) {
  memset(&P, 0, sizeof(P));
  P.N = s->N;
  P.B = s->B;
  P.Bp = s->Bp;
  P.G = s->G;
  P.Gtot = s->G;
  P.R = s->R;
  P.GS = s->GS;
  P.Rz = s->Rz;
  P.GSz = s->GSz;
  P.Rs = s->Rs;
  P.zrows = s->con_h.rows;
  P.h = s->h;
  P.hk = s->h_uniform ? nullptr : s->hk;
  memcpy(P.model_params, s->params, sizeof(P.model_params));
  P.lin = s->lin;
  P.Qd = s->Qd;
  P.Rd = s->Rd;
  P.Qf = s->dense_cost ? s->Qf : nullptr;
  P.Rf = s->dense_cost ? s->Rf : nullptr;
  P.Hf = s->dense_cost ? s->Hf : nullptr;
  P.q = s->q;
  P.r = s->r;
  P.c = s->c;
  P.x0 = s->x0;
  P.xbar = s->xbar;
  P.ubar = s->ubar;
  P.x = s->x;
  P.u = s->u;
  P.y = s->y;
  P.A = s->A;
  P.Bm = s->Bm;
  P.lx = s->lx;
  P.lu = s->lu;
  P.K = s->K;
  P.d = s->d;
  P.P = s->P;
  P.p = s->p;
  P.contab = s->con_h;
  P.z = s->z;
  P.zest = s->zest;
  P.rho = s->rho;
  P.status = s->status;
  P.iters = s->iters;
  P.merit_evals = s->merit_evals;
  P.ls_fail = s->ls_fail;
  P.phi = s->phi;
  P.stat = s->stat;
  P.feas = s->feas;
  P.ls = s->ls;
  P.alpha_eval = s->alpha_eval;
  P.alpha_bt = s->alpha_bt;
  P.nslots = s->nslots;
  P.nstore = s->nslots > 1 ? s->nstore : 0;
  P.follow_deriv = s->follow_deriv;
  P.spec_round1 = s->spec_round1;
  P.prof_tid = getenv("ALTRO_B200_PROF_TID") ? atoi(getenv("ALTRO_B200_PROF_TID")) : 0;
  P.qrc_uniform = s->qrc_uniform_enable ? qrc_uniform(s) : 0;
  P.xs = s->xs;
  P.us = s->us;
  P.phi_s = s->phi_s;
  P.spec_base = s->spec_base;
  P.spec_known = s->spec_known;
  P.sel = s->sel;
  P.stat_acc = s->stat_acc;
  P.feas_acc = s->feas_acc;
  P.phi_eval = s->phi_eval;
  P.phi0 = s->phi0;
  P.dphi0 = s->dphi0;
  P.flags = s->flags;
  P.iter_count = s->iter_count;
  P.prof = nullptr;
  P.ls_hist = s->ls_hist;
  P.opts.iterations_max = s->opts.iterations_max;
  P.opts.tol_primal_feasibility = s->opts.tol_primal_feasibility;
  P.opts.tol_stationarity = s->opts.tol_stationarity;
  P.opts.tol_meritfun_gradient = s->opts.tol_meritfun_gradient;
  P.opts.penalty_initial = s->opts.penalty_initial;
  P.opts.penalty_scaling = s->opts.penalty_scaling;
  P.opts.penalty_max = s->opts.penalty_max;
  P.opts.use_backtracking_linesearch = s->opts.use_backtracking_linesearch;
  P.opts.ls_c1 = s->opts.linesearch_c1;
  P.opts.ls_c2 = s->opts.linesearch_c2;
}

// per-sub-batch stream, pinned stop-counter ring and events, created on first use
static int ensure_sub(altro_b200_solver* s, int i) {
  if (s->sub_ready[i]) return 0;
  PhaseHost& H = s->sub_ph[i];
  memset(&H, 0, sizeof(PhaseHost));
  CUDA_OK(cudaStreamCreateWithFlags(&s->sub_stream[i], cudaStreamNonBlocking));
  CUDA_OK(cudaMallocHost((void**)&H.h_done, PhaseHost::kDoneRing * sizeof(int)));
  CUDA_OK(cudaEventCreate(&H.ev0));
  CUDA_OK(cudaEventCreate(&H.ev1));
  for (int j = 0; j < PhaseHost::kDoneRing; ++j)
    CUDA_OK(cudaEventCreateWithFlags(&H.ev_done[j], cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&s->ev_join[i], cudaEventDisableTiming));
  s->sub_ready[i] = true;
  return 0;
}

int altro_b200_solve_async(altro_b200_solver* s) {  // altro_solver.cpp:257-260
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  CUDA_OK(cudaSetDevice(s->device));
  solve_launcher L = find_launcher(s->model, s->n, s->m, s->params);
  if (!L) return ALTRO_B200_ERR_UNSUPPORTED;
  DeviceProblem P;
  fill_device_problem(s, P);
  const int has_con = con_level(s);
  if (s->solve_mode == 1) {
    int e = L(P, has_con, s->stream, nullptr);
    s->launches++;
    if (e) return ALTRO_B200_ERR_NO_DEVICE;
    return ALTRO_B200_NO_ERROR;
  }
  long before = 0;  // kernel launches so far (the PH_FWD_* entries are sub-phases, not launches)
  for (int i = 0; i <= PH_FORWARD; ++i) before += s->ph.launches[i];
  int nsplit = s->nsplit;
  // DESIGN.md "pipelined sub-batches": the Riccati sweep keeps one warp per group busy, the
  // forward kernel up to six; sub-batches on separate streams let the kernels of different
  // iterations overlap (bicycle 16384, one B200: 45.3 ms with 1, 37.8 with 4, 37.3 with 8, 38.0 with
  // 16).  It only pays when the batch does not fit the machine at once -- two line-search CTAs per
  // SM; below that every split just adds launches to a latency-bound chain (scotty 8192 = 256
  // groups: 57.2 ms with 1, 62.0 with 4)
  if (nsplit <= 0) nsplit = s->G > 2 * s->ph.num_sms ? 8 : 1;
  nsplit = std::min(std::min(nsplit, (int)altro_b200_solver::kMaxSplit), s->G);
  const int per = (s->G + nsplit - 1) / nsplit;
  nsplit = (s->G + per - 1) / per;  // no empty sub-batch
  if (!s->ev_fork) CUDA_OK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
  CUDA_OK(cudaEventRecord(s->ev_fork, s->stream));
  struct Sub {
    DeviceProblem P;
    cudaStream_t st;
    PhaseHost* H;
    int total;
    bool done;
  } subs[altro_b200_solver::kMaxSplit];
  int e = 0;
  for (int i = 0; i < nsplit; ++i) {
    int err = ensure_sub(s, i);
    if (err) return err;
    Sub& sb = subs[i];
    sb.P = P;
    sb.P.g0 = i * per;
    sb.P.G = std::min(per, s->G - sb.P.g0);
    sb.st = s->sub_stream[i];
    sb.H = &s->sub_ph[i];
    sb.total = std::min(s->B, (sb.P.g0 + sb.P.G) * 32) - sb.P.g0 * 32;
    sb.done = false;
    PhaseHost& H = *sb.H;
    H.profile = s->ph.profile;
    H.num_sms = s->ph.num_sms;
    H.smem_per_sm = s->ph.smem_per_sm;
    H.smem_per_cta = s->ph.smem_per_cta;
    H.d_done = s->d_done + i;
    H.d_prof = s->d_prof + 16 * i;
    H.fwd_warps = std::max(4, s->nslots);
    H.fwd_depth = s->fwd_depth;
    H.backward_team = s->backward_team >= 0 ? s->backward_team : (s->n > kUnrollDim ? 1 : 0);
    for (int j = 0; j < PH_COUNT; ++j) {
      H.ms[j] = 0.0;
      H.launches[j] = 0;
      H.units[j] = 0.0;
    }
    H.syncs = 0;
    CUDA_OK(cudaStreamWaitEvent(sb.st, s->ev_fork, 0));
    H.op = OP_SOLVE_PROLOGUE;
    const int rc = L(sb.P, has_con, sb.st, &H);
    if (rc) e = rc;
  }
  // Iterations are enqueued back to back; the stop counter of iteration i is copied to pinned
  // memory right behind it and looked at kLag iterations later, so the device always has work
  // queued while the host decides whether the sub-batch still needs another iteration.
  constexpr int kLag = 2;
  int live = nsplit;
  for (int iter = 0; iter < s->opts.iterations_max && live > 0 && !e; ++iter) {
    for (int i = 0; i < nsplit; ++i) {
      Sub& sb = subs[i];
      if (sb.done) continue;
      PhaseHost& H = *sb.H;
      if (iter >= kLag) {
        const int slot = (iter - kLag) % PhaseHost::kDoneRing;
        CUDA_OK(cudaEventSynchronize(H.ev_done[slot]));
        H.syncs += 1;
        if (H.h_done[slot] >= sb.total) {
          sb.done = true;
          live -= 1;
          continue;
        }
      }
      H.op = OP_SOLVE_ITERATION;
      H.iter = iter;
      const int rc = L(sb.P, has_con, sb.st, &H);
      if (rc) e = rc;
      const int slot = iter % PhaseHost::kDoneRing;
      CUDA_OK(cudaMemcpyAsync(&H.h_done[slot], H.d_done, sizeof(int), cudaMemcpyDeviceToHost, sb.st));
      CUDA_OK(cudaEventRecord(H.ev_done[slot], sb.st));
    }
  }
  for (int i = 0; i < nsplit; ++i) {
    CUDA_OK(cudaEventRecord(s->ev_join[i], subs[i].st));
    CUDA_OK(cudaStreamWaitEvent(s->stream, s->ev_join[i], 0));
    for (int j = 0; j < PH_COUNT; ++j) {
      s->ph.ms[j] += s->sub_ph[i].ms[j];
      s->ph.launches[j] += s->sub_ph[i].launches[j];
      s->ph.units[j] += s->sub_ph[i].units[j];
    }
    s->ph.syncs += s->sub_ph[i].syncs;
  }
  long after = 0;
  for (int i = 0; i <= PH_FORWARD; ++i) after += s->ph.launches[i];
  s->launches += after - before;
  if (e) {
    fprintf(stderr, "altro_b200: CUDA error %d in the phase pipeline\n", e);
    return ALTRO_B200_ERR_NO_DEVICE;
  }
  return ALTRO_B200_NO_ERROR;
}

static int run_host_op(altro_b200_solver* s, int op, double* cost_out) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  CUDA_OK(cudaSetDevice(s->device));
  solve_launcher L = find_launcher(s->model, s->n, s->m, s->params);
  if (!L) return ALTRO_B200_ERR_UNSUPPORTED;
  DeviceProblem P;
  fill_device_problem(s, P);
  s->ph.op = op;
  s->ph.cost_out = cost_out;
  int e = L(P, con_level(s), s->stream, &s->ph);
  s->ph.op = OP_SOLVE_PROLOGUE;
  s->launches++;
  if (e) return ALTRO_B200_ERR_NO_DEVICE;
  return ALTRO_B200_NO_ERROR;
}

// The per-knot expansions of KnotPointData at the WORKING trajectory x_, u_ (as set by SetState /
// SetInput or left by Solve), for every knot of every problem: CalcDynamicsExpansion,
// CalcConstraints, CalcConstraintJacobians, CalcProjectedDuals (z_est = z - rho c), CalcCostGradient
// (knotpoint_data.cpp:406-437, :473-487, :523-595) with the current duals and penalty.  Results are
// read back through altro_b200_get_field ("A" "B" "lx" "lu" "z_est" and the derived views).
int altro_b200_knot_eval(altro_b200_solver* s) { return run_host_op(s, OP_KNOT_EVAL, nullptr); }

// KnotPointData::SetPenalty (knotpoint_data.cpp:180-191) for every constraint of every problem
int altro_b200_set_penalty(altro_b200_solver* s, double rho) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  if (!(rho > 0.0)) return ALTRO_B200_NON_POSITIVE_PENALTY;
  CUDA_OK(cudaSetDevice(s->device));
  k_fill<<<64, 256, 0, s->stream>>>(s->rho, s->Bp, rho);
  s->launches++;
  CUDA_OK(cudaGetLastError());
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_open_loop_rollout(altro_b200_solver* s) {  // altro_solver.cpp:253
  return run_host_op(s, OP_OPEN_LOOP_ROLLOUT, nullptr);
}

int altro_b200_calc_cost(altro_b200_solver* s, double* cost) {  // altro_solver.cpp:313
  if (!cost) return ALTRO_B200_INVALID_POINTER;
  int e = run_host_op(s, OP_CALC_COST, s ? s->phi_eval : nullptr);
  if (e) return e;
  CUDA_OK(cudaMemcpyAsync(cost, s->phi_eval, sizeof(double) * (size_t)s->B, cudaMemcpyDeviceToHost, s->stream));
  CUDA_OK(cudaStreamSynchronize(s->stream));
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_solve_mode(altro_b200_solver* s, int mode) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (mode != 0 && mode != 1) return ALTRO_B200_BAD_INDEX;
  s->solve_mode = mode;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_backward_mode(altro_b200_solver* s, int team) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (team < -1 || team > 1) return ALTRO_B200_BAD_INDEX;
  s->backward_team = team;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_pipeline_split(altro_b200_solver* s, int nsplit) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (nsplit < 0 || nsplit > altro_b200_solver::kMaxSplit) return ALTRO_B200_BAD_INDEX;
  s->nsplit = nsplit;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_speculation(altro_b200_solver* s, int nslots) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (s->initialized) return ALTRO_B200_SOLVER_ALREADY_INITIALIZED;
  if (nslots < 1 || nslots > 8) return ALTRO_B200_BAD_INDEX;
  s->nslots = nslots;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_candidate_store(altro_b200_solver* s, int nstore) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  if (nstore < 0) return ALTRO_B200_BAD_INDEX;
  s->nstore = std::min(nstore, s->nstore_alloc);
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_set_profiling(altro_b200_solver* s, int on) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  s->ph.profile = on != 0;
  for (int i = 0; i < PH_COUNT; ++i) {
    s->ph.ms[i] = 0.0;
    s->ph.launches[i] = 0;
    s->ph.units[i] = 0.0;
  }
  s->ph.syncs = 0;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_get_phase_stats(altro_b200_solver* s, double* ms, long* launches, double* units,
                               long* syncs) {
  if (!s || !ms || !launches || !units) return ALTRO_B200_INVALID_POINTER;
  for (int i = 0; i < PH_COUNT; ++i) {
    ms[i] = s->ph.ms[i];
    launches[i] = s->ph.launches[i];
    units[i] = s->ph.units[i];
  }
  if (syncs) *syncs = s->ph.syncs;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_synchronize(altro_b200_solver* s) {
  if (!s) return ALTRO_B200_INVALID_POINTER;
  CUDA_OK(cudaSetDevice(s->device));
  CUDA_OK(cudaStreamSynchronize(s->stream));
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_solve(altro_b200_solver* s) {
  int e = altro_b200_solve_async(s);
  if (e) return e;
  return altro_b200_synchronize(s);
}

long altro_b200_kernel_launches(const altro_b200_solver* s) { return s ? s->launches : 0; }

int altro_b200_get_linesearch_histogram(altro_b200_solver* s, long* hist32, int reset) {
  if (!s || !hist32) return ALTRO_B200_INVALID_POINTER;
  if (!s->dims_set) return ALTRO_B200_DIMENSION_UNKNOWN;
  CUDA_OK(cudaSetDevice(s->device));
  unsigned long long h[32];
  CUDA_OK(cudaMemcpyAsync(h, s->ls_hist, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
  CUDA_OK(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < 32; ++i) hist32[i] = (long)h[i];
  if (reset) CUDA_OK(cudaMemsetAsync(s->ls_hist, 0, sizeof(h), s->stream));
  return ALTRO_B200_NO_ERROR;
}

#define GETTER_PM(name, field, rows_expr, width_expr)                                 \
  int name(altro_b200_solver* s, double* out) {                            \
    if (!s || !out) return ALTRO_B200_INVALID_POINTER;                     \
    if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;         \
    CUDA_OK(cudaSetDevice(s->device));                                     \
    return download_pm(s, fview(s, s->field, (rows_expr)), (long)(width_expr), out); \
  }
GETTER_PM(altro_b200_get_states, x, s->n, (s->N + 1) * s->n)
GETTER_PM(altro_b200_get_inputs, u, s->m, s->N * s->m)
GETTER_PM(altro_b200_get_dual_dynamics, y, s->n, (s->N + 1) * s->n)
GETTER_PM(altro_b200_get_feedback_gains, K, s->m * s->n, s->N * s->m * s->n)
GETTER_PM(altro_b200_get_feedforward_gains, d, s->m, s->N * s->m)

static int run_host_op(altro_b200_solver* s, int op, double* cost_out);

// scratch stream [G][N+1][rows][32] for the derived views, grown on demand and kept
static int ensure_view(altro_b200_solver* s, long rows) {
  const long need = (long)s->G * (s->N + 1) * rows * 32;
  if (need <= s->view_count) return 0;
  if (s->view_buf) {
    CUDA_OK(cudaStreamSynchronize(s->stream));
    CUDA_OK(cudaFree(s->view_buf));
    s->bytes -= s->view_count * 8;
  }
  CUDA_OK(cudaMalloc((void**)&s->view_buf, (size_t)need * 8));
  s->view_count = need;
  s->bytes += need * 8;
  return 0;
}

// KnotPointData member of every problem by name (knotpoint_data.hpp:160-233).  Stored members:
// x u y (x_, u_, y_: the accepted point after Solve), xbar ubar (x, u), A B, lx lu, K d, P p, q r c,
// z z_est (duals and estimates of all constraints of the knot, slot after slot).  Re-created on
// demand: constraint_val z_proj (same row layout as z), lxx luu lux (CalcCostHessian at the working
// trajectory), rho (uniform over the problem's constraints).  out: [B][knots][rows] with knots =
// N + 1 (fields that do not exist at the terminal knot, and rows of constraints that do not apply
// at a knot, hold zeros / stale data there).
int altro_b200_get_field(altro_b200_solver* s, const char* name, double* out, int* rows_out) {
  if (!s || !name) return ALTRO_B200_INVALID_POINTER;
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  const int n = s->n, m = s->m, zrows = s->con_h.rows;
  struct { const char* nm; double* p; int rows; } tab[] = {
      {"x", s->x, n}, {"u", s->u, m}, {"y", s->y, n}, {"xbar", s->xbar, n}, {"ubar", s->ubar, m},
      {"A", s->A, n * n}, {"B", s->Bm, n * m}, {"lx", s->lx, n}, {"lu", s->lu, m},
      {"K", s->K, m * n}, {"d", s->d, m}, {"P", s->P, n * n}, {"p", s->p, n},
      {"q", s->q, n}, {"r", s->r, m}, {"c", s->c, 1}};
  for (auto& t : tab) {
    if (strcmp(t.nm, name) == 0) {
      if (rows_out) *rows_out = t.rows;
      if (!out) return ALTRO_B200_NO_ERROR;
      CUDA_OK(cudaSetDevice(s->device));
      if (t.p == s->A || t.p == s->Bm) {
        // [A B] may be stored packed (models.cuh, JacPack): expand into the dense scratch stream
        const long rows = (long)n * n + (long)n * m;
        int e = ensure_view(s, rows);
        if (e) return e;
        CUDA_OK(cudaMemsetAsync(s->view_buf, 0, (size_t)s->G * (s->N + 1) * rows * 32 * 8, s->stream));
        e = run_host_op(s, OP_UNPACK_JAC, s->view_buf);
        if (e) return e;
        FieldView v{s->view_buf + (t.p == s->Bm ? (long)n * n * 32 : 0), t.rows, rows * 32,
                    (long)(s->N + 1) * rows * 32};
        return download_pm(s, v, (long)(s->N + 1) * t.rows, out);
      }
      return download_pm(s, fview(s, t.p, t.rows), (long)(s->N + 1) * t.rows, out);
    }
  }
  // duals live in their own record stream [group][knot][z rows | z_est rows][32]
  if (strcmp(name, "z") == 0 || strcmp(name, "z_est") == 0) {
    if (rows_out) *rows_out = zrows;
    if (!out || zrows == 0) return ALTRO_B200_NO_ERROR;
    CUDA_OK(cudaSetDevice(s->device));
    FieldView v{name[1] == 0 ? s->z : s->zest, zrows, s->Rz, s->GSz};
    return download_pm(s, v, (long)(s->N + 1) * zrows, out);
  }
  struct { const char* nm; int view, rows; } derived[] = {
      {"constraint_val", KV_CONSTRAINT_VAL, zrows}, {"z_proj", KV_Z_PROJ, zrows},
      {"lxx", KV_LXX, n * n}, {"luu", KV_LUU, m * m}, {"lux", KV_LUX, m * n}, {"rho", KV_RHO, 1}};
  for (auto& t : derived) {
    if (strcmp(t.nm, name) == 0) {
      if (rows_out) *rows_out = t.rows;
      if (!out || t.rows == 0) return ALTRO_B200_NO_ERROR;
      CUDA_OK(cudaSetDevice(s->device));
      int e = ensure_view(s, t.rows);
      if (e) return e;
      CUDA_OK(cudaMemsetAsync(s->view_buf, 0, (size_t)s->G * (s->N + 1) * t.rows * 32 * 8, s->stream));
      s->ph.view = t.view;
      s->ph.view_rows = t.rows;
      e = run_host_op(s, OP_KNOT_VIEW, s->view_buf);
      if (e) return e;
      FieldView v{s->view_buf, t.rows, (long)t.rows * 32, (long)(s->N + 1) * t.rows * 32};
      return download_pm(s, v, (long)(s->N + 1) * t.rows, out);
    }
  }
  return ALTRO_B200_BAD_INDEX;
}

// SetDualGeneric / GetDualGeneral (altro_solver.hpp:359, :416; declared, never defined in the
// reference): the dual z of constraint slot `constraint` at knot k for every problem, [B][dim].
static int dual_view(altro_b200_solver* s, int constraint, int k, FieldView* v, int* dim) {
  if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;
  if (constraint < 0 || constraint >= s->con_h.ncon) return ALTRO_B200_BAD_INDEX;
  const ConSlot& c = s->con_h.slot[constraint];
  if (k < c.k_start || k >= c.k_stop) return ALTRO_B200_BAD_INDEX;
  *dim = c.dim;
  *v = FieldView{s->z + (long)k * s->Rz + (long)c.row0 * 32, c.dim, 0, s->GSz};
  return 0;
}
int altro_b200_get_dual_general(altro_b200_solver* s, int constraint, int k, double* z) {
  if (!s || !z) return ALTRO_B200_INVALID_POINTER;
  FieldView v;
  int dim = 0;
  int e = dual_view(s, constraint, k, &v, &dim);
  if (e) return e;
  CUDA_OK(cudaSetDevice(s->device));
  return download_pm(s, v, dim, z);
}
int altro_b200_set_dual_general(altro_b200_solver* s, int constraint, int k, const double* z,
                                int per_problem) {
  if (!s || !z) return ALTRO_B200_INVALID_POINTER;
  FieldView v;
  int dim = 0;
  int e = dual_view(s, constraint, k, &v, &dim);
  if (e) return e;
  CUDA_OK(cudaSetDevice(s->device));
  e = per_problem ? upload_pm(s, z, dim, v) : upload_shared(s, z, dim, v);
  if (e) return e;
  CUDA_OK(cudaStreamSynchronize(s->stream));
  return ALTRO_B200_NO_ERROR;
}
int altro_b200_get_num_constraints(const altro_b200_solver* s) { return s ? s->con_h.ncon : 0; }
int altro_b200_get_constraint_dim(const altro_b200_solver* s, int constraint) {
  if (!s || constraint < 0 || constraint >= s->con_h.ncon) return 0;
  return s->con_h.slot[constraint].dim;
}

#define GETTER_VEC(name, field, type)                                                       \
  int name(altro_b200_solver* s, type* out) {                                               \
    if (!s || !out) return ALTRO_B200_INVALID_POINTER;                                      \
    if (!s->initialized) return ALTRO_B200_SOLVER_NOT_INITIALIZED;                          \
    CUDA_OK(cudaSetDevice(s->device));                                                      \
    CUDA_OK(cudaMemcpyAsync(out, s->field, sizeof(type) * (size_t)s->B, cudaMemcpyDeviceToHost, \
                            s->stream));                                                    \
    CUDA_OK(cudaStreamSynchronize(s->stream));                                              \
    return ALTRO_B200_NO_ERROR;                                                             \
  }
GETTER_VEC(altro_b200_get_status, status, int)
GETTER_VEC(altro_b200_get_iterations, iters, int)
GETTER_VEC(altro_b200_get_merit_evals, merit_evals, int)
GETTER_VEC(altro_b200_get_final_objective, phi, double)
GETTER_VEC(altro_b200_get_stationarity, stat, double)
GETTER_VEC(altro_b200_get_primal_feasibility, feas, double)
GETTER_VEC(altro_b200_get_penalty, rho, double)

int altro_b200_get_horizon_length(const altro_b200_solver* s) { return s ? s->N : 0; }
int altro_b200_get_batch(const altro_b200_solver* s) { return s ? s->B : 0; }
int altro_b200_get_state_dim(const altro_b200_solver* s) { return s ? s->n : 0; }
int altro_b200_get_input_dim(const altro_b200_solver* s) { return s ? s->m : 0; }
long altro_b200_device_bytes(const altro_b200_solver* s) { return s ? s->bytes : 0; }

}  // extern "C"
