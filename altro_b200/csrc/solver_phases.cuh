// solver_phases.cuh -- the AL-iLQR iteration as a pipeline of phase kernels over compacted work
// lists (the production path; the single persistent kernel of solver_kernels.cuh stays as the
// differential-testing twin: both must agree bit for bit).
//
// Why: the persistent thread-per-trajectory kernel is latency bound (ncu, profiles/r01: one warp
// per scheduler, 12 % issue utilisation) and loses half its lanes to divergence -- trajectories of
// one warp need different numbers of line-search evaluations and iterations.  Here
//   * every sequential sweep (backward Riccati, rollout, d(phi) scan, convergence criteria) is its
//     own small kernel, launched over a COMPACTED list of the trajectories that actually need it,
//     so warps stay dense and the instruction footprint of each kernel fits the I-cache;
//   * everything that is independent per knot -- dynamics Jacobians (the transcendental-heavy
//     part), projected duals, cost gradients -- runs one thread per (trajectory, knot) and is
//     throughput- instead of latency-bound;
//   * the redundant alpha = 0 rollout of ForwardPass (solver.cpp:241, ~40 % of the reference's
//     merit evaluations) is replaced by a linear scan that provably reproduces it
//     (TrajSolver::phase_phi0_scan).
// Decisions (line search, dual/penalty update, convergence) are identical to the reference.
#pragma once
#include "solver_kernels.cuh"

namespace altro_b200 {

// ------------------------------------------------------------------ list compaction
// Ordered compaction of `in[0..count)` by (flags[in[i]] & mask) != 0 into out; writes the number
// kept to counters[slot] (and, if mask2 != 0, the number that also has mask2 to counters[slot2]).
// `count` bounds the input length; `dcount` (optional) is its exact device-side value.
// One CTA; ordered so that neighbouring lanes keep touching neighbouring problems.
static __global__ void __launch_bounds__(1024) k_compact(const int* __restrict__ in, int count,
                                                  const int* dcount,
                                                  const int* __restrict__ flags, int mask,
                                                  int* __restrict__ out, int* counters, int slot,
                                                  int mask2, int slot2) {
  __shared__ int warp_tot[32];
  __shared__ int warp_tot2[32];
  __shared__ int base_s, base2_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // the input length may live on the device (possibly in the very counter this kernel rewrites
  // at the end, hence read before the first barrier)
  if (dcount) count = min(count, *dcount);
  if (tid == 0) {
    base_s = 0;
    base2_s = 0;
  }
  __syncthreads();
  for (int start = 0; start < count; start += 1024) {
    const int i = start + tid;
    int b = -1, keep = 0, keep2 = 0;
    if (i < count) {
      b = in ? in[i] : i;
      const int f = flags[b];
      keep = (f & mask) != 0;
      keep2 = keep && mask2 && (f & mask2) != 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const unsigned bal2 = __ballot_sync(0xffffffffu, keep2);
    if (lane == 0) {
      warp_tot[wid] = __popc(bal);
      warp_tot2[wid] = __popc(bal2);
    }
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < wid; ++w) off += warp_tot[w];
    if (keep) out[off + __popc(bal & ((1u << lane) - 1u))] = b;
    __syncthreads();
    if (tid == 0) {
      int t = 0, t2 = 0;
      for (int w = 0; w < 32; ++w) {
        t += warp_tot[w];
        t2 += warp_tot2[w];
      }
      base_s += t;
      base2_s += t2;
    }
    __syncthreads();
  }
  if (tid == 0) {
    counters[slot] = base_s;
    if (mask2) counters[slot2] = base2_s;
  }
}

// ------------------------------------------------------------------ phase kernels
// Every list kernel takes an upper bound `count` (sizes the grid) and an optional device-side
// exact count, so a freshly compacted list can be consumed without a host round trip.
__device__ __forceinline__ int list_count(int count, const int* dcount) {
  return dcount ? min(count, *dcount) : count;
}

__device__ __forceinline__ LsOptions ls_options(const DevOptions& o) {
  LsOptions lo;
  lo.try_cubic_first = true;  // solver.cpp:248
  lo.use_backtracking = o.use_backtracking_linesearch != 0;
  lo.c1 = o.ls_c1;
  lo.c2 = o.ls_c2;
  return lo;
}

// this lane's column of the per-warp staging ring (dynamic shared memory; CTA = one warp)
extern __shared__ double altro_stage_ring[];
__device__ __forceinline__ double* lane_ring() { return altro_stage_ring + (threadIdx.x & 31); }

// K0: Solve() prologue, sequential part (solver.cpp:417-423)
template <class Model, bool CON>
__global__ void __launch_bounds__(32) k_phase_init(const DeviceProblem P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  TrajSolver<Model, CON> s(P, b);
  s.phase_init_rollout();
  P.flags[b] = TF_ACTIVE;
  P.iter_count[b] = 0;
  P.merit_evals[b] = 0;
  P.status[b] = SOLVE_UNSOLVED;
  P.ls_fail[b] = 0;
  P.sel[b] = -1;
}

// Expansion, one thread per (list entry, knot).  `mask`: only trajectories whose flags have one of
// these bits are processed (0 = all).  with_dyn: also recompute [A B].  slot_mode: -1 read the
// main trajectory, >= 0 that candidate slot, -2 the slot recorded in sel[b].  dual_first: apply
// the dual update z <- Pi(z_est) of this knot before recomputing the projected duals.
// The prologue calls this BEFORE the penalty reset, which reproduces quirk Q3 (gradient with the
// old rho, solver.cpp:424-430).
template <class Model, bool CON>
__global__ void __launch_bounds__(128) k_phase_expand(const DeviceProblem P, const int* list,
                                                      int count, const int* dcount, int mask,
                                                      bool with_dyn, int slot_mode,
                                                      bool dual_first) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (t >= list_count(count, dcount)) return;
  const int b = list ? list[t] : t;
  if (mask && !(P.flags[b] & mask)) return;
  TrajSolver<Model, CON> s(P, b);
  s.rho = CON ? P.rho[b] : 1.0;
  const int slot = (slot_mode == -2) ? P.sel[b] : slot_mode;
  s.phase_expand_knot(k, with_dyn, slot, dual_first);
}

// penalty reset at the end of the prologue (SetPenalty(penalty_initial), solver.cpp:429)
static __global__ void k_phase_set_rho(double* rho, int B, double value) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) rho[b] = value;
}

// K1: CalcExpansions + BackwardPass + the alpha = 0 half of ForwardPass (solver.cpp:448-450,
// :241-245) and the start of the line search.
template <class Model, bool CON>
__global__ void __launch_bounds__(32) k_phase_backward(const DeviceProblem P, const int* list,
                                                       int count, int depth) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int b = list[t];
  TrajSolver<Model, CON> s(P, b);
  s.rho = CON ? P.rho[b] : 1.0;
  constexpr bool ST = TrajSolver<Model, CON>::kStaged;
  s.template backward_sweep<ST>(lane_ring(), depth);
  double phi0, dphi0;
  s.template phase_phi0_scan<ST>(&phi0, &dphi0, lane_ring(), depth);
  P.phi0[b] = phi0;
  P.dphi0[b] = dphi0;
  P.phi[b] = phi0;
  P.merit_evals[b] += 1;
  P.stat_acc[b] = 0ull;
  P.feas_acc[b] = 0ull;
  P.sel[b] = -1;
  int f = P.flags[b] & TF_ACTIVE;
  if (fabs(dphi0) < P.opts.tol_meritfun_gradient) {
    // MeritFunctionGradientTooSmall: alpha = 0, not fatal (solver.cpp:242-245, :451)
    P.alpha_eval[b] = 0.0;
  } else {
    const LsOptions lo = ls_options(P.opts);
    LsMachine ls;
    if (ls.start(lo, 1.0, phi0, dphi0)) {
      f |= TF_NEED_EVAL | TF_WANT_DERIV;
      P.alpha_eval[b] = ls.alpha;
    } else {
      // NOT_DESCENT_DIRECTION: Run returns 0 without evaluating -> LineSearchFailed (:264-269)
      f |= TF_LS_FAILED;
      P.alpha_eval[b] = 0.0;
    }
    P.ls[b] = ls;
  }
  P.flags[b] = f;
}

// K2: rollouts.  Plain round: one candidate per trajectory (alpha_eval) into the main trajectory.
// Speculative round (backtracking line search): candidate `blockIdx.y` of every trajectory is
// rolled out concurrently into its own slot -- slot 0 is the step the state machine asked for,
// slots j >= 1 are the halvings it will ask for next if it keeps rejecting
// (SimpleBacktracking, linesearch.cpp:385-412), alpha_bt * 2^-(j-1).
template <class Model, bool CON>
__global__ void __launch_bounds__(32) k_phase_rollout(const DeviceProblem P, const int* list,
                                                      int count, const int* dcount,
                                                      bool speculative, int depth) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= list_count(count, dcount)) return;
  const int b = list[t];
  TrajSolver<Model, CON> s(P, b);
  s.rho = CON ? P.rho[b] : 1.0;
  constexpr bool ST = TrajSolver<Model, CON>::kStaged;
  if (!speculative) {
    P.phi_eval[b] = s.template phase_rollout<ST>(P.alpha_eval[b], -1, lane_ring(), depth);
  } else {
    const int slot = blockIdx.y;
    const double alpha = (slot == 0) ? P.alpha_eval[b] : ldexp(P.alpha_bt[b], -(slot - 1));
    P.phi_s[(long)slot * P.Bp + b] = s.template phase_rollout<ST>(alpha, slot, lane_ring(), depth);
  }
}

// K4: d(phi) scan (when requested) + the line-search state machine.  In a speculative round the
// candidates are fed to the machine in the order the reference would have evaluated them, and
// feeding stops at the first one it accepts, so the decisions (and the reported evaluation count)
// are those of the sequential search.
template <class Model, bool CON>
__global__ void __launch_bounds__(32) k_phase_lsupdate(const DeviceProblem P, const int* list,
                                                       int count, const int* dcount,
                                                       bool speculative, int depth) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= list_count(count, dcount)) return;
  const int b = list[t];
  int f = P.flags[b];
  const bool had_deriv = (f & TF_WANT_DERIV) != 0;
  double dphi = 0.0;
  if (had_deriv) {
    TrajSolver<Model, CON> s(P, b);
    dphi = s.template phase_dphi_scan<TrajSolver<Model, CON>::kStaged>(lane_ring(), depth);
  }
  const LsOptions lo = ls_options(P.opts);
  LsMachine ls = P.ls[b];
  bool last_had_deriv = had_deriv;
  int fed = 0;
  if (!speculative) {
    ls.update(lo, P.phi_eval[b], dphi);
    fed = 1;
  } else {
    int last_slot = -1;
    for (int slot = 0; slot < P.nslots && !ls.done(); ++slot) {
      const double cand = (slot == 0) ? P.alpha_eval[b] : ldexp(P.alpha_bt[b], -(slot - 1));
      if (ls.alpha != cand) break;  // not the step the machine is asking for (cannot happen)
      const bool want = ls.want_derivative();
      if (want && !(slot == 0 && had_deriv)) break;
      ls.update(lo, P.phi_s[(long)slot * P.Bp + b], want ? dphi : 0.0);
      last_had_deriv = want;
      last_slot = slot;
      fed += 1;
    }
    if (last_slot >= 0) P.sel[b] = last_slot;
    if (fed == 0) {  // defensive: never stall the pipeline
      ls.finish(LS_NOERROR, ls.alpha);
    }
  }
  P.merit_evals[b] += fed;
  f &= ~(TF_NEED_EVAL | TF_WANT_DERIV | TF_SPECULATE);
  if (!ls.done()) {
    f |= TF_NEED_EVAL;
    if (ls.want_derivative()) f |= TF_WANT_DERIV;
    P.alpha_eval[b] = ls.alpha;
    if (lo.use_backtracking) {
      f |= TF_SPECULATE;
      // first halving the machine will ask for after the pending step is rejected
      P.alpha_bt[b] = (ls.phase == LsMachine::P_CUBIC_FIRST ? ls.alpha0 : ls.alpha) * lo.beta_decrease;
    }
  } else {
    const double alpha = ls.alpha;
    P.alpha_eval[b] = alpha;
    if (ls.n_iters > 0) P.phi[b] = ls.phi;
    // the accepted candidate lives in a slot and/or has no derivative information yet
    // (solver.cpp:256-262): copy it into x_, u_ and expand it
    if (P.sel[b] >= 0 || (lo.use_backtracking && fabs(alpha - 1.0) > 0 && !last_had_deriv))
      f |= TF_REFRESH_DYN;
    if (isnan(alpha) || !(ls.status == LS_MINIMUM_FOUND || ls.status == LS_HIT_MAX_STEPSIZE))
      f |= TF_LS_FAILED;
  }
  P.ls[b] = ls;
  P.flags[b] = f;
}

// K5a/b: costates of the accepted point, then stationarity / feasibility residuals and
// CopyTrajectory -- one thread per (trajectory, knot)
template <class Model, bool CON>
__global__ void __launch_bounds__(128) k_phase_costate(const DeviceProblem P, const int* list,
                                                       int count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  TrajSolver<Model, CON> s(P, list[t]);
  s.phase_costate_knot(blockIdx.y);
}

template <class Model, bool CON>
__global__ void __launch_bounds__(128) k_phase_residual(const DeviceProblem P, const int* list,
                                                        int count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  TrajSolver<Model, CON> s(P, list[t]);
  s.phase_residual_knot(blockIdx.y);
}

// K5c: convergence test, dual / penalty update decision (solver.cpp:459-489, :503-506)
template <bool CON>
__global__ void __launch_bounds__(128) k_phase_decide(const DeviceProblem P, const int* list,
                                                      int count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int b = list[t];
  const DevOptions& o = P.opts;
  const double stationarity = __longlong_as_double((long long)P.stat_acc[b]);
  const double feasibility = __longlong_as_double((long long)P.feas_acc[b]);
  int f = P.flags[b];
  bool stop = (f & TF_LS_FAILED) != 0;
  int status = SOLVE_UNSOLVED;
  if (fabs(stationarity) < o.tol_stationarity && feasibility < o.tol_primal_feasibility) {
    stop = true;
    status = SOLVE_SUCCESS;
  }
  f &= ~(TF_REFRESH_DYN | TF_REFRESH_GRAD);
  if (stationarity < sqrt(o.tol_stationarity)) {
    if constexpr (CON) {
      // z <- Pi(z_est) is applied knot by knot by the expansion that follows (dual_first)
      if (feasibility > o.tol_primal_feasibility)
        P.rho[b] = fmin(P.rho[b] * o.penalty_scaling, o.penalty_max);
      f |= TF_REFRESH_GRAD;
    }
  }
  const int iter = P.iter_count[b] + 1;  // iterations completed
  P.iter_count[b] = iter;
  if (!stop && iter >= o.iterations_max) {
    stop = true;
    status = SOLVE_MAX_ITERATIONS;
  }
  P.stat[b] = stationarity;
  P.feas[b] = feasibility;
  if (stop) {
    f &= ~TF_ACTIVE;
    P.status[b] = status;
    // stats.iterations = iter + 1 with the loop counter at exit (quirk Q4): a loop that ran to
    // exhaustion reports iterations_max + 1
    P.iters[b] = (status == SOLVE_MAX_ITERATIONS) ? iter + 1 : iter;
    P.ls_fail[b] = (f & TF_LS_FAILED) ? 1 : 0;
  }
  P.flags[b] = f;
}

// ALTROSolver::OpenLoopRollout (solver.cpp:116-131): x_[k+1] = f(x_[k], u_[k]) from the initial state
template <class Model, bool CON>
__global__ void __launch_bounds__(32) k_open_loop_rollout(const DeviceProblem P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  TrajSolver<Model, CON> s(P, b);
  constexpr int n = Model::n, m = Model::m;
  double x[n], u[m], xn[n];
  load_block<n>(s.G(P.x0, n), 0, 0, x);
  for (int k = 0; k < P.N; ++k) {
    load_block<m>(s.F(P.u), s.S, k, u);
    s.dynamics(k, x, u, xn);
    store_block<n>(s.F(P.x), s.S, k, x);
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = xn[i];
  }
  store_block<n>(s.F(P.x), s.S, P.N, x);
}

// ALTROSolver::CalcCost (solver.cpp:163-174): sum_k cost(k) incl. the AL terms at the working
// trajectory; refreshes the projected duals like the reference does.
template <class Model, bool CON>
__global__ void __launch_bounds__(32) k_calc_cost(const DeviceProblem P, double* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  TrajSolver<Model, CON> s(P, b);
  s.rho = CON ? P.rho[b] : 1.0;
  constexpr int n = Model::n, m = Model::m;
  double cost = 0.0;
  for (int k = 0; k <= P.N; ++k) {
    const bool terminal = (k == P.N);
    double x[n], u[m], q[n], r[m];
    load_block<n>(s.F(P.x), s.S, k, x);
    load_block<n>(s.F(P.q), s.S, k, q);
    if (!terminal) {
      load_block<m>(s.F(P.u), s.S, k, u);
      load_block<m>(s.F(P.r), s.S, k, r);
    } else {
#pragma unroll
      for (int i = 0; i < m; ++i) u[i] = 0.0;
    }
    cost += s.stage_cost(k, x, u, q, r, terminal) + s.al_terms(k, x, u, terminal, false, nullptr, nullptr);
  }
  out[b] = cost;
}

}  // namespace altro_b200
