// solver_phases.cuh -- the AL-iLQR iteration as TWO kernels per iteration over the GROUPS of the
// batch, with no host in the loop (the production path; the single thread-per-trajectory kernel of
// solver_kernels.cuh stays as the differential-testing twin: both must agree bit for bit).
//
// A group is 32 consecutive problems = one stream of knot records (device_problem.h).  Groups never
// exchange data, so every kernel is launched over all groups of a sub-batch and a CTA whose
// problems have all stopped returns at once: no work lists, no compaction, no counter read-back.
//   k_phase_backward  one warp per group: CalcExpansions + the Riccati sweep + the alpha = 0 merit
//                     evaluation of ForwardPass as a linear scan (solver.cpp:241 re-simulates the
//                     accepted trajectory, ~40 % of the reference's merit evaluations; the scan
//                     provably reproduces it, TrajSolver::phi0_step), streaming the knot records
//                     through a shared-memory ring of TMA bulk copies (linalg.cuh, BulkRing).
//   k_phase_forward   one CTA per group: the whole line search (rounds of one pass each: rollout
//                     warp, follower warp for the derivative half, speculating warps; state
//                     machines), then one fused expansion / costate / residual / copy pass, the
//                     convergence / AL decision and the post-update expansion.
// The host enqueues iterations back to back on the sub-batch's stream and looks at a stop counter
// two iterations late (an asynchronous 4-byte copy), so the GPU never waits for the host.
// Decisions (line search, dual/penalty update, convergence) are identical to the reference.
#pragma once
#include "solver_kernels.cuh"

namespace altro_b200 {

// ------------------------------------------------------------------ phase kernels

__device__ __forceinline__ LsOptions ls_options(const DevOptions& o) {
  LsOptions lo;
  lo.try_cubic_first = true;  // solver.cpp:248
  lo.use_backtracking = o.use_backtracking_linesearch != 0;
  lo.c1 = o.ls_c1;
  lo.c2 = o.ls_c2;
  return lo;
}

// dynamic shared memory of the sweep kernels: mbarriers + staging ring (linalg.cuh)
extern __shared__ __align__(128) unsigned char altro_smem[];

// Copies the cost weights Qd [(N+1) n], Rd [N m] into shared memory behind the staging ring (when
// the launch reserved room for them: wcount > 0) and points the solver at the copy.  All threads
// of the CTA call it; the caller synchronises afterwards.
template <class TS>
__device__ __forceinline__ void stage_weights(TS& s, const DeviceProblem& P, double* wsm, int wcount) {
  if (wcount <= 0) return;
  const int nq = (P.N + 1) * TS::n, nr = P.N * TS::m;
  for (int i = threadIdx.x; i < nq; i += blockDim.x) wsm[i] = P.Qd[i];
  for (int i = threadIdx.x; i < nr; i += blockDim.x) wsm[nq + i] = P.Rd[i];
  s.Qd = wsm;
  s.Rd = wsm + nq;
}

// K0: Solve() prologue, sequential part (solver.cpp:417-423)
template <class Model, int CON>
__global__ void __launch_bounds__(32) k_phase_init(const __grid_constant__ DeviceProblem P) {
  const int b = (P.g0 + blockIdx.x) * 32 + threadIdx.x;
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

// Expansion of the Solve() prologue (solver.cpp:425-428), one thread per (problem, knot): [A B],
// projected duals, cost gradients at the initial rollout.  Launched BEFORE the penalty reset,
// which reproduces quirk Q3 (gradient with the old rho, solver.cpp:424-430).
template <class Model, int CON>
__global__ void __launch_bounds__(128) k_phase_expand(const __grid_constant__ DeviceProblem P, int count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  const int gi = t >> 5;
  if (gi >= count) return;
  const int b = (P.g0 + gi) * 32 + (t & 31);
  if (b >= P.B) return;
  TrajSolver<Model, CON> s(P, b);
  s.rho = CON ? P.rho[b] : 1.0;
  s.phase_expand_knot(k, true, -1, false);
}

// penalty reset at the end of the prologue (SetPenalty(penalty_initial), solver.cpp:429)
static __global__ void k_phase_set_rho(double* rho, int b0, int b1, double value) {
  const int b = b0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (b < b1) rho[b] = value;
}

// The alpha = 0 half of ForwardPass (solver.cpp:241) as a linear scan over the knots, by ONE warp
// (lane = problem) behind the Riccati sweep of either backward kernel; `ring` continues the sweep's
// stage / parity state.
template <class Model, int CON>
__device__ __forceinline__ void backward_scans(const DeviceProblem& P, TrajSolver<Model, CON>& s, BulkRing& ring,
                                               int depth, int g, int lane, int b, bool active, int first,
                                               double& phi0, double& dphi0) {
  using TS = TrajSolver<Model, CON>;
  constexpr int n = Model::n, m = Model::m;
  constexpr int kV = TS::kV;
  constexpr int kRowsPhi = TS::kRowsPhi;
  const int zr = CON ? 2 * P.zrows : 0;
  const double* rec = P.xbar + (long)g * P.GS;  // row 0 of the group's knot-0 record
  const double* zrec = CON ? P.z + (long)g * P.GSz : nullptr;
  {
    // phi0 / dphi0 scan, knots ascending
    // K, d were just written by the lanes of this warp through the generic proxy; the bulk copies
    // read them through the async proxy: every writer fences, then the leader issues
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    if (CON == 0 && !first) {
      // Unconstrained problems after the first iteration: merit(0) at the accepted point IS the
      // merit value the line search accepted it with (same stage costs summed in the same
      // order), and lx, lu in HBM are already those of the accepted point; only the directional
      // derivative depends on the new gains.  Scan [K d] [J] [lx lu] instead of
      // [q r c K d x u J] and skip the lx, lu write-back.
      constexpr int kRows1 = m * n + m;
      auto fetch_d = [&](int k, int st) {
        ring.expect(st, TS::kRowsDphi * 256);
        ring.copy(st, 0, rec + (long)k * P.R + TS::rK * 32, kRows1 * 256);
        ring.copy(st, kRows1, rec + (long)k * P.R + TS::rA * 32, kV * 256);
        ring.copy(st, kRows1 + kV, rec + (long)k * P.R + TS::rLx * 32, (n + m) * 256);
      };
      if (lane == 0)
        for (int j = 0; j < depth; ++j)
          if (j < P.N) fetch_d(j, (ring.s + j) % depth);
      double dxda[n];
#pragma unroll
      for (int i = 0; i < n; ++i) dxda[i] = 0.0;
      for (int k = 0; k < P.N; ++k) {
        const double* st = ring.wait();
        double K[m * n], d[m], A[n * n], Bm[n * m], lx[n], lu[m];
        if (active) {
          unstage_block<m * n>(st, 0, lane, K);
          unstage_block<m>(st, m * n, lane, d);
          s.unstage_jac(st, kRows1, lane, k, A, Bm);
          unstage_block<n>(st, kRows1 + kV, lane, lx);
          unstage_block<m>(st, kRows1 + kV + n, lane, lu);
          s.dphi_step(K, d, A, Bm, lx, lu, dxda, dphi0);
        }
        __syncwarp();
        if (lane == 0 && k + depth < P.N) fetch_d(k + depth, ring.s);
        ring.advance();
      }
      if (active) {
        dphi0 = s.dphi_terminal(dxda, dphi0);
        phi0 = P.phi[b];
      }
    } else {
    auto fetch_phi = [&](int k, int st) {
      ring.expect(st, (kRowsPhi + zr) * 256);
      ring.copy(st, 0, rec + (long)k * P.R + TS::rQ * 32, kRowsPhi * 256);
      if (zr) ring.copy(st, kRowsPhi, zrec + (long)k * P.Rz, zr * 256);
    };
    if (lane == 0) {
      for (int j = 0; j < depth; ++j)
        if (j < P.N) fetch_phi(j, (ring.s + j) % depth);
    }
    double dxda[n];
#pragma unroll
    for (int i = 0; i < n; ++i) dxda[i] = 0.0;
    constexpr int oR = n, oC = n + m, oK = oC + 1, oD = oK + m * n, oX = oD + m, oU = oX + n,
                  oJ = oU + m;
    for (int k = 0; k < P.N; ++k) {
      const double* st = ring.wait();
      double x[n], u[m], q[n], r[m], K[m * n], d[m], A[n * n], Bm[n * m], cval = 0.0;
      if (active) {
        unstage_block<n>(st, 0, lane, q);
        unstage_block<m>(st, oR, lane, r);
        cval = st[oC * 32 + lane];
        unstage_block<m * n>(st, oK, lane, K);
        unstage_block<m>(st, oD, lane, d);
        unstage_block<n>(st, oX, lane, x);
        unstage_block<m>(st, oU, lane, u);
        s.unstage_jac(st, oJ, lane, k, A, Bm);
      }
      if (zr) s.zstage = st + kRowsPhi * 32 + lane;
      if (active) s.phi0_step(k, x, u, q, r, cval, K, d, A, Bm, dxda, phi0, dphi0);
      s.zstage = nullptr;
      __syncwarp();
      if (lane == 0 && k + depth < P.N) fetch_phi(k + depth, ring.s);
      ring.advance();
    }
    if (active) s.phi0_terminal(dxda, phi0, dphi0);
    }
  }
}

// Start of the line search from (phi0, dphi0) (solver.cpp:242-252) and the per-iteration resets
__device__ __forceinline__ void backward_finish(const DeviceProblem& P, int b, bool active, double phi0,
                                                double dphi0) {
  if (!active) return;
  P.phi0[b] = phi0;
  P.dphi0[b] = dphi0;
  P.phi[b] = phi0;
  P.merit_evals[b] += 1;
  P.stat_acc[b] = 0ull;
  P.feas_acc[b] = 0ull;
  P.sel[b] = -1;
  int f = TF_ACTIVE;
  if (fabs(dphi0) < P.opts.tol_meritfun_gradient) {
    // MeritFunctionGradientTooSmall: alpha = 0, not fatal (solver.cpp:242-245, :451)
    P.alpha_eval[b] = 0.0;
    if (P.ls_hist) atomicAdd(P.ls_hist + 19, 1ull);
  } else {
    const LsOptions lo = ls_options(P.opts);
    LsMachine ls;
    if (ls.start(lo, 1.0, phi0, dphi0)) {
      f |= TF_NEED_EVAL | TF_WANT_DERIV;
      P.alpha_eval[b] = ls.alpha;
      P.spec_known[b] = 0;
      if (lo.use_backtracking && P.nslots > 1 && P.spec_round1) {
        // the halvings SimpleBacktracking(alpha0 * beta_decrease) will try if alpha0 and the
        // cubic-first probe are rejected (linesearch.cpp:130-132, :385-412)
        f |= TF_SPECULATE;
        P.spec_base[b] = 1;
        P.alpha_bt[b] = ls.alpha0 * lo.beta_decrease;
      }
    } else {
      // NOT_DESCENT_DIRECTION: Run returns 0 without evaluating -> LineSearchFailed (:264-269)
      f |= TF_LS_FAILED;
      P.alpha_eval[b] = 0.0;
    }
    P.ls[b] = ls;
  }
  P.flags[b] = f;
}

// K1: CalcExpansions + BackwardPass + the alpha = 0 half of ForwardPass (solver.cpp:448-450,
// :241-245) and the start of the line search.  One warp per group of the sub-batch.
template <class Model, int CON>
__global__ void __launch_bounds__(32) k_phase_backward(const __grid_constant__ DeviceProblem P, int depth,
                                                       int wcount, int first) {
  using TS = TrajSolver<Model, CON>;
  constexpr int n = Model::n, m = Model::m;
  const int g = P.g0 + blockIdx.x;
  const int lane = threadIdx.x;
  const int b = g * 32 + lane;
  const bool active = b < P.B && (P.flags[b] & TF_ACTIVE);
  if (!__any_sync(0xffffffffu, active)) return;  // every problem of the group has stopped
  TS s(P, active ? b : g * 32);
  s.rho = (CON && active) ? P.rho[b] : 1.0;
  double phi0 = 0.0, dphi0 = 0.0;
  if constexpr (TS::kStaged) {
    // stage contents: Riccati sweep [J | lx lu]; phi0 scan [q r c K d x u J]  (J = packed [A B])
    constexpr int kV = TS::kV;
    constexpr int kRowsBw = TS::kRowsBw, kRowsPhi = TS::kRowsPhi;
    // constrained problems also stage the knot's dual record [z | z_est] behind the main rows
    const int zr = CON ? 2 * P.zrows : 0;
    const int kStage = (TS::kRowsBackwardKernel + zr) * 32;
    BulkRing ring;
    ring.init(altro_smem, depth, kStage, lane == 0);
    stage_weights(s, P, reinterpret_cast<double*>(altro_smem + BulkRing::bytes(depth, kStage)), wcount);
    __syncwarp();
    const double* rec = P.xbar + (long)g * P.GS;  // row 0 of the group's knot-0 record
    const double* zrec = CON ? P.z + (long)g * P.GSz : nullptr;
    auto fetch_bw = [&](int k, int st) {
      ring.expect(st, (kRowsBw + zr) * 256);
      ring.copy(st, 0, rec + (long)k * P.R + TS::rA * 32, kV * 256);
      ring.copy(st, kV, rec + (long)k * P.R + TS::rLx * 32, (n + m) * 256);
      if (zr) ring.copy(st, kRowsBw, zrec + (long)k * P.Rz, zr * 256);
    };
    if (lane == 0)
      for (int j = 0; j < depth; ++j)
        if (P.N - 1 - j >= 0) fetch_bw(P.N - 1 - j, (ring.s + j) % depth);
    double Pn[n * n], pn[n];
    if (active) s.riccati_terminal(Pn, pn);
    bool alive = active;
    for (int k = P.N - 1; k >= 0; --k) {
      const double* st = ring.wait();
      double A[n * n], Bm[n * m], Qx[n], Qu[m];
      if (alive) {
        s.unstage_jac(st, 0, lane, k, A, Bm);
        unstage_block<n>(st, kV, lane, Qx);
        unstage_block<m>(st, kV + n, lane, Qu);
      }
      if (zr) s.zstage = st + kRowsBw * 32 + lane;
      if (alive) alive = s.riccati_step(k, A, Bm, Qx, Qu, Pn, pn);
      s.zstage = nullptr;
      // release the stage only after the step has consumed what was read from it (see BulkRing)
      __syncwarp();
      if (lane == 0 && k - depth >= 0) fetch_bw(k - depth, ring.s);
      ring.advance();
    }
    backward_scans<Model, CON>(P, s, ring, depth, g, lane, b, active, first, phi0, dphi0);
  } else {
    if (active) {
      s.backward_sweep();
      s.phase_phi0_scan(&phi0, &dphi0);
    }
  }
  backward_finish(P, b, active, phi0, dphi0);
}

// Sub-phases of k_phase_forward, timed per CTA with %globaltimer when DeviceProblem::prof is set
// (profile mode): prof[s] accumulates the nanoseconds thread 0 of every CTA spent in sub-phase s,
// prof[FS_COUNT] the CTAs that did work.
enum FwdSub { FS_ROLLOUT = 0, FS_EXPAND = 1, FS_DPHI_LS = 2, FS_CRITERIA = 3, FS_COUNT = 4 };

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Work items of the knot-parallel sub-phases of k_phase_forward: one item per (flagged lane of
// the group, knot).  Consecutive threads take different lanes of the same knot, so a warp's loads
// and stores stay inside one or two 256-byte rows of the knot record.
template <class Fn>
__device__ __forceinline__ void for_knot_items(unsigned lanemask, int g, int knots, Fn&& fn) {
  const int nl = __popc(lanemask);
  if (nl == 0) return;
  const int items = nl * knots;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int k = i / nl;
    const int l = __fns(lanemask, 0, i % nl + 1);  // the (i % nl)-th flagged lane
    fn(g * 32 + l, k);
  }
}

// K2: everything of one iLQR iteration after the backward pass -- ForwardPass (solver.cpp:237-271)
// with the whole line search (linesearch.cpp:37-412), Stationarity / Feasibility / CopyTrajectory
// (:207-231, :148-157), the convergence test and the dual / penalty update (:459-489) -- for ONE
// group of 32 problems per CTA, with no host in the loop.  The CTA iterates line-search ROUNDS
// until every lane's state machine is done:
//   rollout pass    the ROLLOUT warp (warp 0, lane = problem) rolls out the step each lane's machine
//                   asked for (alpha_eval) into the working trajectory x_, u_; the speculating
//                   warps roll out, for the lanes whose search is (or is about to be)
//                   backtracking, the halvings SimpleBacktracking would try next
//                   (linesearch.cpp:385-412) -- the (lane, halving) pairs are packed densely over
//                   their threads; halvings <= nstore keep their trajectory in a candidate slot,
//                   deeper ones return the merit value only (an accepted one is rolled out again,
//                   TF_REROLL).  The warps with work consume ONE staged copy of the knot rows
//                   [xbar ubar q r c K d] through RollPipe (linalg.cuh): no CTA barrier per knot,
//                   nobody waits on a slower warp, the last warp out of a stage refills it.
//   derivatives     FOLLOW variant (default): the FOLLOWER warp (warp 1) does [A B], projected duals,
//                   lx, lu and the phi' recurrence of the requested step (solver.cpp:303-315) one
//                   knot behind the rollout warp, from x_k, u_k handed over in the stage.
//                   Otherwise: knot-parallel expansion by warps >= 1 and a staged phi' scan by warp 0
//                   behind them.
//   update          warp 0: every lane's LsMachine is fed the value of its request and, while it
//                   keeps backtracking, the precomputed halvings in the order the reference would
//                   have evaluated them; feeding stops at the first accept, so decisions and
//                   evaluation counts are those of the sequential search.
// then, still in the same launch: ONE pass that does the post-search expansion of an accepted
// backtracking step (:256-262), the costates, the residuals and CopyTrajectory
// (TrajSolver::post_chunk), the decision, and the expansion after a dual update (:483-486).
// active_out: incremented by the number of problems of the group that stopped in this iteration.
// CTA shape of the staged variants (register budget = 65536 / (threads * CTAs per SM))
#ifndef ALTRO_FWD_THREADS
#define ALTRO_FWD_THREADS 192
#endif
#ifndef ALTRO_FWD_CTAS
#define ALTRO_FWD_CTAS 2
#endif
constexpr int kFwdThreads = ALTRO_FWD_THREADS;
constexpr int kFwdCtasPerSm = ALTRO_FWD_CTAS;

// FOLLOW: the derivative half of a merit evaluation (solver.cpp:303-315: [A B], gradients, the phi'
// recurrence) is done by a FOLLOWER warp (warp 1) right behind the rollout warp, which hands it
// x_k, u_k through extra rows of the knot's stage (TrajSolver::follow_step) -- no separate
// expansion / d(phi) scan and no re-read of x, u, [J], lx, lu from HBM, while the state recursion
// of warp 0, the critical path of a pass, stays as short as a plain rollout.  Speculative
// candidates use warps >= 2.  It is a template parameter so that the variant with the separate
// knot-parallel expansion + scan (warps >= 1 speculate) stays available for comparison.
template <class Model, int CON, bool FOLLOW>
__global__ void __launch_bounds__((Model::n > kUnrollDim) ? 256 : kFwdThreads, (Model::n > kUnrollDim) ? 1 : kFwdCtasPerSm)
    k_phase_forward(const __grid_constant__ DeviceProblem P, int depth, int stage_rows, int wcount, int* done_out) {
  using TS = TrajSolver<Model, CON>;
  constexpr int n = Model::n, m = Model::m;
  constexpr unsigned kAll = 0xffffffffu;
  const int g = P.g0 + blockIdx.x;
  const int tid = threadIdx.x, lid = tid & 31, wid = tid >> 5;
  // speculative candidates per round: one warp each, at most what SetSpeculation asked for (extra
  // warps only serve the knot-parallel sub-phases)
  constexpr bool follow = TrajSolver<Model, CON>::kStaged && FOLLOW;
  constexpr int kSpecWarp0 = follow ? 2 : 1;  // first warp that rolls out speculative candidates
  const int nspec = max(0, min((int)(blockDim.x >> 5) - kSpecWarp0, P.nslots - 1));
  const int bl = g * 32 + lid;                   // the problem this thread's lane index names
  const bool valid = bl < P.B;
  int fl = valid ? P.flags[bl] : 0;
  // every warp reads the same 32 flags before anyone writes them: CTA-uniform
  const unsigned act_mask = __ballot_sync(kAll, (fl & TF_ACTIVE) != 0);
  if (!act_mask) return;

  // two pipes over the same stages: `pipe` for the rollout passes (the warps with work consume),
  // `scan` for the d(phi) scans (warp 0 alone); barriers armed once, never invalidated.  Follower
  // mode has no scan: scan.full[] serve as the "x_k, u_k of the rollout warp are in the stage"
  // barriers, with the follower's own parity bits.
  RollPipe pipe;
  BulkPipe scan;
  pipe.setup(altro_smem, reinterpret_cast<double*>(altro_smem + 256), depth, stage_rows * 32, P.N);
  scan.setup(altro_smem + BulkPipe::kBarBytes, reinterpret_cast<double*>(altro_smem + 256), depth, stage_rows * 32);
  if (TS::kStaged && tid == 0) {
    pipe.init();
    scan.init(1);
  }
  unsigned xubits = 0u;
  double* wsm = reinterpret_cast<double*>(altro_smem + BulkPipe::bytes(depth, stage_rows * 32));
  const int nq = (P.N + 1) * n, nr = P.N * m;
  if (wcount > 0) {
    for (int i = tid; i < nq; i += blockDim.x) wsm[i] = P.Qd[i];
    for (int i = tid; i < nr; i += blockDim.x) wsm[nq + i] = P.Rd[i];
  }
  // cost weights of the solver objects below come from shared memory when they were staged
  auto weights = [&](TS& s) {
    if (wcount > 0) {
      s.Qd = wsm;
      s.Rd = wsm + nq;
    }
  };
  unsigned long long t_prev = 0;
  const bool prof = P.prof != nullptr && tid == P.prof_tid;
  if (prof) t_prev = global_ns();
  auto tick = [&](int sub) {
    if (prof) {
      const unsigned long long t = global_ns();
      atomicAdd(P.prof + sub, t - t_prev);
      t_prev = t;
    }
  };
  const int zr = CON ? 2 * P.zrows : 0;
  const double* rec = P.xbar + (long)g * P.GS;  // row 0 of the group's knot-0 record
  const double* zrec = CON ? P.z + (long)g * P.GSz : nullptr;
  const LsOptions lo = ls_options(P.opts);
  // ready[k]: flagged lanes whose expansion of knot k is published (this round)
  int* ready = reinterpret_cast<int*>(wsm + wcount);
  // uniform linear cost terms: [q r c] of knot 0 for the group's 32 problems, kept in shared memory
  // for the whole kernel and laid out like the rows rQ.. of a stage
  constexpr int kQrcRows = TS::rK - TS::rQ;
  double* qrc0 = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ready) + ((P.N + 1) * 4 + 15) / 16 * 16);
  const bool quni = TS::kStaged && P.qrc_uniform != 0;
  if (quni)
    for (int i = tid; i < kQrcRows * 32; i += blockDim.x) qrc0[i] = rec[TS::rQ * 32 + i];
  // follower mode: x_N of the rollout warp [n][32] and the follower's phi' [32]
  double* xN_sm = qrc0 + kQrcRows * 32;
  double* dphi_sm = xN_sm + n * 32;
  __syncthreads();

  // ================================================================= line-search rounds
  for (int round = 0;; ++round) {
    if (round > 4 * kMaxHalvings + 64) {  // a search takes <= 25 evaluations (+ re-rollouts)
      if (tid == 0) printf("altro_b200: line search of group %d does not terminate\n", g);
      __trap();
    }
    fl = valid ? P.flags[bl] : 0;
    const unsigned pend = __ballot_sync(kAll, (fl & (TF_NEED_EVAL | TF_REROLL)) != 0);
    if (!pend) break;
    if (TS::kStaged)  // cleared for this round; the barrier behind the rollout pass orders it
      for (int i = tid; i <= P.N; i += blockDim.x) ready[i] = 0;
    const unsigned needy = __ballot_sync(kAll, (fl & TF_NEED_EVAL) && (fl & TF_SPECULATE));
    const unsigned dmask = __ballot_sync(kAll, (fl & TF_WANT_DERIV) != 0);
    const int nneedy = nspec > 0 ? __popc(needy) : 0;

    // ---- rollout pass.  Warp 0: thread = problem lane, candidate 0 (the requested step).  Follower
    // (warp 1 in follower mode): lane = problem, the derivative half of the requests that want it.
    // Speculating warps: pair p -> lane rank p % nneedy (consecutive threads = different lanes:
    // conflict-free shared-memory columns, neighbouring global stores), halving p / nneedy + 1.
    // Warps without work sit the pass out.
    const bool is_follow = follow && wid == 1;
    const unsigned fmask =
        follow ? __ballot_sync(kAll, (fl & TF_NEED_EVAL) && (fl & TF_WANT_DERIV) && !(fl & TF_REROLL)) : 0u;
    const int spec_warps = (nneedy * nspec + 31) >> 5;
    const int nact = 1 + (fmask ? 1 : 0) + spec_warps;  // consumer warps of this pass
    int lane = lid, slot = 0;
    bool need, warp_in;
    if (wid == 0) {
      need = (fl & (TF_NEED_EVAL | TF_REROLL)) != 0;
      warp_in = true;
    } else if (is_follow) {
      need = (fmask >> lid) & 1u;
      warp_in = fmask != 0u;
    } else {
      const int p = (wid - kSpecWarp0) * 32 + lid;
      need = p < nneedy * nspec;
      warp_in = wid - kSpecWarp0 < spec_warps;
      if (need) {
        lane = __fns(needy, 0, p % nneedy + 1);
        slot = p / nneedy + 1;
      }
    }
    const int b = g * 32 + lane;
    TS s(P, need ? b : g * 32);
    weights(s);
    s.rho = (CON && need) ? P.rho[b] : 1.0;
    double fdxda[follow ? n : 1], fdphi = 0.0;  // the follower's phi' recurrence
    if (warp_in) {
      double alpha = 0.0;
      double *xo = nullptr, *uo = nullptr;
      long so = 0;
      if (need && !is_follow) {
        if (slot == 0) {
          alpha = P.alpha_eval[b];
          xo = s.xw(-1);
          uo = s.uw(-1);
          so = s.sw(-1);
        } else {
          alpha = ldexp(P.alpha_bt[b], -(slot - 1));  // halving spec_base + slot - 1
          if (slot <= P.nstore) {
            xo = s.xw(slot - 1);
            uo = s.uw(slot - 1);
            so = s.sw(slot - 1);
          }
        }
      }
      double phi = 0.0;
      if constexpr (TS::kStaged) {
        constexpr int kRows = TS::kRowsRoll;  // [xbar ubar q r c K d] (+ the dual record)
        const int xu_row = kRows + zr;        // follower mode: x_k, u_k of the rollout warp
        auto fetch = [&](int k, int st) {
          if (quni) {  // [xbar ubar] and [K d] only
            pipe.arm(st, (unsigned)(kRows - kQrcRows + zr) * 256u);
            pipe.copy(st, 0, rec + (long)k * P.R, TS::rQ * 256);
            pipe.copy(st, TS::rK, rec + (long)k * P.R + TS::rK * 32, (kRows - TS::rK) * 256);
          } else {
            pipe.arm(st, (unsigned)(kRows + zr) * 256u);
            pipe.copy(st, 0, rec + (long)k * P.R, kRows * 256);
          }
          if (zr) pipe.copy(st, kRows, zrec + (long)k * P.Rz, zr * 256);
        };
        if (wid == 0 && lid == 0)
          for (int j = 0; j < depth && j < P.N; ++j) fetch(j, j);
        double x[n];
        if (need && !is_follow) load_block<n>(s.G(P.x0, n), 0, 0, x);
        if constexpr (follow) {
#pragma unroll
          for (int i = 0; i < n; ++i) fdxda[i] = 0.0;
        }
        long long c_full = 0, c_wr = 0, c_t0 = 0, c_pass = 0;  // profile mode: clocks of thread 0
        if (prof) c_pass = clock64();
        int st = 0, st_prev = 0;     // stage of knot k = k mod depth, kept incrementally
        unsigned out_prev = ~0u;     // lane 0: what the count-out of knot k - 1 returned
        for (int k = 0; k < P.N; ++k) {
          if (prof) c_t0 = clock64();
          double* stg = pipe.wait(st);
          if (prof) c_full += clock64() - c_t0;
          const double* qst = quni ? qrc0 - TS::rQ * 32 : stg;
          double u[m], q[n], r[m], K[m * n], d[m], cval = 0.0;
          if (need) {
            unstage_block<n>(qst, TS::rQ, lane, q);
            unstage_block<m>(qst, TS::rR, lane, r);
            unstage_block<m * n>(stg, TS::rK, lane, K);
            unstage_block<m>(stg, TS::rD, lane, d);
          }
          if (zr) s.zstage = stg + kRows * 32 + lane;
          if (follow && is_follow) {
            // x_k, u_k arrive from the rollout warp through the stage
            RollPipe::mbar_wait((unsigned)__cvta_generic_to_shared(scan.full + st), (xubits >> st) & 1u);
            xubits ^= 1u << st;
            if (need) {
              unstage_block<n>(stg, xu_row, lane, x);
              unstage_block<m>(stg, xu_row + n, lane, u);
              s.follow_step(k, x, u, q, r, K, d, fdxda, fdphi);
            }
          } else {
            if (need) {
              double xb[n], ub[m];
              unstage_block<n>(stg, TS::rXbar, lane, xb);
              unstage_block<m>(stg, TS::rUbar, lane, ub);
              cval = qst[TS::rC * 32 + lane];
              s.rollout_control(alpha, xb, ub, K, d, x, u);
            }
            if (follow && wid == 0 && fmask) {
              if (need) {
                double* xu = stg + xu_row * 32 + lane;
#pragma unroll
                for (int i = 0; i < n; ++i) xu[i * 32] = x[i];
#pragma unroll
                for (int i = 0; i < m; ++i) xu[(n + i) * 32] = u[i];
              }
              __syncwarp();
              if (lid == 0) {
                const unsigned a = (unsigned)__cvta_generic_to_shared(scan.full + st);
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
              }
            }
            if (need) s.rollout_advance(k, u, q, r, cval, x, xo, uo, so, phi);
          }
          s.zstage = nullptr;
          // count out of the stage only after the step consumed what was read from it (see
          // BulkRing).  Lane 0 looks at the count it got for the PREVIOUS knot -- that atomic's
          // latency is long over -- and refills that stage if this warp was the last one out.
          if (prof) c_t0 = clock64();
          __syncwarp();
          if (lid == 0) {
            if (out_prev == (unsigned)nact - 1u) {
              pipe.reset(st_prev);
              if (k - 1 + depth < P.N) fetch(k - 1 + depth, st_prev);
            }
            out_prev = pipe.count_out(st);
          }
          st_prev = st;
          st = (st + 1 == depth) ? 0 : st + 1;
          if (prof) c_wr += clock64() - c_t0;
        }
        if (lid == 0 && out_prev == (unsigned)nact - 1u) pipe.reset(st_prev);
        if (prof) {  // rollout warp: cycles waiting for a stage to land / counting out, per pass
          const unsigned long long c_all = (unsigned long long)(clock64() - c_pass);
          atomicAdd(P.prof + 5, (unsigned long long)c_full);
          atomicAdd(P.prof + 6, (unsigned long long)c_wr);
          atomicAdd(P.prof + 7, c_all);
          atomicAdd(P.prof + 8 + min(round, 3), c_all);
          atomicAdd(P.prof + 12 + min(round, 3), 1ull);
        }
        if (need && !is_follow) s.rollout_terminal(x, xo, so, phi);
        if constexpr (follow) {
          // hand x_N to the follower, which finishes phi' behind the CTA barrier below
          if (wid == 0 && need && fmask) {
#pragma unroll
            for (int i = 0; i < n; ++i) xN_sm[i * 32 + lane] = x[i];
          }
        }
      } else {
        if (need) phi = s.phase_rollout(alpha, xo, uo, so);
      }
      if (need && !is_follow) {
        if (slot == 0)
          P.phi_eval[b] = phi;
        else
          P.phi_s[(long)min(P.spec_base[b] + slot - 1, kMaxHalvings) * P.Bp + b] = phi;
      }
    } else if (TS::kStaged) {
      pipe.skip_pass();
    }
    if constexpr (follow) {
      if (fmask) {  // CTA-uniform
        __syncthreads();
        if (is_follow) {
          if (need) {
            double x[n];
            unstage_block<n>(xN_sm, 0, lane, x);
            s.follow_terminal(x, fdxda, fdphi);
          }
          dphi_sm[lid] = fdphi;
        }
      }
    }
    __syncthreads();
    tick(FS_ROLLOUT);

    // ---- expansion of the trial point of the lanes that asked for the derivative.  Staged models:
    // warps >= 1 expand in knot-major order and publish every finished knot (ready[k] counts the
    // flagged lanes done), so warp 0 can run the sequential d(phi) scan BEHIND them instead of
    // after them -- the scan's bulk copy of knot k is issued once ready[k] is complete.
    // follower mode: nothing left to expand or scan, the follower warp left phi' in shared memory
    const unsigned dmask_sep = follow ? 0u : dmask;
    const int nl = __popc(dmask_sep);
    if (dmask_sep) {
      if constexpr (TS::kStaged) {
        if (wid > 0) {
          const int items = nl * (P.N + 1), step = (int)blockDim.x - 32;
          for (int i = (wid - 1) * 32 + lid; i < items; i += step) {
            const int k = i / nl;
            const int b = g * 32 + __fns(dmask_sep, 0, i % nl + 1);
            TS s(P, b);
            weights(s);
            s.rho = CON ? P.rho[b] : 1.0;
            if (i + step < items) {  // the next item's rows on their way into L2 meanwhile
              TS sn(P, g * 32 + __fns(dmask_sep, 0, (i + step) % nl + 1));
              sn.phase_expand_prefetch((i + step) / nl, -1);
            }
            s.phase_expand_knot(k, true, -1, false);
            // [J] [lx lu] went through the generic proxy; the scan reads them with bulk copies
            // (async proxy): fence, then publish
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            atomicAdd(ready + k, 1);
          }
        }
      } else {
        for_knot_items(dmask_sep, g, P.N + 1, [&](int b, int k) {
          TS s(P, b);
          weights(s);
          s.rho = CON ? P.rho[b] : 1.0;
          s.phase_expand_knot(k, true, -1, false);
        });
        __syncthreads();
      }
    }
    tick(FS_EXPAND);
    // all lanes of warp 0: wait until every flagged lane's expansion of knot k has been published
    auto wait_ready = [&](int k) {
      volatile int* r = ready;
      unsigned long long spins = 0;
      while (r[k] < nl) {
        if (++spins > (1ull << 26)) {
          if (lid == 0) printf("altro_b200: expansion of knot %d never published (group %d)\n", k, g);
          __trap();
        }
      }
      __threadfence_block();
    };

    // ---- d(phi) scan + line-search machines: warp 0, lane = problem
    if (wid == 0) {
      const int b = bl;
      int f = fl;
      const bool pending = (f & (TF_NEED_EVAL | TF_REROLL)) != 0;
      const bool had_deriv = (f & TF_NEED_EVAL) && (f & TF_WANT_DERIV) && !(f & TF_REROLL);
      double dphi = (follow && had_deriv) ? dphi_sm[lid] : 0.0;
      if (dmask_sep) {
        TS s(P, had_deriv ? b : g * 32);
        if constexpr (TS::kStaged) {
          // stage contents: [K d] [J] [lx lu]
          constexpr int kV = TS::kV;
          constexpr int kRows1 = m * n + m;
          auto fetch = [&](int k) {
            const int st = scan.acquire(k, (unsigned)TS::kRowsDphi * 256u);
            scan.copy(st, 0, rec + (long)k * P.R + TS::rK * 32, kRows1 * 256);
            scan.copy(st, kRows1, rec + (long)k * P.R + TS::rA * 32, kV * 256);
            scan.copy(st, kRows1 + kV, rec + (long)k * P.R + TS::rLx * 32, (n + m) * 256);
          };
          for (int j = 0; j < depth && j < P.N; ++j) {
            wait_ready(j);
            if (lid == 0) fetch(j);
          }
          double dxda[n];
#pragma unroll
          for (int i = 0; i < n; ++i) dxda[i] = 0.0;
          for (int k = 0; k < P.N; ++k) {
            const double* st = scan.wait(k);
            double K[m * n], d[m], A[n * n], Bm[n * m], lx[n], lu[m];
            if (had_deriv) {
              unstage_block<m * n>(st, 0, lid, K);
              unstage_block<m>(st, m * n, lid, d);
              s.unstage_jac(st, kRows1, lid, k, A, Bm);
              unstage_block<n>(st, kRows1 + kV, lid, lx);
              unstage_block<m>(st, kRows1 + kV + n, lid, lu);
              s.dphi_step(K, d, A, Bm, lx, lu, dxda, dphi);
            }
            scan.release(k, lid);
            if (k >= 1 && k - 1 + depth < P.N) {
              scan.wait_writable(k - 1 + depth);
              wait_ready(k - 1 + depth);
              if (lid == 0) fetch(k - 1 + depth);
            }
          }
          scan.end_pass(P.N);
          wait_ready(P.N);  // lx of the terminal knot (read below with plain loads)
          if (had_deriv) dphi = s.dphi_terminal(dxda, dphi);
        } else {
          if (had_deriv) dphi = s.phase_dphi_scan();
        }
      }
      if (pending) {
        if (f & TF_REROLL) {
          // the accepted candidate has just been rolled out again into x_, u_; it still needs its
          // expansion (TF_REFRESH_DYN stays set)
          f &= ~(TF_REROLL | TF_NEED_EVAL | TF_WANT_DERIV | TF_SPECULATE);
          P.sel[b] = -1;
          P.flags[b] = f;
        } else {
          LsMachine ls = P.ls[b];
          bool last_had_deriv = had_deriv;
          int known = P.spec_known[b];  // halvings 1..known have their merit value in phi_s
          const int base = P.spec_base[b];
          const bool was_backtrack = ls.phase == LsMachine::P_BACKTRACK;
          int fed = 1, winner = 0;  // winner: halving index the search returned (0: the request)
          ls.update(lo, P.phi_eval[b], dphi);
          if (f & TF_SPECULATE) {
            // this round produced halvings base .. base+nspec-1; the request itself was halving
            // base-1 when the machine was already backtracking
            if (was_backtrack && base >= 2)
              P.phi_s[(long)min(base - 1, kMaxHalvings) * P.Bp + b] = P.phi_eval[b];
            known = min(base + nspec - 1, kMaxHalvings);
          }
          // feed the precomputed halvings in the order the sequential search would evaluate them
          while (!ls.done() && ls.phase == LsMachine::P_BACKTRACK) {
            // halving index of the step the machine asks for: alpha = alpha0 * 2^-j
            int j = 1;
            double cand = ls.alpha0 * lo.beta_decrease;
            while (j <= known && cand != ls.alpha) {
              cand *= lo.beta_decrease;
              ++j;
            }
            if (j > known) break;  // not precomputed: needs another round
            ls.update(lo, P.phi_s[(long)j * P.Bp + b], 0.0);
            last_had_deriv = false;
            winner = j;
            fed += 1;
          }
          if (!ls.done()) winner = 0;
          P.merit_evals[b] += fed;
          P.spec_known[b] = known;
          f &= ~(TF_NEED_EVAL | TF_WANT_DERIV | TF_SPECULATE);
          if (!ls.done()) {
            f |= TF_NEED_EVAL;
            if (ls.want_derivative()) f |= TF_WANT_DERIV;
            P.alpha_eval[b] = ls.alpha;
            if (lo.use_backtracking && nspec > 0) {
              if (ls.phase == LsMachine::P_BACKTRACK) {
                // ran out of precomputed halvings: the request is halving known+1, speculate the
                // next ones
                f |= TF_SPECULATE;
                P.spec_base[b] = known + 2;
                P.alpha_bt[b] = ls.alpha * lo.beta_decrease;
              } else if (ls.phase == LsMachine::P_CUBIC_FIRST) {
                // the cubic-first probe is next (linesearch.cpp:96-127).  If it is rejected the
                // search backtracks through the halvings; when none of the known ones passes the
                // Armijo test the ones after them ride along with the probe
                bool any = false;
                double cand = ls.alpha0;
                for (int j = 1; j <= known; ++j) {
                  cand *= lo.beta_decrease;
                  any = any || (P.phi_s[(long)j * P.Bp + b] <= ls.phi0 + lo.c1 * cand * ls.dphi0);
                }
                if (!any && known + 1 <= kMaxHalvings) {
                  f |= TF_SPECULATE;
                  P.spec_base[b] = known + 1;
                  P.alpha_bt[b] = cand * lo.beta_decrease;
                }
              }
            }
          } else {
            const double alpha = ls.alpha;
            P.alpha_eval[b] = alpha;
            if (ls.n_iters > 0) P.phi[b] = ls.phi;
            if (P.ls_hist) {
              int bin = 17;
              if (!(ls.status == LS_MINIMUM_FOUND || ls.status == LS_HIT_MAX_STEPSIZE)) bin = 18;
              else if (winner > 0) bin = winner < 15 ? winner : 15;
              else if (alpha == ls.alpha0) bin = 0;
              else if (lo.use_backtracking && last_had_deriv) bin = 16;
              else if (lo.use_backtracking) {  // a halving evaluated as the request of a later round
                int j = 1;
                double a = ls.alpha0 * lo.beta_decrease;
                while (j < 15 && a != alpha) { a *= lo.beta_decrease; ++j; }
                bin = j;
              }
              atomicAdd(P.ls_hist + bin, 1ull);
            }
            if (winner > 0) {
              // the step the search returns was only evaluated as a speculative candidate of the
              // round that started at halving `wbase`
              const int wbase = (winner >= base) ? base : 1;
              const int wslot = winner - wbase + 1;
              if (winner >= base && wslot >= 1 && wslot <= P.nstore)
                P.sel[b] = wslot - 1;  // its trajectory is in a candidate slot: copy + expand it
              else
                f |= TF_REROLL;        // merit-only / overwritten candidate: roll it out again
              f |= TF_REFRESH_DYN;
            } else if (lo.use_backtracking && fabs(alpha - 1.0) > 0 && !last_had_deriv) {
              f |= TF_REFRESH_DYN;  // accepted point has no derivative information yet (:256-262)
            }
            if (isnan(alpha) || !(ls.status == LS_MINIMUM_FOUND || ls.status == LS_HIT_MAX_STEPSIZE))
              f |= TF_LS_FAILED;
          }
          P.ls[b] = ls;
          P.flags[b] = f;
        }
      }
    }
    __syncthreads();
    tick(FS_DPHI_LS);
  }

  // ================================================================= after the search
  // accepted backtracking step: A, B, lx, lu at the accepted point, read from the candidate slot
  // that holds it (solver.cpp:256-262)
  // -- fused with the costates of the accepted point, the stationarity / feasibility residuals and
  // CopyTrajectory into one pass: warp w walks the w-th chunk of the knots downwards, lane =
  // problem (TrajSolver::post_chunk)
  {
    const int W = (int)(blockDim.x >> 5);
    const int C = (P.N + 1 + W - 1) / W;
    const int k0 = wid * C, k1 = min(k0 + C, P.N + 1);
    const bool act = valid && (fl & TF_ACTIVE) && k0 < k1;
    TS s(P, act ? bl : g * 32);
    weights(s);
    s.rho = (CON && act) ? P.rho[bl] : 1.0;
    const bool refresh = (fl & TF_REFRESH_DYN) != 0;
    const int slot = (act && refresh) ? P.sel[bl] : -1;
    double yn[n];
    if (act && k1 <= P.N) s.post_boundary(k1, slot, yn);
    __syncthreads();
    if (act) s.post_chunk(k0, k1, refresh, slot, yn);
    __syncthreads();
  }
  // convergence test, dual / penalty update decision (solver.cpp:459-489, :503-506)
  if (wid == 0) {
    const int b = bl;
    bool stopped = false;
    if (valid && (fl & TF_ACTIVE)) {
      const DevOptions& o = P.opts;
      const double stationarity = __longlong_as_double((long long)P.stat_acc[b]);
      const double feasibility = __longlong_as_double((long long)P.feas_acc[b]);
      int f = fl;
      bool stop = (f & TF_LS_FAILED) != 0;
      int status = SOLVE_UNSOLVED;
      if (fabs(stationarity) < o.tol_stationarity && feasibility < o.tol_primal_feasibility) {
        stop = true;
        status = SOLVE_SUCCESS;
      }
      f &= ~(TF_REFRESH_DYN | TF_REFRESH_GRAD);
      if (stationarity < sqrt(o.tol_stationarity)) {
        if constexpr (CON) {
          // z <- Pi(z_est) is applied knot by knot by the expansion below (dual_first)
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
        stopped = true;
      }
      P.flags[b] = f;
      fl = f;
    }
    const unsigned smask = __ballot_sync(kAll, stopped);
    if (lid == 0 && smask && done_out) atomicAdd(done_out, __popc(smask));
  }
  __syncthreads();
  if constexpr (CON) {
    // CalcProjectedDuals + CalcCostGradient after the dual / penalty update (solver.cpp:475-486),
    // for the problems that just stopped as well (the reference updates, then leaves the loop).
    // The flag is consumed here: a stopped problem must not be updated again while its group
    // mates keep iterating.
    fl = valid ? P.flags[bl] : 0;
    const unsigned gmask = __ballot_sync(kAll, (fl & TF_REFRESH_GRAD) != 0);
    if (gmask) {
      for_knot_items(gmask, g, P.N + 1, [&](int b, int k) {
        TS s(P, b);
        weights(s);
        s.rho = P.rho[b];
        s.phase_expand_knot(k, false, -1, true);
      });
      __syncthreads();
      if (wid == 0 && (fl & TF_REFRESH_GRAD)) P.flags[bl] = fl & ~TF_REFRESH_GRAD;
    }
  }
  tick(FS_CRITERIA);
  if (prof) atomicAdd(P.prof + FS_COUNT, 1ull);
}

// ALTROSolver::OpenLoopRollout (solver.cpp:116-131): x_[k+1] = f(x_[k], u_[k]) from the initial state
template <class Model, int CON>
__global__ void __launch_bounds__(32) k_open_loop_rollout(const __grid_constant__ DeviceProblem P) {
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

// Dense [A B] of every knot from the packed Jacobian rows (host views A_, B_ of KnotPointData)
template <class Model, int CON>
__global__ void __launch_bounds__(128) k_unpack_jac(const __grid_constant__ DeviceProblem P, double* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (b >= P.B) return;
  TrajSolver<Model, CON> s(P, b);
  constexpr int n = Model::n, m = Model::m, rows = n * n + n * m;
  double A[n * n], Bm[n * m];
  s.load_jac(k, A, Bm);
  double* o = out + ((long)(b >> 5) * (P.N + 1) + k) * rows * 32 + (b & 31);
  for (int e = 0; e < n * n; ++e) o[e * 32] = A[e];
  for (int e = 0; e < n * m; ++e) o[(n * n + e) * 32] = Bm[e];
}

// KnotPointData members that are not kept in HBM, re-created on demand for the host views
// (altro_b200_get_field): constraint_val_, z_proj_ (knotpoint_data.hpp:187-198) and the cost
// expansion lxx_, luu_, lux_ = CalcCostHessian (knotpoint_data.cpp:439-448) at the working
// trajectory with the stored z_est and the current rho.  out: [group][knot][rows][32].
// (enum KnotView: launchers.h)
template <class Model, int CON>
__global__ void __launch_bounds__(128) k_knot_view(const __grid_constant__ DeviceProblem P, int what, int rows,
                                                   double* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (b >= P.B) return;
  using TS = TrajSolver<Model, CON>;
  constexpr int n = Model::n, m = Model::m;
  TS s(P, b);
  s.rho = CON ? P.rho[b] : 1.0;
  const bool terminal = (k == P.N);
  double* o = out + ((long)(b >> 5) * (P.N + 1) + k) * rows * 32 + (b & 31);
  if (what == KV_RHO) {
    o[0] = s.rho;
  } else if (what == KV_LXX || what == KV_LUU || what == KV_LUX) {
    double lxx[n * n], luu[m * m], lux[m * n];
    for (int i = 0; i < m * m; ++i) luu[i] = 0.0;
    for (int i = 0; i < m * n; ++i) lux[i] = 0.0;
    s.cost_hessian(k, terminal, lxx, luu, lux);
    s.al_hessian(k, terminal, lxx, luu, lux);
    if (what == KV_LXX)
      for (int i = 0; i < n * n; ++i) o[i * 32] = lxx[i];
    if (what == KV_LUU)
      for (int i = 0; i < m * m; ++i) o[i * 32] = luu[i];
    if (what == KV_LUX)
      for (int i = 0; i < m * n; ++i) o[i * 32] = lux[i];
  } else if constexpr (CON != 0) {
    double x[n], u[m];
    load_block<n>(s.F(P.x), s.S, k, x);
    if (!terminal) {
      load_block<m>(s.F(P.u), s.S, k, u);
    } else {
      for (int i = 0; i < m; ++i) u[i] = 0.0;
    }
    const ConTable& T = P.contab;
    for (int j = 0; j < T.ncon; ++j) {
      const ConSlot& c = T.slot[j];
      if (k < c.k_start || k >= c.k_stop) continue;
      double val[kMaxConDim];
      if (what == KV_CONSTRAINT_VAL) {
        if constexpr (CON == 2) {
          s.con_eval(c, x, u, val, nullptr);
        } else {
          for (int i = 0; i < c.dim; ++i) val[i] = s.row_value(c, i, x, u);
        }
      } else {  // KV_Z_PROJ: projection of the stored z_est onto the dual cone
        double zt[kMaxConDim];
        const long zrow = s.zoff(k, c.row0);
        for (int i = 0; i < c.dim; ++i) zt[i] = P.zest[zrow + i * 32];
        if (CON == 2 && c.cone == CONE_SOC) {
          soc_projection(c.dim, zt, val);
        } else {
          for (int i = 0; i < c.dim; ++i) {
            double v = 0.0;
            if (c.cone == CONE_EQUALITY) v = zt[i];
            if (c.cone == CONE_INEQUALITY) v = fmin(0.0, zt[i]);
            val[i] = v;
          }
        }
      }
      for (int i = 0; i < c.dim; ++i) o[(c.row0 + i) * 32] = val[i];
    }
  }
}

// ALTROSolver::CalcCost (solver.cpp:163-174): sum_k cost(k) incl. the AL terms at the working
// trajectory; refreshes the projected duals like the reference does.
template <class Model, int CON>
__global__ void __launch_bounds__(32) k_calc_cost(const __grid_constant__ DeviceProblem P, double* out) {
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

#include "solver_team.cuh"
