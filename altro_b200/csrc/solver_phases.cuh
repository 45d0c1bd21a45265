// solver_phases.cuh -- the AL-iLQR iteration as a pipeline of phase kernels over compacted lists of
// GROUPS (the production path; the single persistent kernel of solver_kernels.cuh stays as the
// differential-testing twin: both must agree bit for bit).
//
// A group is 32 consecutive problems = one warp = one stream of knot records (device_problem.h).
//   * Every sequential sweep (backward Riccati, phi0 scan, rollout, d(phi) scan) is a kernel with
//     one warp per group that streams the group's knot records through a shared-memory ring with
//     TMA bulk copies (linalg.cuh, BulkRing): one cp.async.bulk per knot and contiguous range, a
//     few knots ahead, instead of dozens of dependent 256-byte loads per knot.  The pipeline is
//     HBM-bound (ncu r01 v2: 50-65 % of DRAM peak in every kernel at only 3.5 warps/SM), so what
//     counts is bytes per knot and how fast a lone warp can stream them.
//   * Lists are compacted per GROUP (a group stays listed while any of its lanes needs the
//     phase); lanes that do not need it are predicated off.  Records are fetched whole anyway.
//   * Everything that is independent per knot -- dynamics Jacobians (the transcendental-heavy
//     part), projected duals, cost gradients, costates, residuals -- runs one thread per
//     (problem, knot).
//   * Backtracking line search: the first rollout round evaluates the requested step AND the
//     halvings that SimpleBacktracking (linesearch.cpp:385-412) would try next, one warp per
//     candidate in the SAME CTA, all fed from one staged copy of the knot data; only slot 0 and
//     the first `nstore` halvings write their trajectory, the others return the merit value
//     alone (an accepted one is rolled out once more, TF_REROLL).  The state machine is then fed
//     the values in the order the reference would have evaluated them, so decisions and the
//     reported evaluation counts are those of the sequential search.
//   * The redundant alpha = 0 rollout of ForwardPass (solver.cpp:241, ~40 % of the reference's
//     merit evaluations) is replaced by a linear scan that provably reproduces it
//     (TrajSolver::phi0_step).
// Decisions (line search, dual/penalty update, convergence) are identical to the reference.
#pragma once
#include "solver_kernels.cuh"

namespace altro_b200 {

// ------------------------------------------------------------------ list compaction (groups)
// Ordered compaction of the groups in[0..count) that have at least one lane with
// (flags & mask) != 0 into out; writes the number kept to counters[slot], and the number of kept
// groups that also have such a lane with mask2 / mask3 to counters[slot2] / counters[slot3].
// `count` bounds the input length; `dcount` (optional) is its exact device-side value.  One CTA.
static __global__ void __launch_bounds__(1024) k_compact(const int* __restrict__ in, int count,
                                                  const int* dcount,
                                                  const int* __restrict__ flags, int mask,
                                                  int* __restrict__ out, int* counters, int slot,
                                                  int mask2, int slot2, int mask3, int slot3,
                                                  int gbase) {
  __shared__ int warp_tot[32], warp_tot2[32], warp_tot3[32];
  __shared__ int base_s, base2_s, base3_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // the input length may live in the very counter this kernel rewrites at the end, hence read
  // before the first barrier
  if (dcount) count = min(count, *dcount);
  if (tid == 0) {
    base_s = 0;
    base2_s = 0;
    base3_s = 0;
  }
  __syncthreads();
  for (int start = 0; start < count; start += 1024) {
    const int i = start + tid;
    int g = -1, keep = 0, keep2 = 0, keep3 = 0;
    if (i < count) {
      g = in ? in[i] : gbase + i;
      int f = 0;  // OR of the flags of the lanes that have `mask`
      const int4* fp = reinterpret_cast<const int4*>(flags + (long)g * 32);
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        const int4 v = fp[l];
        f |= (v.x & mask) ? v.x : 0;
        f |= (v.y & mask) ? v.y : 0;
        f |= (v.z & mask) ? v.z : 0;
        f |= (v.w & mask) ? v.w : 0;
      }
      keep = (f & mask) != 0;
      keep2 = keep && mask2 && (f & mask2) != 0;
      keep3 = keep && mask3 && (f & mask3) != 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const unsigned bal2 = __ballot_sync(0xffffffffu, keep2);
    const unsigned bal3 = __ballot_sync(0xffffffffu, keep3);
    if (lane == 0) {
      warp_tot[wid] = __popc(bal);
      warp_tot2[wid] = __popc(bal2);
      warp_tot3[wid] = __popc(bal3);
    }
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < wid; ++w) off += warp_tot[w];
    if (keep) out[off + __popc(bal & ((1u << lane) - 1u))] = g;
    __syncthreads();
    if (tid == 0) {
      int t = 0, t2 = 0, t3 = 0;
      for (int w = 0; w < 32; ++w) {
        t += warp_tot[w];
        t2 += warp_tot2[w];
        t3 += warp_tot3[w];
      }
      base_s += t;
      base2_s += t2;
      base3_s += t3;
    }
    __syncthreads();
  }
  if (tid == 0) {
    counters[slot] = base_s;
    if (mask2) counters[slot2] = base2_s;
    if (mask3) counters[slot3] = base3_s;
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

// Expansion, one thread per (problem of a listed group, knot).  `mask`: only problems whose flags
// have one of these bits are processed (0 = all).  with_dyn: also recompute [A B].  slot_mode: -1
// read the main trajectory, >= 0 that candidate slot, -2 the slot recorded in sel[b].
// dual_first: apply the dual update z <- Pi(z_est) of this knot before recomputing the projected
// duals.  The prologue calls this BEFORE the penalty reset, which reproduces quirk Q3 (gradient
// with the old rho, solver.cpp:424-430).
template <class Model, int CON>
__global__ void __launch_bounds__(128) k_phase_expand(const __grid_constant__ DeviceProblem P, const int* list,
                                                      int count, const int* dcount, int mask,
                                                      bool with_dyn, int slot_mode,
                                                      bool dual_first) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  const int gi = t >> 5;
  if (gi >= list_count(count, dcount)) return;
  const int b = (list ? list[gi] : P.g0 + gi) * 32 + (t & 31);
  if (b >= P.B) return;
  if (mask && !(P.flags[b] & mask)) return;
  TrajSolver<Model, CON> s(P, b);
  s.rho = CON ? P.rho[b] : 1.0;
  const int slot = (slot_mode == -2) ? P.sel[b] : slot_mode;
  s.phase_expand_knot(k, with_dyn, slot, dual_first);
}

// penalty reset at the end of the prologue (SetPenalty(penalty_initial), solver.cpp:429)
static __global__ void k_phase_set_rho(double* rho, int b0, int b1, double value) {
  const int b = b0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (b < b1) rho[b] = value;
}

// K1: CalcExpansions + BackwardPass + the alpha = 0 half of ForwardPass (solver.cpp:448-450,
// :241-245) and the start of the line search.  One warp per listed group.
template <class Model, int CON>
__global__ void __launch_bounds__(32) k_phase_backward(const __grid_constant__ DeviceProblem P, const int* list,
                                                       int count, int depth, int wcount, int first) {
  using TS = TrajSolver<Model, CON>;
  constexpr int n = Model::n, m = Model::m;
  if ((int)blockIdx.x >= count) return;
  const int g = list[blockIdx.x];
  const int lane = threadIdx.x;
  const int b = g * 32 + lane;
  const bool active = b < P.B && (P.flags[b] & TF_ACTIVE);
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
        s.unstage_jac(st, 0, lane, A, Bm);
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
          s.unstage_jac(st, kRows1, lane, A, Bm);
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
        s.unstage_jac(st, oJ, lane, A, Bm);
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
  } else {
    if (active) {
      s.backward_sweep();
      s.phase_phi0_scan(&phi0, &dphi0);
    }
  }
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
      if (lo.use_backtracking && P.nslots > 1) {
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

// K2: rollouts.  One CTA per listed group, one warp per candidate step of the group's problems:
// warp 0 rolls out the step the state machine asked for (alpha_eval) into the main trajectory;
// warps j >= 1 (launched in speculative rounds) roll out the halving alpha_bt * 2^-(j-1) for the
// lanes flagged TF_SPECULATE -- into candidate slot j-1 when j <= nstore, merit value only
// otherwise.  All warps consume the SAME staged copy of the knot data [xbar ubar q r c K d].
template <class Model, int CON>
__global__ void __launch_bounds__(32 * 16) k_phase_rollout(const __grid_constant__ DeviceProblem P, const int* list,
                                                           int count, const int* dcount, int depth,
                                                           int wcount) {
  using TS = TrajSolver<Model, CON>;
  constexpr int n = Model::n, m = Model::m;
  if ((int)blockIdx.x >= list_count(count, dcount)) return;
  const int g = list[blockIdx.x];
  // Work assignment.  Warp 0: thread = problem lane, candidate 0 (the requested step).  Warps
  // >= 1: the (lane, halving) pairs of the lanes flagged TF_SPECULATE, PACKED densely over the
  // threads -- in the first round every lane speculates and warp j is simply halving j, but in the
  // later rounds only a few lanes per group still need halvings and all their candidates fit in
  // one or two warps instead of keeping nslots-1 mostly idle warps busy for the whole sweep.
  // Pair p -> lane rank p % nneedy (consecutive threads = different lanes: conflict-free smem
  // columns, neighbouring global stores), halving p / nneedy + 1.
  int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
  const int nspec = (int)(blockDim.x >> 5) - 1;
  bool need;
  {
    const int bl = g * 32 + lane;
    const int fl = bl < P.B ? P.flags[bl] : 0;
    const unsigned needy = __ballot_sync(0xffffffffu, (fl & TF_NEED_EVAL) && (fl & TF_SPECULATE));
    if (slot == 0) {
      need = (fl & (TF_NEED_EVAL | TF_REROLL)) != 0;
    } else {
      const int nneedy = __popc(needy);
      const int p = (slot - 1) * 32 + lane;
      need = nneedy > 0 && p < nneedy * nspec;
      if (need) {
        lane = __fns(needy, 0, p % nneedy + 1);  // the (p % nneedy)-th flagged lane
        slot = p / nneedy + 1;
      }
    }
  }
  const int b = g * 32 + lane;
  TS s(P, need ? b : g * 32);
  s.rho = (CON && need) ? P.rho[b] : 1.0;
  double alpha = 0.0;
  double *xo = nullptr, *uo = nullptr;
  long so = 0;
  if (need) {
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
    const int zr = CON ? 2 * P.zrows : 0;
    BulkRing ring;
    ring.init(altro_smem, depth, (kRows + zr) * 32, threadIdx.x == 0);
    stage_weights(s, P, reinterpret_cast<double*>(altro_smem + BulkRing::bytes(depth, (kRows + zr) * 32)), wcount);
    __syncthreads();
    const double* rec = P.xbar + (long)g * P.GS;
    const double* zrec = CON ? P.z + (long)g * P.GSz : nullptr;
    auto fetch = [&](int k, int st) {
      ring.expect(st, (kRows + zr) * 256);
      ring.copy(st, 0, rec + (long)k * P.R, kRows * 256);
      if (zr) ring.copy(st, kRows, zrec + (long)k * P.Rz, zr * 256);
    };
    if (threadIdx.x == 0)
      for (int j = 0; j < depth; ++j)
        if (j < P.N) fetch(j, j);
    double x[n];
    if (need) load_block<n>(s.G(P.x0, n), 0, 0, x);
    for (int k = 0; k < P.N; ++k) {
      const double* st = ring.wait();
      double xb[n], ub[m], q[n], r[m], K[m * n], d[m], cval = 0.0;
      if (need) {
        unstage_block<n>(st, TS::rXbar, lane, xb);
        unstage_block<m>(st, TS::rUbar, lane, ub);
        unstage_block<n>(st, TS::rQ, lane, q);
        unstage_block<m>(st, TS::rR, lane, r);
        cval = st[TS::rC * 32 + lane];
        unstage_block<m * n>(st, TS::rK, lane, K);
        unstage_block<m>(st, TS::rD, lane, d);
      }
      if (zr) s.zstage = st + kRows * 32 + lane;
      if (need) s.rollout_step(k, alpha, xb, ub, K, d, q, r, cval, x, xo, uo, so, phi);
      s.zstage = nullptr;
      __syncthreads();
      if (threadIdx.x == 0 && k + depth < P.N) fetch(k + depth, ring.s);
      ring.advance();
    }
    if (need) s.rollout_terminal(x, xo, so, phi);
  } else {
    if (need) phi = s.phase_rollout(alpha, xo, uo, so);
  }
  if (need) {
    if (slot == 0)
      P.phi_eval[b] = phi;
    else
      P.phi_s[(long)min(P.spec_base[b] + slot - 1, kMaxHalvings) * P.Bp + b] = phi;
  }
}

// K4: d(phi) scan (when requested) + the line-search state machine, one warp per listed group.
// The machine is fed the value of the requested step and then, while it keeps backtracking, the
// precomputed merit values of the halvings in the order the reference would have evaluated
// them; feeding stops at the first one it accepts, so the decisions (and the reported evaluation
// count) are those of the sequential search.
template <class Model, int CON>
__global__ void __launch_bounds__(32) k_phase_lsupdate(const __grid_constant__ DeviceProblem P, const int* list,
                                                       int count, const int* dcount, int depth,
                                                       int nspec) {
  using TS = TrajSolver<Model, CON>;
  constexpr int n = Model::n, m = Model::m;
  if ((int)blockIdx.x >= list_count(count, dcount)) return;
  const int g = list[blockIdx.x];
  const int lane = threadIdx.x;
  const int b = g * 32 + lane;
  int f = b < P.B ? P.flags[b] : 0;
  const bool pending = (f & (TF_NEED_EVAL | TF_REROLL)) != 0;
  const bool had_deriv = (f & TF_NEED_EVAL) && (f & TF_WANT_DERIV) && !(f & TF_REROLL);
  double dphi = 0.0;
  if (__any_sync(0xffffffffu, had_deriv)) {
    TS s(P, had_deriv ? b : g * 32);
    if constexpr (TS::kStaged) {
      // stage contents: [K d] [J] [lx lu]
      constexpr int kV = TS::kV;
      constexpr int kRows1 = m * n + m;
      BulkRing ring;
      ring.init(altro_smem, depth, TS::kRowsDphi * 32, lane == 0);
      __syncwarp();
      const double* rec = P.xbar + (long)g * P.GS;
      auto fetch = [&](int k, int st) {
        ring.expect(st, TS::kRowsDphi * 256);
        ring.copy(st, 0, rec + (long)k * P.R + TS::rK * 32, kRows1 * 256);
        ring.copy(st, kRows1, rec + (long)k * P.R + TS::rA * 32, kV * 256);
        ring.copy(st, kRows1 + kV, rec + (long)k * P.R + TS::rLx * 32, (n + m) * 256);
      };
      if (lane == 0)
        for (int j = 0; j < depth; ++j)
          if (j < P.N) fetch(j, j);
      double dxda[n];
#pragma unroll
      for (int i = 0; i < n; ++i) dxda[i] = 0.0;
      for (int k = 0; k < P.N; ++k) {
        const double* st = ring.wait();
        double K[m * n], d[m], A[n * n], Bm[n * m], lx[n], lu[m];
        if (had_deriv) {
          unstage_block<m * n>(st, 0, lane, K);
          unstage_block<m>(st, m * n, lane, d);
          s.unstage_jac(st, kRows1, lane, A, Bm);
          unstage_block<n>(st, kRows1 + kV, lane, lx);
          unstage_block<m>(st, kRows1 + kV + n, lane, lu);
        }
        if (had_deriv) s.dphi_step(K, d, A, Bm, lx, lu, dxda, dphi);
        __syncwarp();
        if (lane == 0 && k + depth < P.N) fetch(k + depth, ring.s);
        ring.advance();
      }
      if (had_deriv) dphi = s.dphi_terminal(dxda, dphi);
    } else {
      if (had_deriv) dphi = s.phase_dphi_scan();
    }
  }
  if (!pending) return;
  const LsOptions lo = ls_options(P.opts);
  if (f & TF_REROLL) {
    // the accepted candidate has just been rolled out again into x_, u_; it still needs its
    // expansion (TF_REFRESH_DYN stays set)
    f &= ~(TF_REROLL | TF_NEED_EVAL | TF_WANT_DERIV | TF_SPECULATE);
    P.sel[b] = -1;
    P.flags[b] = f;
    return;
  }
  LsMachine ls = P.ls[b];
  bool last_had_deriv = had_deriv;
  // nspec: halvings the round just executed rolled out besides the request (warps - 1)
  int known = P.spec_known[b];     // halvings 1..known have their merit value in phi_s
  const int base = P.spec_base[b];
  const bool was_backtrack = ls.phase == LsMachine::P_BACKTRACK;
  int fed = 1, winner = 0;         // winner: halving index the search returned (0: the request)
  ls.update(lo, P.phi_eval[b], dphi);
  if (f & TF_SPECULATE) {
    // this round produced halvings base .. base+nspec-1; the request itself was halving base-1
    // when the machine was already backtracking
    if (was_backtrack && base >= 2) P.phi_s[(long)min(base - 1, kMaxHalvings) * P.Bp + b] = P.phi_eval[b];
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
    if (lo.use_backtracking && P.nslots > 1) {
      if (ls.phase == LsMachine::P_BACKTRACK) {
        // ran out of precomputed halvings: the request is halving known+1, speculate the next ones
        f |= TF_SPECULATE;
        P.spec_base[b] = known + 2;
        P.alpha_bt[b] = ls.alpha * lo.beta_decrease;
      } else if (ls.phase == LsMachine::P_CUBIC_FIRST) {
        // the cubic-first probe is next (linesearch.cpp:96-127).  If it is rejected the search
        // backtracks through the halvings; when none of the known ones passes the Armijo test the
        // ones after them ride along with the probe
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
      // the step the search returns was only evaluated as a speculative candidate of the round
      // that started at halving `wbase`
      const int wbase = (winner >= base) ? base : 1;
      const int wslot = winner - wbase + 1;
      if (winner >= base && wslot >= 1 && wslot <= P.nstore)
        P.sel[b] = wslot - 1;  // its trajectory is in a candidate slot: copy + expand it
      else
        f |= TF_REROLL;        // merit-only / overwritten candidate: roll it out again, then expand
      f |= TF_REFRESH_DYN;
    } else if (lo.use_backtracking && fabs(alpha - 1.0) > 0 && !last_had_deriv) {
      f |= TF_REFRESH_DYN;  // accepted point has no derivative information yet (solver.cpp:256-262)
    }
    if (isnan(alpha) || !(ls.status == LS_MINIMUM_FOUND || ls.status == LS_HIT_MAX_STEPSIZE))
      f |= TF_LS_FAILED;
  }
  P.ls[b] = ls;
  P.flags[b] = f;
}

// K5a/b: costates of the accepted point, then stationarity / feasibility residuals and
// CopyTrajectory -- one thread per (problem of a listed group, knot)
template <class Model, int CON>
__global__ void __launch_bounds__(128) k_phase_costate(const __grid_constant__ DeviceProblem P, const int* list,
                                                       int count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 5) >= count) return;
  const int b = list[t >> 5] * 32 + (t & 31);
  if (b >= P.B || !(P.flags[b] & TF_ACTIVE)) return;
  TrajSolver<Model, CON> s(P, b);
  s.phase_costate_knot(blockIdx.y);
}

template <class Model, int CON>
__global__ void __launch_bounds__(128) k_phase_residual(const __grid_constant__ DeviceProblem P, const int* list,
                                                        int count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 5) >= count) return;
  const int b = list[t >> 5] * 32 + (t & 31);
  if (b >= P.B || !(P.flags[b] & TF_ACTIVE)) return;
  TrajSolver<Model, CON> s(P, b);
  s.phase_residual_knot(blockIdx.y);
}

// K5c: convergence test, dual / penalty update decision (solver.cpp:459-489, :503-506)
template <int CON>
__global__ void __launch_bounds__(128) k_phase_decide(const __grid_constant__ DeviceProblem P, const int* list,
                                                      int count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 5) >= count) return;
  const int b = list[t >> 5] * 32 + (t & 31);
  if (b >= P.B || !(P.flags[b] & TF_ACTIVE)) return;
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
