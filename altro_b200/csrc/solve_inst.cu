// solve_inst.cu -- explicit instantiations of the solve kernel, one group per translation unit
// (compiled with -DALTRO_INST=<group>) so altro_b200/build.py can compile them in parallel.
#include <algorithm>

#include "launchers.h"
#include "solver_phases.cuh"

namespace altro_b200 {

// One operation of the phase pipeline for the sub-batch [P.g0, P.g0 + P.G): the Solve() prologue
// (OP_SOLVE_PROLOGUE) or iLQR iteration H->iter (OP_SOLVE_ITERATION = two launches).  Nothing here
// waits for the device; capi.cu enqueues iterations back to back and reads the stop counter late.
template <class Model, int CON>
static int run_phased(const DeviceProblem& P, cudaStream_t st, PhaseHost* H) {
  using TS = TrajSolver<Model, CON>;
  constexpr int n = Model::n, m = Model::m;
  const int B = P.B, N = P.N, G = P.G;
  auto g128 = [](int count) { return (count + 127) / 128; };
  // launch wrapper: counts launches/units and, in profile mode, times the kernel with events
  auto timed = [&](int phase, double units, auto&& launch) {
    if (H->profile) cudaEventRecord(H->ev0, st);
    launch();
    if (H->profile) {
      cudaEventRecord(H->ev1, st);
      cudaEventSynchronize(H->ev1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, H->ev0, H->ev1);
      H->ms[phase] += ms;
    }
    H->launches[phase] += 1;
    H->units[phase] += units;
  };
  const int b0 = P.g0 * 32, b1 = std::min(B, (P.g0 + G) * 32);  // problems of this sub-batch

  if (H->op == OP_SOLVE_PROLOGUE) {  // solver.cpp:417-430
    {
      const int mx = (int)H->smem_per_cta;
      if (TS::kStaged) {
        cudaFuncSetAttribute(k_phase_backward<Model, CON>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_phase_forward<Model, CON, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_phase_forward<Model, CON, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
      }
      cudaFuncSetAttribute(k_phase_backward_team<Model, CON>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    }
    cudaMemsetAsync(H->d_done, 0, sizeof(int), st);
    timed(PH_INIT, b1 - b0, [&] { k_phase_init<Model, CON><<<G, 32, 0, st>>>(P); });
    timed(PH_EXPAND, (double)G * 32 * (N + 1), [&] {  // with the OLD penalty (quirk Q3) ...
      k_phase_expand<Model, CON><<<dim3(g128(G * 32), N + 1), 128, 0, st>>>(P, G);
    });
    if (CON)  // ... then reset
      k_phase_set_rho<<<g128(b1 - b0), 128, 0, st>>>(P.rho, b0, b1, P.opts.penalty_initial);
    return (int)cudaGetLastError();
  }

  // ---- OP_SOLVE_ITERATION
  // TMA staging of the sequential sweeps (linalg.cuh): depth = knots in flight per CTA, as deep as
  // shared memory allows with `per_sm` CTAs of the launch resident (<= kMaxStageDepth); the cost
  // weights ride in shared memory behind the ring (stage_weights) when they are small.
  const int wdoubles = (N + 1) * n + N * m;
  const int wcount = (TS::kStaged && wdoubles * 8 <= 16 * 1024) ? wdoubles : 0;
  const size_t wbytes = (size_t)wcount * 8;
  auto ring_depth = [&](int per_sm, size_t stage_bytes, size_t fixed, int max_depth = kMaxStageDepth) -> int {
    const size_t budget =
        std::min<size_t>(H->smem_per_sm / std::max(per_sm, 1) - 1024, H->smem_per_cta) - fixed - wbytes;
    return (int)std::max<size_t>(2, std::min<size_t>((size_t)max_depth, budget / stage_bytes));
  };
  const int zr = CON ? 2 * P.zrows : 0;  // dual record rows staged with the main rows
  if (H->backward_team) {
    // Riccati sweep by the W warps of a CTA per group (solver_team.cuh)
    using Shape = TeamShape<n, m>;
    const int per_sm = (P.Gtot + H->num_sms - 1) / H->num_sms;
    const size_t xch = (size_t)Shape::kXchRows * 256;
    const size_t sweep_stage = (size_t)(TS::kRowsBw + zr) * 256;
    const size_t scan_stage = (size_t)(TS::kRowsBackwardKernel + zr) * 256;
    int depth, rdepth = 0;
    size_t stage_bytes;
    if (TS::kStaged) {
      depth = ring_depth(per_sm, sweep_stage, 256 + xch);
      rdepth = ring_depth(per_sm, scan_stage, 256 + 128 + xch);
      stage_bytes = std::max(BulkPipe::bytes(depth, (TS::kRowsBw + zr) * 32),
                             256 + BulkRing::bytes(rdepth, (TS::kRowsBackwardKernel + zr) * 32));
    } else {
      depth = ring_depth(1, sweep_stage, 256 + xch);  // one CTA per SM
      stage_bytes = BulkPipe::bytes(depth, (TS::kRowsBw + zr) * 32);
    }
    const size_t sm = stage_bytes + xch + wbytes;
    timed(PH_BACKWARD, (double)G * 32, [&] {
      k_phase_backward_team<Model, CON><<<G, 32 * Shape::W, sm, st>>>(P, depth, rdepth, wcount, H->iter == 0);
    });
  } else {
    int depth = 0;
    size_t sm = 0;
    if (TS::kStaged) {
      const int rows = TS::kRowsBackwardKernel + zr;
      // budget for every group of the HANDLE resident at once (other sub-batches' kernels share
      // the SMs)
      depth = ring_depth((P.Gtot + H->num_sms - 1) / H->num_sms, (size_t)rows * 256, 128);
      sm = BulkRing::bytes(depth, rows * 32) + wbytes;
    }
    timed(PH_BACKWARD, (double)G * 32, [&] {
      k_phase_backward<Model, CON><<<G, 32, sm, st>>>(P, depth, wcount, H->iter == 0);
    });
  }
  {
    int depth = 0, rows = 0;
    size_t sm = 0;
    // follower variant: the derivative half of a merit evaluation by a warp behind the rollout warp
    const bool follow = TS::kStaged && P.follow_deriv != 0;
    if (TS::kStaged) {
      // stage rows: the rollout's [xbar ubar q r c K d] (+ duals) plus, in follower mode, x_k, u_k of
      // the rollout warp; otherwise the d(phi) scan's [K d] [J] [lx lu] must fit as well
      rows = follow ? TS::kRowsRoll + zr + n + m : std::max(TS::kRowsRoll + zr, TS::kRowsDphi);
      // behind the ring: cost weights, ready[] counters, the resident [q r c] rows of knot 0, x_N and phi'
      const size_t tail = (size_t)((N + 1) * 4 + 15) / 16 * 16 + (size_t)(TS::rK - TS::rQ + n + 1) * 256;
      // The forward CTAs are register-limited to two per SM, so the ring may use up to half an SM's
      // shared memory minus what a co-resident backward CTA needs: up to kFwdStageDepth knots in
      // flight (ncu r02e: a third of the kernel's stall samples waited for a stage to land at depth 4)
      const int per_sm = std::min(kFwdCtasPerSm, (P.Gtot + H->num_sms - 1) / H->num_sms);
      depth = ring_depth(per_sm, (size_t)rows * 256, 256 + tail + (size_t)(48 * 1024) / std::max(per_sm, 1), H->fwd_depth);
      sm = BulkPipe::bytes(depth, rows * 32) + wbytes + tail;
    }
    // 192 threads at most (256 for the large-block models); the follower variant needs the rollout
    // warp, the follower and the speculating warps
    const int warps = std::max(follow ? 2 : 1, std::min(H->fwd_warps + (follow ? 1 : 0), TS::kStaged ? kFwdThreads / 32 : 8));
    DeviceProblem Pf = P;
    Pf.prof = H->profile ? H->d_prof : nullptr;
    if (H->profile) cudaMemsetAsync(H->d_prof, 0, 16 * sizeof(unsigned long long), st);
    const double ms_before = H->ms[PH_FORWARD];
    timed(PH_FORWARD, (double)G * 32, [&] {
      if (follow)
        k_phase_forward<Model, CON, true><<<G, 32 * warps, sm, st>>>(Pf, depth, rows, wcount, H->d_done);
      else
        k_phase_forward<Model, CON, false><<<G, 32 * warps, sm, st>>>(Pf, depth, rows, wcount, H->d_done);
    });
    if (H->profile) {  // split the kernel's time by the in-kernel sub-phase clocks
      unsigned long long ns[16];
      cudaMemcpy(ns, H->d_prof, sizeof(ns), cudaMemcpyDeviceToHost);
      if (getenv("ALTRO_B200_PROF_DUMP"))
        fprintf(stderr,
                "fwd prof iter %d: ctas %llu  ns rollout %llu expand %llu dphi_ls %llu criteria %llu | rollout warp cycles: "
                "wait-full %llu release %llu pass %llu | pass cycles by round %llu %llu %llu %llu passes %llu %llu %llu %llu\n",
                H->iter, ns[4], ns[0], ns[1], ns[2], ns[3], ns[5], ns[6], ns[7], ns[8], ns[9], ns[10], ns[11], ns[12],
                ns[13], ns[14], ns[15]);
      const double tot = (double)(ns[0] + ns[1] + ns[2] + ns[3]);
      const double ms = H->ms[PH_FORWARD] - ms_before;
      for (int i = 0; i < 4; ++i) {
        if (tot > 0) H->ms[PH_FWD_ROLLOUT + i] += ms * (double)ns[i] / tot;
        H->launches[PH_FWD_ROLLOUT + i] += 1;
      }
    }
  }
  return (int)cudaGetLastError();
}

// has_con: 0 unconstrained, 1 linear cones only, 2 with second-order cones (TrajSolver's CON).
// The secondary entry points (open-loop rollout, cost, persistent twin) only distinguish
// unconstrained / constrained and use the general instantiation for the latter.
template <class Model>
static int launch_solve(const DeviceProblem& P, int has_con, cudaStream_t st, PhaseHost* host) {
  if (host && host->op == OP_OPEN_LOOP_ROLLOUT) {
    if (has_con)
      k_open_loop_rollout<Model, 2><<<(P.B + 31) / 32, 32, 0, st>>>(P);
    else
      k_open_loop_rollout<Model, 0><<<(P.B + 31) / 32, 32, 0, st>>>(P);
    return (int)cudaGetLastError();
  }
  if (host && host->op == OP_CALC_COST) {
    if (has_con)
      k_calc_cost<Model, 2><<<(P.B + 31) / 32, 32, 0, st>>>(P, host->cost_out);
    else
      k_calc_cost<Model, 0><<<(P.B + 31) / 32, 32, 0, st>>>(P, host->cost_out);
    return (int)cudaGetLastError();
  }
  if (host && host->op == OP_KNOT_EVAL) {
    const dim3 grid((P.G * 32 + 127) / 128, P.N + 1);
    if (has_con == 0)
      k_phase_expand<Model, 0><<<grid, 128, 0, st>>>(P, P.G);
    else if (has_con == 1)
      k_phase_expand<Model, 1><<<grid, 128, 0, st>>>(P, P.G);
    else
      k_phase_expand<Model, 2><<<grid, 128, 0, st>>>(P, P.G);
    return (int)cudaGetLastError();
  }
  if (host && host->op == OP_KNOT_VIEW) {
    const dim3 grid((P.B + 127) / 128, P.N + 1);
    if (has_con == 0)
      k_knot_view<Model, 0><<<grid, 128, 0, st>>>(P, host->view, host->view_rows, host->cost_out);
    else if (has_con == 1)
      k_knot_view<Model, 1><<<grid, 128, 0, st>>>(P, host->view, host->view_rows, host->cost_out);
    else
      k_knot_view<Model, 2><<<grid, 128, 0, st>>>(P, host->view, host->view_rows, host->cost_out);
    return (int)cudaGetLastError();
  }
  if (host && host->op == OP_UNPACK_JAC) {
    k_unpack_jac<Model, 0><<<dim3((P.B + 127) / 128, P.N), 128, 0, st>>>(P, host->cost_out);
    return (int)cudaGetLastError();
  }
  if (host) {
    if (has_con == 0) return run_phased<Model, 0>(P, st, host);
    if (has_con == 1) return run_phased<Model, 1>(P, st, host);
    return run_phased<Model, 2>(P, st, host);
  }
  const int threads = 32;  // one warp per CTA: 32 consecutive problems
  const int blocks = (P.B + threads - 1) / threads;
  if (has_con)
    solve_kernel<Model, 2><<<blocks, threads, 0, st>>>(P);
  else
    solve_kernel<Model, 0><<<blocks, threads, 0, st>>>(P);
  return (int)cudaGetLastError();
}

#define ALTRO_DEFINE_LAUNCHER(name, ...)                                                        \
  int name(const DeviceProblem& P, int has_constraints, cudaStream_t stream, PhaseHost* host) { \
    return launch_solve<__VA_ARGS__>(P, has_constraints, stream, host);                         \
  }

#if ALTRO_INST == 0
ALTRO_DEFINE_LAUNCHER(launch_solve_linear_4_2, LinearModel<4, 2>)
ALTRO_DEFINE_LAUNCHER(launch_solve_linear_2_1, LinearModel<2, 1>)
ALTRO_DEFINE_LAUNCHER(launch_solve_di_1, DoubleIntegrator<1>)
ALTRO_DEFINE_LAUNCHER(launch_solve_di_2, DoubleIntegrator<2>)
#elif ALTRO_INST == 1
ALTRO_DEFINE_LAUNCHER(launch_solve_pendulum, Pendulum)
ALTRO_DEFINE_LAUNCHER(launch_solve_bicycle4, Bicycle4)
#elif ALTRO_INST == 2
ALTRO_DEFINE_LAUNCHER(launch_solve_bicycle5, Bicycle5)
#elif ALTRO_INST == 3
ALTRO_DEFINE_LAUNCHER(launch_solve_chain_4_2, Chain<4, 2>)
ALTRO_DEFINE_LAUNCHER(launch_solve_chain_4_4, Chain<4, 4>)
#elif ALTRO_INST == 4
ALTRO_DEFINE_LAUNCHER(launch_solve_chain_6_2, Chain<6, 2>)
ALTRO_DEFINE_LAUNCHER(launch_solve_chain_6_4, Chain<6, 4>)
#elif ALTRO_INST == 5
ALTRO_DEFINE_LAUNCHER(launch_solve_chain_12_2, Chain<12, 2>)
#elif ALTRO_INST == 6
ALTRO_DEFINE_LAUNCHER(launch_solve_chain_12_4, Chain<12, 4>)
#elif ALTRO_INST == 7
ALTRO_DEFINE_LAUNCHER(launch_solve_linear_6_3, LinearModel<6, 3>)
ALTRO_DEFINE_LAUNCHER(launch_solve_di_3, DoubleIntegrator<3>)
#else
#error "ALTRO_INST must be 0..7"
#endif

}  // namespace altro_b200
