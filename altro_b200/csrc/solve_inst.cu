// solve_inst.cu -- explicit instantiations of the solve kernel, one group per translation unit
// (compiled with -DALTRO_INST=<group>) so altro_b200/build.py can compile them in parallel.
#include <algorithm>

#include "launchers.h"
#include "solver_phases.cuh"

namespace altro_b200 {

template <class Model, int CON>
static int run_phased(const DeviceProblem& P, cudaStream_t st, PhaseHost* H) {
  using TS = TrajSolver<Model, CON>;
  constexpr int n = Model::n, m = Model::m;
  const int B = P.B, N = P.N, G = P.G;
  auto g128 = [](int count) { return (count + 127) / 128; };
  cudaError_t err = cudaSuccess;
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
  auto readback = [&]() {
    cudaMemcpyAsync(H->h_counters, P.counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, st);
    err = cudaStreamSynchronize(st);
    H->syncs += 1;
  };
  // lists hold GROUP ids; `count` = groups
  auto expand = [&](const int* list, int count, const int* dcount, int mask, bool with_dyn,
                    int slot_mode, bool dual_first) {
    timed(PH_EXPAND, (double)count * 32 * (N + 1), [&] {
      k_phase_expand<Model, CON><<<dim3(g128(count * 32), N + 1), 128, 0, st>>>(
          P, list, count, dcount, mask, with_dyn, slot_mode, dual_first);
    });
  };
  auto compact = [&](const int* in, int count, const int* dcount, int mask, int* out, int slot,
                     int mask2, int slot2, int mask3, int slot3) {
    timed(PH_COMPACT, count, [&] {
      k_compact<<<1, 1024, 0, st>>>(in, count, dcount, P.flags, mask, out, P.counters, slot, mask2,
                                    slot2, mask3, slot3, P.g0);
    });
  };
  // TMA staging ring of the sequential sweeps (linalg.cuh): depth = knots in flight per CTA, as
  // deep as shared memory allows with every CTA of the launch resident (<= kMaxStageDepth).
  // cost weights staged in shared memory behind the ring (stage_weights) when they are small
  const int wdoubles = (N + 1) * n + N * m;
  const int wcount = (TS::kStaged && wdoubles * 8 <= 16 * 1024) ? wdoubles : 0;
  auto ring = [&](int ctas, int stage_rows, int* depth, bool weights) -> size_t {
    if (!TS::kStaged) {
      *depth = 0;
      return 0;
    }
    const size_t wbytes = weights ? (size_t)wcount * 8 : 0;
    const size_t stage_bytes = (size_t)stage_rows * 256;
    const int per_sm = (ctas + H->num_sms - 1) / H->num_sms;
    const size_t budget =
        std::min<size_t>(H->smem_per_sm / std::max(per_sm, 1) - 1024, H->smem_per_cta) - 128 - wbytes;
    *depth = (int)std::max<size_t>(2, std::min<size_t>(kMaxStageDepth, budget / stage_bytes));
    return BulkRing::bytes(*depth, stage_rows * 32) + wbytes;
  };
  if (TS::kStaged) {
    const int mx = (int)H->smem_per_cta;
    cudaFuncSetAttribute(k_phase_backward<Model, CON>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_phase_rollout<Model, CON>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_phase_lsupdate<Model, CON>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  }
  const int zr = CON ? 2 * P.zrows : 0;  // dual record rows staged with the main rows
  const int kRowsBackward = TS::kRowsBackwardKernel + zr;
  const int kRowsRollout = TS::kRowsRoll + zr;
  constexpr int kRowsDphi = TS::kRowsDphi;
  // warps = candidate steps rolled out per group (1 = only the requested step)
  auto rollout = [&](const int* list, int count, const int* dcount, int warps) {
    int depth;
    const size_t sm = ring(count, kRowsRollout, &depth, true);
    timed(PH_ROLLOUT, (double)count * 32 * warps, [&] {
      k_phase_rollout<Model, CON><<<count, 32 * warps, sm, st>>>(P, list, count, dcount, depth, wcount);
    });
  };
  auto lsupdate = [&](const int* list, int count, const int* dcount, int warps) {
    int depth;
    const size_t sm = ring(count, kRowsDphi, &depth, false);
    timed(PH_LSUPDATE, (double)count * 32, [&] {
      k_phase_lsupdate<Model, CON><<<count, 32, sm, st>>>(P, list, count, dcount, depth, warps - 1);
    });
  };

  // ---- prologue (solver.cpp:417-430)
  const int b0 = P.g0 * 32, b1 = std::min(B, (P.g0 + G) * 32);  // problems of this sub-batch
  timed(PH_INIT, b1 - b0, [&] { k_phase_init<Model, CON><<<G, 32, 0, st>>>(P); });
  expand(nullptr, G, nullptr, 0, true, -1, false);  // with the OLD penalty (quirk Q3) ...
  if (CON)  // ... then reset
    k_phase_set_rho<<<g128(b1 - b0), 128, 0, st>>>(P.rho, b0, b1, P.opts.penalty_initial);
  int* list_iter = P.list_iter;
  int* list_iter_next = H->list_aux;
  compact(nullptr, G, nullptr, TF_ACTIVE, list_iter, PC_ITER, 0, 0, 0, 0);
  int count_iter = G;
  const bool backtracking = P.opts.use_backtracking_linesearch != 0;
  // candidate steps per round.  (Going wider in the late rounds that only serve the lanes that
  // keep halving was measured and lost: nearly every group has such a lane, so a 16-warp tail
  // round costs more issue slots than the round it saves -- 29.2 vs 21.9 ms of rollouts.)
  const int spec_warps = (backtracking && P.nslots > 1) ? std::min(P.nslots, 16) : 1;
  const int LS_MASK = TF_NEED_EVAL | TF_REROLL;

  for (int iter = 0; iter < P.opts.iterations_max && count_iter > 0; ++iter) {
    {
      int depth;
      const size_t sm = ring(count_iter, kRowsBackward, &depth, true);
      timed(PH_BACKWARD, (double)count_iter * 32, [&] {
        k_phase_backward<Model, CON><<<count_iter, 32, sm, st>>>(P, list_iter, count_iter, depth, wcount,
                                                                 iter == 0);
      });
    }
    int* cur = P.list_ls;
    int* nxt = P.list_tmp;
    compact(list_iter, count_iter, nullptr, LS_MASK, cur, PC_LS, TF_WANT_DERIV, PC_DERIV, TF_SPECULATE, PC_SPEC);
    // round 1: the requested step (alpha0 = 1, with derivative) for every problem still searching,
    // plus, for the backtracking search, the halvings it would try next.  The exact list length
    // stays on the device (no host round trip); count_iter bounds the grid.
    const int* dcount = P.counters + PC_LS;
    rollout(cur, count_iter, dcount, spec_warps);
    expand(cur, count_iter, dcount, TF_WANT_DERIV, true, -1, false);
    lsupdate(cur, count_iter, dcount, spec_warps);
    compact(cur, count_iter, dcount, LS_MASK, nxt, PC_LS, TF_WANT_DERIV, PC_DERIV, TF_SPECULATE, PC_SPEC);
    readback();
    if (err != cudaSuccess) return (int)err;
    int count_ls = H->h_counters[PC_LS], count_deriv = H->h_counters[PC_DERIV],
        count_spec = H->h_counters[PC_SPEC];
    std::swap(cur, nxt);
    while (count_ls > 0) {  // further rounds: cubic-first probe, zoom steps, re-rollouts, deeper halvings
      const int warps = count_spec > 0 ? spec_warps : 1;
      rollout(cur, count_ls, nullptr, warps);
      if (count_deriv > 0) expand(cur, count_ls, nullptr, TF_WANT_DERIV, true, -1, false);
      lsupdate(cur, count_ls, nullptr, warps);
      compact(cur, count_ls, nullptr, LS_MASK, nxt, PC_LS, TF_WANT_DERIV, PC_DERIV, TF_SPECULATE, PC_SPEC);
      readback();
      if (err != cudaSuccess) return (int)err;
      count_ls = H->h_counters[PC_LS];
      count_deriv = H->h_counters[PC_DERIV];
      count_spec = H->h_counters[PC_SPEC];
      std::swap(cur, nxt);
    }
    if (backtracking) expand(list_iter, count_iter, nullptr, TF_REFRESH_DYN, true, -2, false);  // solver.cpp:256-262
    timed(PH_CRITERIA, (double)count_iter * 32 * (N + 1), [&] {
      k_phase_costate<Model, CON><<<dim3(g128(count_iter * 32), N + 1), 128, 0, st>>>(P, list_iter, count_iter);
      k_phase_residual<Model, CON><<<dim3(g128(count_iter * 32), N + 1), 128, 0, st>>>(P, list_iter, count_iter);
      k_phase_decide<CON><<<g128(count_iter * 32), 128, 0, st>>>(P, list_iter, count_iter);
    });
    H->launches[PH_CRITERIA] += 2;
    if (CON) expand(list_iter, count_iter, nullptr, TF_REFRESH_GRAD, false, -1, true);  // solver.cpp:475-486
    compact(list_iter, count_iter, nullptr, TF_ACTIVE, list_iter_next, PC_ITER, 0, 0, 0, 0);
    readback();
    if (err != cudaSuccess) return (int)err;
    count_iter = H->h_counters[PC_ITER];
    std::swap(list_iter, list_iter_next);
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
