// launchers.h -- one host launcher per compiled-in (model, n, m); defined in solve_inst.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "device_problem.h"

namespace altro_b200 {

// Host-side state of the phase-kernel pipeline (solver_phases.cuh).  One PhaseHost per sub-batch.
// Phases: kernels timed with CUDA events in profile mode (INIT, EXPAND = prologue expansion,
// BACKWARD, FORWARD) followed by the sub-phases of k_phase_forward, whose `ms` is the forward
// kernel's time split by the in-kernel %globaltimer shares (rollout passes, expansions, d(phi)
// scan + line-search machines, criteria + AL update).
enum Phase {
  PH_INIT = 0, PH_EXPAND, PH_BACKWARD, PH_FORWARD, PH_FWD_ROLLOUT, PH_FWD_EXPAND, PH_FWD_DPHI_LS,
  PH_FWD_CRITERIA, PH_COUNT
};

enum HostOp {
  OP_SOLVE_PROLOGUE = 0,   // Solve() up to the loop (solver.cpp:417-430) for the sub-batch [g0, g0+G)
  OP_OPEN_LOOP_ROLLOUT = 1,
  OP_CALC_COST = 2,
  OP_UNPACK_JAC = 3,
  OP_SOLVE_ITERATION = 4,  // one iLQR iteration: k_phase_backward + k_phase_forward (iter = PhaseHost::iter)
  OP_KNOT_VIEW = 5,        // derived KnotPointData member PhaseHost::view into cost_out
  OP_KNOT_EVAL = 6,        // expansions of every knot at the working trajectory (k_phase_expand)
};

// derived KnotPointData views served by k_knot_view (solver_phases.cuh)
enum KnotView { KV_CONSTRAINT_VAL = 0, KV_Z_PROJ = 1, KV_LXX = 2, KV_LUU = 3, KV_LUX = 4, KV_RHO = 5 };

struct PhaseHost {
  int op;              // HostOp
  int iter;            // OP_SOLVE_ITERATION: iteration number (0 = first)
  int view, view_rows; // OP_KNOT_VIEW: KnotView id and rows per knot of the output stream
  double* cost_out;    // OP_CALC_COST: device array [Bp]; OP_UNPACK_JAC: dense [A B] stream
                       // [group][knot][n*n + n*m][32]
  int* d_done;         // device: problems of the sub-batch that have stopped (zeroed by the prologue)
  int* h_done;         // pinned [kDoneRing]: copies of *d_done taken after each iteration
  unsigned long long* d_prof;  // device [16]: sub-phase clocks of k_phase_forward (profile mode)
  bool profile;        // record CUDA events around every launch
  double ms[PH_COUNT];         // accumulated kernel time per phase (profile mode)
  long launches[PH_COUNT];     // launches per phase
  double units[PH_COUNT];      // trajectories (or trajectory-knots for PH_EXPAND) processed
  long syncs;                  // host waits on the device (lagged stop-counter checks)
  int fwd_warps;               // warps per CTA of k_phase_forward (1 + speculative candidates)
  int fwd_depth;               // cap of the forward kernel's staging depth (knots in flight, <= 8)
  int backward_team;           // 1: Riccati sweep by the warps of a CTA (solver_team.cuh), 0: one warp
  cudaEvent_t ev0, ev1;
  static constexpr int kDoneRing = 4;
  cudaEvent_t ev_done[kDoneRing];  // recorded after the h_done copy of iteration i % kDoneRing
  // device limits that size the shared-memory staging rings of the sequential sweeps
  int num_sms;
  size_t smem_per_sm, smem_per_cta;
};

// has_constraints selects the AL-enabled instantiation.  host == nullptr: the single persistent
// kernel (solver_kernels.cuh); otherwise the operation host->op.  Returns 0 or a cudaError_t.
typedef int (*solve_launcher)(const DeviceProblem&, int has_constraints, cudaStream_t, PhaseHost* host);

#define ALTRO_DECLARE_LAUNCHER(name) \
  int name(const DeviceProblem& P, int has_constraints, cudaStream_t stream, PhaseHost* host)

ALTRO_DECLARE_LAUNCHER(launch_solve_linear_4_2);
ALTRO_DECLARE_LAUNCHER(launch_solve_linear_2_1);
ALTRO_DECLARE_LAUNCHER(launch_solve_linear_6_3);
ALTRO_DECLARE_LAUNCHER(launch_solve_di_1);
ALTRO_DECLARE_LAUNCHER(launch_solve_di_2);
ALTRO_DECLARE_LAUNCHER(launch_solve_di_3);
ALTRO_DECLARE_LAUNCHER(launch_solve_pendulum);
ALTRO_DECLARE_LAUNCHER(launch_solve_bicycle4);
ALTRO_DECLARE_LAUNCHER(launch_solve_bicycle5);
ALTRO_DECLARE_LAUNCHER(launch_solve_chain_4_2);
ALTRO_DECLARE_LAUNCHER(launch_solve_chain_4_4);
ALTRO_DECLARE_LAUNCHER(launch_solve_chain_6_2);
ALTRO_DECLARE_LAUNCHER(launch_solve_chain_6_4);
ALTRO_DECLARE_LAUNCHER(launch_solve_chain_12_2);
ALTRO_DECLARE_LAUNCHER(launch_solve_chain_12_4);

}  // namespace altro_b200
