// launchers.h -- one host launcher per compiled-in (model, n, m); defined in solve_inst.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "device_problem.h"

namespace altro_b200 {

// Host-side state of the phase-kernel pipeline (solver_phases.cuh)
enum Phase { PH_INIT = 0, PH_EXPAND, PH_BACKWARD, PH_ROLLOUT, PH_LSUPDATE, PH_CRITERIA, PH_COMPACT, PH_COUNT };

enum HostOp { OP_SOLVE = 0, OP_OPEN_LOOP_ROLLOUT = 1, OP_CALC_COST = 2, OP_UNPACK_JAC = 3 };

struct PhaseHost {
  int op;              // HostOp
  double* cost_out;    // OP_CALC_COST: device array [Bp]; OP_UNPACK_JAC: dense [A B] stream
                       // [group][knot][n*n + n*m][32]
  int* h_counters;     // pinned, 8 ints
  int* list_aux;       // second buffer for list_iter
  bool profile;        // record CUDA events around every launch
  double ms[PH_COUNT];         // accumulated kernel time per phase (profile mode)
  long launches[PH_COUNT];     // launches per phase
  double units[PH_COUNT];      // trajectories (or trajectory-knots for PH_EXPAND) processed
  long syncs;                  // host<->device count readbacks
  cudaEvent_t ev0, ev1;
  // device limits that size the shared-memory staging rings of the sequential sweeps
  int num_sms;
  size_t smem_per_sm, smem_per_cta;
};

// has_constraints selects the AL-enabled instantiation.  host == nullptr: the single persistent
// kernel (solver_kernels.cuh); otherwise the phase pipeline.  Returns 0 or a cudaError_t.
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
