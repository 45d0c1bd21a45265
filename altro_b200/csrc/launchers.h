// launchers.h -- one host launcher per compiled-in (model, n, m); defined in solve_inst.cu.
#pragma once
#include <cuda_runtime.h>

#include "device_problem.h"

namespace altro_b200 {

typedef void (*solve_launcher)(const DeviceProblem&, int has_constraints, cudaStream_t);

#define ALTRO_DECLARE_LAUNCHER(name) \
  void name(const DeviceProblem& P, int has_constraints, cudaStream_t stream)

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
