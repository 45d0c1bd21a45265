// solve_inst.cu -- explicit instantiations of the solve kernel, one group per translation unit
// (compiled with -DALTRO_INST=<group>) so altro_b200/build.py can compile them in parallel.
#include "launchers.h"
#include "solver_kernels.cuh"

namespace altro_b200 {

template <class Model>
static void launch_solve(const DeviceProblem& P, int has_con, cudaStream_t st) {
  const int threads = 32;  // one warp per CTA: 32 consecutive problems
  const int blocks = (P.B + threads - 1) / threads;
  if (has_con)
    solve_kernel<Model, true><<<blocks, threads, 0, st>>>(P);
  else
    solve_kernel<Model, false><<<blocks, threads, 0, st>>>(P);
}

#define ALTRO_DEFINE_LAUNCHER(name, ...)                                          \
  void name(const DeviceProblem& P, int has_constraints, cudaStream_t stream) {   \
    launch_solve<__VA_ARGS__>(P, has_constraints, stream);                        \
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
