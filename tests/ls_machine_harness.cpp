// Host harness for altro_b200/csrc/linesearch.cuh (the device line-search state machine),
// compiled with g++ by tests/test_linesearch_machine.py.  Drives the machine to completion
// with a C callback so it can be compared with the oracle port / the reference build.
#include "../altro_b200/csrc/linesearch.cuh"

extern "C" {
typedef void (*merit_cb)(void* ctx, double alpha, double* phi, double* dphi);

double lsm_run(merit_cb f, void* ctx, double alpha0, double phi0, double dphi0, double c1,
               double c2, int try_cubic_first, int use_backtracking, int* status, int* iters,
               double* phi_out, double* dphi_out, int* sd, int* cv) {
  altro_b200::LsOptions o;
  o.c1 = c1;
  o.c2 = c2;
  o.try_cubic_first = try_cubic_first != 0;
  o.use_backtracking = use_backtracking != 0;
  altro_b200::LsMachine m;
  m.start(o, alpha0, phi0, dphi0);
  while (!m.done()) {
    double phi = 0, dphi = 0;
    f(ctx, m.alpha, &phi, m.want_derivative() ? &dphi : nullptr);
    m.update(o, phi, dphi);
  }
  *status = m.status;
  *iters = m.n_iters;
  *phi_out = m.phi;
  *dphi_out = m.dphi;
  *sd = m.sufficient_decrease;
  *cv = m.curvature;
  return m.alpha;
}
}
