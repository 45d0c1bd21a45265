// ref_shim.cpp -- ORACLE support (test infrastructure only).
// C-ABI window onto the REFERENCE's own compiled line search (src/linesearch/linesearch.cpp,
// src/linesearch/cubicspline.c, built by path from /root/reference into oracle/_ref/ -- see
// Makefile target `ref`).  Used only by tests/test_oracle_linesearch.py to cross-check
// oracle/linesearch_port.c.  No reference source is copied here.
#include <functional>

#include "linesearch/linesearch.hpp"
extern "C" {
#include "linesearch/cubicspline.h"
}

extern "C" {

typedef void (*merit_cb)(void* ctx, double alpha, double* phi, double* dphi);

// Runs CubicLineSearch::Run (linesearch.cpp:37) and reports alpha, status, iteration count.
double ref_linesearch_run(merit_cb f, void* ctx, double alpha0, double phi0, double dphi0,
                          double c1, double c2, int try_cubic_first, int use_backtracking,
                          int* status, int* iters, double* phi_out, double* dphi_out,
                          int* sufficient_decrease, int* curvature) {
  linesearch::CubicLineSearch ls;
  ls.SetOptimalityTolerances(c1, c2);
  ls.try_cubic_first = try_cubic_first != 0;
  ls.use_backtracking_linesearch = use_backtracking != 0;
  auto merit = [f, ctx](double a, double* phi, double* dphi) { f(ctx, a, phi, dphi); };
  double alpha = ls.Run(merit, alpha0, phi0, dphi0);
  *status = static_cast<int>(ls.GetStatus());
  *iters = ls.Iterations();
  ls.GetFinalMeritValues(phi_out, dphi_out);
  *sufficient_decrease = ls.SufficientDecreaseSatisfied();
  *curvature = ls.CurvatureConditionSatisfied();
  return alpha;
}

// CubicSpline_From2Points + CubicSpline_ArgMin (cubicspline.c:18, :111)
double ref_cubic_argmin(double x1, double y1, double d1, double x2, double y2, double d2,
                        int* err_build, int* err_argmin, double* coeffs) {
  enum CubicSplineReturnCodes e1, e2;
  CubicSpline p = CubicSpline_From2Points(x1, y1, d1, x2, y2, d2, &e1);
  *err_build = static_cast<int>(e1);
  coeffs[0] = p.x0; coeffs[1] = p.a; coeffs[2] = p.b; coeffs[3] = p.c; coeffs[4] = p.d;
  if (e1 != CS_NOERROR) { *err_argmin = -1; return 0.0; }
  double x = CubicSpline_ArgMin(&p, &e2);
  *err_argmin = static_cast<int>(e2);
  return x;
}

}  // extern "C"
