/*
 * linesearch_port.c -- ORACLE (test infrastructure only, see altro_oracle.h).
 *
 * Plain-C restatement of the reference's strong-Wolfe cubic line search:
 *   src/linesearch/cubicspline.c:18-42   (CubicSpline_From2Points)
 *   src/linesearch/cubicspline.c:111-181 (CubicSpline_ArgMin)
 *   src/linesearch/cubicspline.c:229-246 (QuadraticFormula)
 *   src/linesearch/linesearch.cpp:37-217 (CubicLineSearch::Run)
 *   src/linesearch/linesearch.cpp:233-351 (Zoom)
 *   src/linesearch/linesearch.cpp:385-412 (SimpleBacktracking)
 * Pinned by tests/test_oracle_linesearch.py against the known answers in
 * src/linesearch/test/linesearch_tests.cpp and, when oracle/_ref/ is built, against the
 * reference's own compiled sources on randomised merit functions.
 */
#include <math.h>
#include <stddef.h>

#include "altro_oracle.h"

#define LS_TOL 1e-6 /* cubicspline.c:10 LINESEARCH_TOL */

/* CubicSplineReturnCodes, cubicspline.h:8-19 */
enum {
  CS_NOERROR,
  CS_FOUND_MINIMUM,
  CS_INVALIDPOINTER,
  CS_SADDLEPOINT,
  CS_NOMINIMUM,
  CS_IS_POSITIVE_QUADRATIC,
  CS_IS_LINEAR,
  CS_IS_CONSTANT,
  CS_UNEXPECTED_ERROR,
  CS_SAME_POINT,
};

/* CubicLineSearch::ReturnCodes, linesearch.hpp:16-25 */
enum {
  LS_NOERROR,
  LS_MINIMUM_FOUND,
  LS_INVALID_POINTER,
  LS_NOT_DESCENT_DIRECTION,
  LS_WINDOW_TOO_SMALL,
  LS_GOT_NONFINITE_STEP_SIZE,
  LS_MAX_ITERATIONS,
  LS_HIT_MAX_STEPSIZE,
};

/* cubicspline.c:18-42 */
int oracle_spline_from2points(oracle_spline *p, double x1, double y1, double d1, double x2,
                              double y2, double d2) {
  double delta = x2 - x1;
  if (fabs(delta) < LS_TOL) {
    p->x0 = p->a = p->b = p->c = p->d = NAN;
    return CS_SAME_POINT;
  }
  p->x0 = x1;
  p->a = y1;
  p->b = d1;
  p->c = 3 * (y2 - y1) / (delta * delta) - (d2 + 2 * d1) / delta;
  p->d = (d2 + d1) / (delta * delta) - 2 * (y2 - y1) / (delta * delta * delta);
  return CS_NOERROR;
}

/* cubicspline.c:229-246 */
static int quadratic_formula(double *x1, double *x2, double a, double b, double c) {
  if (fabs(a) < LS_TOL) return CS_IS_LINEAR;
  double s2 = b * b - 4 * a * c;
  double s;
  if (fabs(s2) < LS_TOL) {
    s = 0.0;
  } else if (s2 < 0) {
    return CS_NOMINIMUM;
  } else {
    s = sqrt(s2);
  }
  *x1 = (-b + s) / (2 * a);
  *x2 = (-b - s) / (2 * a);
  return CS_NOERROR;
}

/* cubicspline.c:111-181 */
double oracle_spline_argmin(const oracle_spline *p, int *err) {
  double b = p->b, c = p->c, d = p->d;
  int is_quadratic = fabs(d) < LS_TOL;
  int is_linear = is_quadratic && (fabs(c) < LS_TOL);
  int is_constant = is_linear && (fabs(b) < LS_TOL);
  if (is_quadratic) {
    if (is_linear) {
      *err = is_constant ? CS_IS_CONSTANT : CS_IS_LINEAR;
      return NAN;
    } else if (c <= 0) {
      *err = CS_IS_POSITIVE_QUADRATIC;
      return NAN;
    } else {
      *err = CS_FOUND_MINIMUM;
      return -b / (2 * c) + p->x0;
    }
  }
  double d1, d2;
  int e = quadratic_formula(&d1, &d2, 3 * d, 2 * c, b);
  if (e != CS_NOERROR) {
    *err = e;
    return NAN;
  }
  double curv1 = 2 * c + 6 * d * d1;
  double curv2 = 2 * c + 6 * d * d2;
  double x1 = d1 + p->x0;
  double x2 = d2 + p->x0;
  if (fabs(curv1) < LS_TOL && fabs(curv2) < LS_TOL) {
    *err = CS_SADDLEPOINT;
    return NAN;
  } else if (curv1 > 0 && curv2 < 0) {
    *err = CS_FOUND_MINIMUM;
    return x1;
  } else if (curv1 < 0 && curv2 > 0) {
    *err = CS_FOUND_MINIMUM;
    return x2;
  }
  *err = CS_UNEXPECTED_ERROR;
  return NAN;
}

/* linesearch.hpp:41-56 defaults */
void oracle_ls_init(oracle_linesearch *ls) {
  ls->max_iters = 25;
  ls->alpha_max = 2.0;
  ls->beta_increase = 1.5;
  ls->beta_decrease = 0.5;
  ls->min_interval_size = 1e-6;
  ls->try_cubic_first = 0;
  ls->use_backtracking_linesearch = 0;
  ls->c1 = 1e-4;
  ls->c2 = 0.9;
  ls->return_code = LS_NOERROR;
  ls->n_iters = 0;
  ls->phi0 = ls->phi = ls->phi_lo = ls->phi_hi = 0;
  ls->dphi0 = ls->dphi = ls->dphi_lo = ls->dphi_hi = 0;
  ls->sufficient_decrease = ls->curvature = 0;
}

/* linesearch.cpp:233-351 */
static double ls_zoom(oracle_linesearch *ls, oracle_merit_fn f, void *ctx, double alo,
                      double ahi) {
  double alpha = alo;
  if (!isfinite(alo) || !isfinite(ahi)) {
    ls->return_code = LS_GOT_NONFINITE_STEP_SIZE;
    return 0;
  }
  double c1 = ls->c1, c2 = ls->c2, phi0 = ls->phi0, dphi0 = ls->dphi0;
  double phi_lo = ls->phi_lo, phi_hi = ls->phi_hi, dphi_lo = ls->dphi_lo, dphi_hi = ls->dphi_hi;

  for (int zoom_iter = ls->n_iters + 1; zoom_iter < ls->max_iters; ++zoom_iter) {
    if (fabs(alo - ahi) < ls->min_interval_size) {
      alpha = (alo + ahi) / 2.0;
      ls->n_iters += 1;
      f(ctx, alpha, &ls->phi, &ls->dphi);
      ls->sufficient_decrease = ls->phi <= phi0 + c1 * alpha * dphi0;
      ls->curvature = fabs(ls->dphi) <= -c2 * dphi0;
      ls->return_code =
          (ls->sufficient_decrease && ls->curvature) ? LS_MINIMUM_FOUND : LS_WINDOW_TOO_SMALL;
      return alpha;
    }
    oracle_spline p;
    int cs_err = oracle_spline_from2points(&p, alo, phi_lo, dphi_lo, ahi, phi_hi, dphi_hi);
    int cubic_spline_failed = 1;
    if (cs_err == CS_NOERROR) {
      alpha = oracle_spline_argmin(&p, &cs_err);
      if (cs_err == CS_FOUND_MINIMUM && isfinite(alpha)) cubic_spline_failed = 0;
    }
    if (cubic_spline_failed) alpha = (alo + ahi) / 2;

    ls->n_iters += 1;
    f(ctx, alpha, &ls->phi, &ls->dphi);
    double phi = ls->phi, dphi = ls->dphi;
    int sufficient_decrease = phi <= phi0 + c1 * alpha * dphi0;
    int higher_than_lo = phi > phi_lo;
    int curvature = fabs(dphi) <= -c2 * dphi0;
    if (sufficient_decrease && curvature) {
      ls->sufficient_decrease = 1;
      ls->curvature = 1;
      ls->return_code = LS_MINIMUM_FOUND;
      return alpha;
    }
    if (!sufficient_decrease || higher_than_lo) {
      ahi = alpha;
      phi_hi = phi;
      dphi_hi = dphi;
    } else {
      int reset_ahi = dphi * (ahi - alo) <= 0;
      if (reset_ahi) {
        ahi = alo;
        phi_hi = phi_lo;
        dphi_hi = dphi_lo;
      }
      alo = alpha;
      phi_lo = phi;
      dphi_lo = dphi;
    }
  }
  ls->return_code = LS_MAX_ITERATIONS;
  return alpha;
}

/* linesearch.cpp:385-412 */
static double ls_simple_backtracking(oracle_linesearch *ls, oracle_merit_fn f, void *ctx,
                                     double alpha0) {
  double alpha = alpha0;
  double c1 = ls->c1, phi0 = ls->phi0, dphi0 = ls->dphi0;
  for (int iter = 1; iter < ls->max_iters; ++iter) {
    ls->n_iters += 1;
    f(ctx, alpha, &ls->phi, NULL);
    int sufficient_decrease_satisfied = ls->phi <= phi0 + c1 * alpha * dphi0;
    if (sufficient_decrease_satisfied) {
      ls->sufficient_decrease = 1;
      ls->curvature = 1;
      ls->return_code = LS_MINIMUM_FOUND;
      return alpha;
    } else {
      alpha *= ls->beta_decrease;
    }
  }
  return alpha;
}

/* linesearch.cpp:37-217 */
double oracle_ls_run(oracle_linesearch *ls, oracle_merit_fn f, void *ctx, double alpha0,
                     double phi0, double dphi0) {
  ls->phi0 = phi0;
  ls->dphi0 = dphi0;
  ls->n_iters = 0;
  ls->sufficient_decrease = 0;
  ls->curvature = 0;
  ls->return_code = LS_NOERROR;
  if (dphi0 >= 0.0) {
    ls->return_code = LS_NOT_DESCENT_DIRECTION;
    return 0.0;
  }
  double alpha_prev = 0.0, phi_prev = phi0, dphi_prev = dphi0;
  double alpha = alpha0;
  double c1 = ls->c1, c2 = ls->c2;
  int hit_max_alpha = 0;

  for (int iter = 0; iter < ls->max_iters; ++iter) {
    ls->n_iters += 1;
    f(ctx, alpha, &ls->phi, &ls->dphi);
    double phi = ls->phi, dphi = ls->dphi;
    int sufficient_decrease_satisfied = phi <= phi0 + c1 * alpha * dphi0;
    int function_not_decreasing = phi >= phi_prev;
    int strong_wolfe_satisfied = fabs(dphi) <= -c2 * dphi0;

    if (sufficient_decrease_satisfied && strong_wolfe_satisfied) {
      ls->sufficient_decrease = 1;
      ls->curvature = 1;
      ls->return_code = LS_MINIMUM_FOUND;
      return alpha;
    } else if (iter == 0 && ls->try_cubic_first) {
      oracle_spline p;
      int cs_err = oracle_spline_from2points(&p, 0, phi0, dphi0, alpha, phi, dphi);
      int cubic_spline_failed = 1;
      double alpha_cubic = NAN;
      if (cs_err == CS_NOERROR) {
        alpha_cubic = oracle_spline_argmin(&p, &cs_err);
        if (cs_err == CS_FOUND_MINIMUM && isfinite(alpha_cubic)) cubic_spline_failed = 0;
      }
      if (!cubic_spline_failed) {
        ls->n_iters += 1;
        double phi_cubic, dphi_cubic;
        ++iter;
        f(ctx, alpha_cubic, &phi_cubic, &dphi_cubic);
        int sd_cubic = phi_cubic <= phi0 + c1 * alpha_cubic * dphi0;
        int sw_cubic = fabs(dphi_cubic) <= -c2 * dphi0;
        if (sd_cubic && sw_cubic) {
          ls->phi = phi_cubic;
          ls->dphi = dphi_cubic;
          ls->sufficient_decrease = 1;
          ls->curvature = 1;
          ls->return_code = LS_MINIMUM_FOUND;
          return alpha_cubic;
        }
      }
    }

    if (ls->use_backtracking_linesearch) {
      return ls_simple_backtracking(ls, f, ctx, alpha0 * ls->beta_decrease);
    }

    if (!sufficient_decrease_satisfied || (iter > 0 && function_not_decreasing)) {
      double alo = alpha_prev;
      ls->phi_lo = phi_prev;
      ls->dphi_lo = dphi_prev;
      double ahi = alpha;
      ls->phi_hi = phi;
      ls->dphi_hi = dphi;
      return ls_zoom(ls, f, ctx, alo, ahi);
    }

    if (dphi >= 0) {
      double alo = alpha;
      double ahi = alpha_prev;
      ls->phi_lo = phi;
      ls->dphi_lo = dphi;
      ls->phi_hi = phi_prev;
      ls->dphi_hi = dphi_prev;
      return ls_zoom(ls, f, ctx, alo, ahi);
    }

    alpha_prev = alpha;
    alpha = alpha * ls->beta_increase;
    if (alpha > ls->alpha_max) {
      alpha = ls->alpha_max;
      if (hit_max_alpha) {
        ls->return_code = LS_HIT_MAX_STEPSIZE;
        ls->sufficient_decrease = sufficient_decrease_satisfied;
        ls->curvature = strong_wolfe_satisfied;
        return alpha;
      } else {
        hit_max_alpha = 1;
      }
    }
    phi_prev = phi;
    dphi_prev = dphi;
  }
  return alpha;
}
