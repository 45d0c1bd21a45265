// linesearch.cuh -- per-trajectory strong-Wolfe line search as a resumable state machine.
//
// Behaviour follows the reference's CubicLineSearch (src/linesearch/linesearch.cpp:37-412,
// src/linesearch/cubicspline.c:18-246) decision for decision, but is re-shaped for SIMT: the
// reference calls the merit function from five different places (Run x2, Zoom x2,
// SimpleBacktracking); here the search is a state machine that *asks* for the next step length
// (`ls_next`) and is *told* the result (`ls_update`), so the expensive rollout has a single
// call site and all lanes of a warp that still need an evaluation execute it together no
// matter which branch of the search each of them is in.
//
// Compiles for host and device (ALTRO_HD) so that tests/test_linesearch_machine.py can check
// it on the CPU against the oracle port and the reference's own compiled sources.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define ALTRO_HD __host__ __device__ __forceinline__
#else
#define ALTRO_HD inline
#endif

namespace altro_b200 {

// CubicLineSearch::ReturnCodes (linesearch.hpp:16-25), same values
enum LsStatus {
  LS_NOERROR = 0,
  LS_MINIMUM_FOUND = 1,
  LS_INVALID_POINTER = 2,
  LS_NOT_DESCENT_DIRECTION = 3,
  LS_WINDOW_TOO_SMALL = 4,
  LS_GOT_NONFINITE_STEP_SIZE = 5,
  LS_MAX_ITERATIONS = 6,
  LS_HIT_MAX_STEPSIZE = 7,
};

constexpr double kLsTol = 1e-6;  // cubicspline.c:10

// Minimiser of the Hermite cubic through (x1,y1,d1),(x2,y2,d2); returns false when the
// reference would report anything but CS_FOUND_MINIMUM with a finite argument.
// cubicspline.c:18-42 (From2Points), :111-181 (ArgMin), :229-246 (QuadraticFormula)
ALTRO_HD bool cubic_argmin(double x1, double y1, double d1, double x2, double y2, double d2,
                           double* xmin) {
  double delta = x2 - x1;
  if (fabs(delta) < kLsTol) return false;  // CS_SAME_POINT
  double b = d1;
  double c = 3 * (y2 - y1) / (delta * delta) - (d2 + 2 * d1) / delta;
  double d = (d2 + d1) / (delta * delta) - 2 * (y2 - y1) / (delta * delta * delta);
  double x;
  if (fabs(d) < kLsTol) {                    // quadratic
    if (fabs(c) < kLsTol) return false;      // linear / constant
    if (c <= 0) return false;                // CS_IS_POSITIVE_QUADRATIC
    x = -b / (2 * c) + x1;
  } else {
    double qa = 3 * d, qb = 2 * c, qc = b;
    if (fabs(qa) < kLsTol) return false;     // CS_IS_LINEAR
    double s2 = qb * qb - 4 * qa * qc;
    double s;
    if (fabs(s2) < kLsTol) {
      s = 0.0;
    } else if (s2 < 0) {
      return false;                          // CS_NOMINIMUM
    } else {
      s = sqrt(s2);
    }
    double r1 = (-qb + s) / (2 * qa);
    double r2 = (-qb - s) / (2 * qa);
    double curv1 = 2 * c + 6 * d * r1;
    double curv2 = 2 * c + 6 * d * r2;
    if (fabs(curv1) < kLsTol && fabs(curv2) < kLsTol) return false;  // CS_SADDLEPOINT
    if (curv1 > 0 && curv2 < 0) {
      x = r1 + x1;
    } else if (curv1 < 0 && curv2 > 0) {
      x = r2 + x1;
    } else {
      return false;                          // CS_UNEXPECTED_ERROR
    }
  }
  if (!isfinite(x)) return false;
  *xmin = x;
  return true;
}

struct LsOptions {  // linesearch.hpp:41-56
  int max_iters = 25;
  double alpha_max = 2.0;
  double beta_increase = 1.5;
  double beta_decrease = 0.5;
  double min_interval_size = 1e-6;
  bool try_cubic_first = false;
  bool use_backtracking = false;
  double c1 = 1e-4;
  double c2 = 0.9;
};

struct LsMachine {
  enum Phase { P_BRACKET, P_CUBIC_FIRST, P_BACKTRACK, P_ZOOM, P_ZOOM_WINDOW, P_DONE };
  int phase;
  int status;   // LsStatus
  int n_iters;  // merit evaluations so far (CubicLineSearch::Iterations)
  int iter;     // loop counter of Run / SimpleBacktracking / Zoom
  bool hit_max_alpha;
  bool sufficient_decrease, curvature;
  double alpha0, phi0, dphi0;
  double phi, dphi;  // phi_/dphi_ (GetFinalMeritValues)
  double alpha;      // step being evaluated / returned
  double alpha_prev, phi_prev, dphi_prev;
  // values of the first bracket evaluation, needed after a rejected cubic-first probe
  double alpha_b, phi_b, dphi_b;
  bool sd_b, sw_b;
  // zoom interval
  double alo, ahi, phi_lo, phi_hi, dphi_lo, dphi_hi;

  // linesearch.cpp:37-60.  Returns false if no evaluation is needed (not a descent direction).
  ALTRO_HD bool start(const LsOptions& o, double alpha0_, double phi0_, double dphi0_) {
    (void)o;
    alpha0 = alpha0_;
    phi0 = phi0_;
    dphi0 = dphi0_;
    n_iters = 0;
    sufficient_decrease = false;
    curvature = false;
    status = LS_NOERROR;
    hit_max_alpha = false;
    iter = 0;
    phi = 0;
    dphi = 0;
    if (dphi0 >= 0.0) {
      status = LS_NOT_DESCENT_DIRECTION;
      alpha = 0.0;
      phase = P_DONE;
      return false;
    }
    alpha_prev = 0.0;
    phi_prev = phi0;
    dphi_prev = dphi0;
    alpha = alpha0;
    phase = P_BRACKET;
    return true;
  }

  ALTRO_HD bool done() const { return phase == P_DONE; }
  // does the evaluation at `alpha` need the derivative?  (linesearch.cpp:395 passes nullptr)
  ALTRO_HD bool want_derivative() const { return phase != P_BACKTRACK; }

  ALTRO_HD void finish(int st, double a) {
    status = st;
    alpha = a;
    phase = P_DONE;
  }

  // Prepare the next zoom probe (linesearch.cpp:256-296) or finish.
  ALTRO_HD void zoom_prepare(const LsOptions& o) {
    if (!(iter < o.max_iters)) {  // loop exhausted, :349-350
      finish(LS_MAX_ITERATIONS, alpha);
      return;
    }
    if (fabs(alo - ahi) < o.min_interval_size) {  // :258-274
      alpha = (alo + ahi) / 2.0;
      phase = P_ZOOM_WINDOW;
      return;
    }
    double a;
    if (cubic_argmin(alo, phi_lo, dphi_lo, ahi, phi_hi, dphi_hi, &a)) {
      alpha = a;
    } else {
      alpha = (alo + ahi) / 2;
    }
    phase = P_ZOOM;
  }

  // Enter Zoom(alo, ahi) (linesearch.cpp:233-254)
  ALTRO_HD void zoom_enter(const LsOptions& o, double alo_, double ahi_) {
    alo = alo_;
    ahi = ahi_;
    alpha = alo_;
    if (!isfinite(alo) || !isfinite(ahi)) {
      finish(LS_GOT_NONFINITE_STEP_SIZE, 0.0);
      return;
    }
    iter = n_iters + 1;
    zoom_prepare(o);
  }

  // Continuation of Run's loop body after the acceptance test failed (:130-213)
  ALTRO_HD void bracket_continue(const LsOptions& o, double a, double ph, double dph, bool sd,
                                 bool sw) {
    if (o.use_backtracking) {  // :130-132, SimpleBacktracking(alpha0 * beta_decrease)
      alpha = alpha0 * o.beta_decrease;
      iter = 1;
      if (iter < o.max_iters) {
        phase = P_BACKTRACK;
      } else {
        finish(status, alpha);
      }
      return;
    }
    bool function_not_decreasing = ph >= phi_prev;
    if (!sd || (iter > 0 && function_not_decreasing)) {  // :144-162
      phi_lo = phi_prev;
      dphi_lo = dphi_prev;
      phi_hi = ph;
      dphi_hi = dph;
      zoom_enter(o, alpha_prev, a);
      return;
    }
    if (dph >= 0) {  // :171-187
      phi_lo = ph;
      dphi_lo = dph;
      phi_hi = phi_prev;
      dphi_hi = dphi_prev;
      zoom_enter(o, a, alpha_prev);
      return;
    }
    // expand the interval, :189-211
    alpha_prev = a;
    double an = a * o.beta_increase;
    if (an > o.alpha_max) {
      an = o.alpha_max;
      if (hit_max_alpha) {
        sufficient_decrease = sd;
        curvature = sw;
        finish(LS_HIT_MAX_STEPSIZE, an);
        return;
      }
      hit_max_alpha = true;
    }
    phi_prev = ph;
    dphi_prev = dph;
    alpha = an;
    iter += 1;
    if (iter < o.max_iters) {
      phase = P_BRACKET;
    } else {
      finish(status, alpha);  // loop exhausted: `return alpha`, :216
    }
  }

  // Feed the merit value (and derivative, if want_derivative()) at `alpha`.
  ALTRO_HD void update(const LsOptions& o, double ph, double dph) {
    const double c1 = o.c1, c2 = o.c2;
    n_iters += 1;
    switch (phase) {
      case P_BRACKET: {  // :74-128
        phi = ph;
        dphi = dph;
        bool sd = ph <= phi0 + c1 * alpha * dphi0;
        bool sw = fabs(dph) <= -c2 * dphi0;
        if (sd && sw) {
          sufficient_decrease = true;
          curvature = true;
          finish(LS_MINIMUM_FOUND, alpha);
          return;
        }
        if (iter == 0 && o.try_cubic_first) {
          double ac;
          if (cubic_argmin(0, phi0, dphi0, alpha, ph, dph, &ac)) {
            alpha_b = alpha;
            phi_b = ph;
            dphi_b = dph;
            sd_b = sd;
            sw_b = sw;
            iter += 1;  // ++iter, :105
            alpha = ac;
            phase = P_CUBIC_FIRST;
            return;
          }
        }
        bracket_continue(o, alpha, ph, dph, sd, sw);
        return;
      }
      case P_CUBIC_FIRST: {  // :106-126; phi_/dphi_ keep the bracket values unless accepted
        bool sd = ph <= phi0 + c1 * alpha * dphi0;
        bool sw = fabs(dph) <= -c2 * dphi0;
        if (sd && sw) {
          phi = ph;
          dphi = dph;
          sufficient_decrease = true;
          curvature = true;
          finish(LS_MINIMUM_FOUND, alpha);
          return;
        }
        bracket_continue(o, alpha_b, phi_b, dphi_b, sd_b, sw_b);
        return;
      }
      case P_BACKTRACK: {  // :385-412
        phi = ph;
        bool sd = ph <= phi0 + c1 * alpha * dphi0;
        if (sd) {
          sufficient_decrease = true;
          curvature = true;
          finish(LS_MINIMUM_FOUND, alpha);
          return;
        }
        alpha *= o.beta_decrease;
        iter += 1;
        if (!(iter < o.max_iters)) finish(status, alpha);
        return;
      }
      case P_ZOOM_WINDOW: {  // :258-274
        phi = ph;
        dphi = dph;
        sufficient_decrease = ph <= phi0 + c1 * alpha * dphi0;
        curvature = fabs(dph) <= -c2 * dphi0;
        finish((sufficient_decrease && curvature) ? LS_MINIMUM_FOUND : LS_WINDOW_TOO_SMALL, alpha);
        return;
      }
      case P_ZOOM: {  // :298-347
        phi = ph;
        dphi = dph;
        bool sd = ph <= phi0 + c1 * alpha * dphi0;
        bool higher_than_lo = ph > phi_lo;
        bool cv = fabs(dph) <= -c2 * dphi0;
        if (sd && cv) {
          sufficient_decrease = true;
          curvature = true;
          finish(LS_MINIMUM_FOUND, alpha);
          return;
        }
        if (!sd || higher_than_lo) {
          ahi = alpha;
          phi_hi = ph;
          dphi_hi = dph;
        } else {
          bool reset_ahi = dph * (ahi - alo) <= 0;
          if (reset_ahi) {
            ahi = alo;
            phi_hi = phi_lo;
            dphi_hi = dphi_lo;
          }
          alo = alpha;
          phi_lo = ph;
          dphi_lo = dph;
        }
        iter += 1;
        zoom_prepare(o);
        return;
      }
      default:
        return;
    }
  }
};

}  // namespace altro_b200
