/*
 * models.c -- ORACLE (test infrastructure only, see altro_oracle.h).
 *
 * The dynamics models every configuration uses, restated from the reference's test helpers:
 *   double integrator   test/test_utils.cpp:18-41   (closed-form discrete; `b = h*h/2` in float)
 *   pendulum            test/test_utils.cpp:43-82
 *   explicit midpoint   test/test_utils.cpp:84-132  (`h/2` is a FLOAT division, then promoted)
 *   bicycle (n=4)       test/test_utils.cpp:134-238 (centre-of-gravity frame, L=2.7, lr=1.5)
 * plus the two models that do not exist in the reference and are defined once here and once
 * in altro_b200/csrc/models.cuh:
 *   bicycle5  state [x,y,theta,delta,v], input [a, delta_dot]: bicycle-4 with v a state, v'=a
 *   chain     q in R^{n/2}: q_i'' = -g sin(q_i) - b q_i' + kc (q_{i-1} - 2 q_i + q_{i+1}) + u_i[i<m]
 * Jacobians are column-major n x (n+m) = [A B] (altro_solver.hpp:69-70).
 */
#include <math.h>
#include <string.h>

#include "altro_oracle.h"

#define NMAX 16

static const double kPendulumMass = 1.0;
static const double kPendulumLength = 0.5;
static const double kPendulumFrictionCoeff = 0.1;
static const double kPendulumGravity = 9.81;

static const double kChainGravity = 9.81;
static const double kChainDamping = 0.1;
static const double kChainCoupling = 1.0;

void oracle_model_dims(int model_id, const double *params, int *n, int *m) {
  switch (model_id) {
    case ORACLE_MODEL_DOUBLE_INTEGRATOR: {
      int dim = (int)params[0];
      *n = 2 * dim;
      *m = dim;
      break;
    }
    case ORACLE_MODEL_PENDULUM:
      *n = 2;
      *m = 1;
      break;
    case ORACLE_MODEL_BICYCLE4:
      *n = 4;
      *m = 2;
      break;
    case ORACLE_MODEL_BICYCLE5:
      *n = 5;
      *m = 2;
      break;
    case ORACLE_MODEL_CHAIN:
      *n = (int)params[0];
      *m = (int)params[1];
      break;
    default:
      *n = 0;
      *m = 0;
  }
}

/* test_utils.cpp:43-57 */
static void pendulum_xdot(double *xdot, const double *x, const double *u) {
  double l = kPendulumLength, g = kPendulumGravity, b = kPendulumFrictionCoeff;
  double m = kPendulumMass * l * l;
  double theta = x[0], omega = x[1];
  double omega_dot = u[0] / m - g * sin(theta) / l - b * omega / m;
  xdot[0] = omega;
  xdot[1] = omega_dot;
}

/* test_utils.cpp:59-82 */
static void pendulum_cjac(double *jac, const double *x, const double *u) {
  (void)u;
  double l = kPendulumLength, g = kPendulumGravity, b = kPendulumFrictionCoeff;
  double m = kPendulumMass * l * l;
  jac[0] = 0.0;
  jac[1] = -g * cos(x[0]) / l;
  jac[2] = 1.0;
  jac[3] = -b / m;
  jac[4] = 0.0;
  jac[5] = 1 / m;
}

/* test_utils.cpp:134-167, CenterOfGravity frame; params: L, lr */
static void bicycle4_xdot(const double *prm, double *xdot, const double *x, const double *u) {
  double L = prm[0], lr = prm[1];
  double v = u[0], delta_dot = u[1], theta = x[2], delta = x[3];
  double beta = atan2(lr * delta, L);
  double omega = v * cos(beta) * tan(delta) / L;
  double stheta = sin(theta + beta), ctheta = cos(theta + beta);
  xdot[0] = v * ctheta;
  xdot[1] = v * stheta;
  xdot[2] = omega;
  xdot[3] = delta_dot;
}

/* test_utils.cpp:169-238 */
static void bicycle4_cjac(const double *prm, double *jac, const double *x, const double *u) {
  double L = prm[0], lr = prm[1];
  double v = u[0], theta = x[2], delta = x[3];
  double by = lr * delta, bx = L;
  double beta = atan2(by, bx);
  double dbeta_ddelta = bx / (bx * bx + by * by) * lr;
  double domega_ddelta =
      v / L * (-sin(beta) * tan(delta) * dbeta_ddelta + cos(beta) / (cos(delta) * cos(delta)));
  double domega_dv = cos(beta) * tan(delta) / L;
  double stheta = sin(theta + beta), ctheta = cos(theta + beta);
  double ds_dtheta = +cos(theta + beta), dc_dtheta = -sin(theta + beta);
  double ds_ddelta = +cos(theta + beta) * dbeta_ddelta;
  double dc_ddelta = -sin(theta + beta) * dbeta_ddelta;
  const int n = 4;
  memset(jac, 0, sizeof(double) * 4 * 6);
#define J(i, j) jac[(i) + n * (j)]
  J(0, 2) = v * dc_dtheta;
  J(0, 3) = v * dc_ddelta;
  J(0, 4) = ctheta;
  J(1, 2) = v * ds_dtheta;
  J(1, 3) = v * ds_ddelta;
  J(1, 4) = stheta;
  J(2, 3) = domega_ddelta;
  J(2, 4) = domega_dv;
  J(3, 5) = 1.0;
#undef J
}

/* bicycle5: bicycle-4 equations with the speed as fifth state.  This model does
 * not exist in the reference; it is DEFINED here (and identically in altro_b200/csrc/models.cuh)
 * in the algebraic form that avoids atan2: with beta = atan(lr*delta/L),
 *   cos(beta) = L / hypot,  sin(beta) = lr*delta / hypot,  hypot = sqrt(L^2 + (lr*delta)^2),
 * and sin/cos(theta+beta) by the angle-addition formulas. */
static void bicycle5_trig(const double *prm, const double *x, double *sb, double *cb, double *tand,
                          double *s_tb, double *c_tb, double *icd_out, double *dbeta) {
  /* Written with reciprocals so that one evaluation costs two divisions (1/hyp, 1/cos(delta))
   * instead of six: the rollout kernels are bound by instruction issue, and an IEEE double
   * division is ~25 instructions on the GPU.  altro_b200/csrc/models.cuh (Bicycle5C::trig) uses
   * exactly this operation order. */
  double L = prm[0], lr = prm[1];
  double theta = x[2], delta = x[3];
  double sd = sin(delta), cd = cos(delta), st = sin(theta), ct = cos(theta);
  double by = lr * delta;
  double h2 = L * L + by * by;
  double hyp = sqrt(h2);
  double inv = 1.0 / hyp;
  double icd = 1.0 / cd;
  *cb = L * inv;
  *sb = by * inv;
  *tand = sd * icd;
  *s_tb = st * (*cb) + ct * (*sb);
  *c_tb = ct * (*cb) - st * (*sb);
  *icd_out = icd;
  *dbeta = (L * lr) * (inv * inv);
}
static void bicycle5_xdot(const double *prm, double *xdot, const double *x, const double *u) {
  double invL = 1.0 / prm[0];
  double v = x[4];
  double sb, cb, tand, s_tb, c_tb, icd, dbeta;
  bicycle5_trig(prm, x, &sb, &cb, &tand, &s_tb, &c_tb, &icd, &dbeta);
  xdot[0] = v * c_tb;
  xdot[1] = v * s_tb;
  xdot[2] = v * cb * tand * invL;
  xdot[3] = u[1];
  xdot[4] = u[0];
}
static void bicycle5_cjac(const double *prm, double *jac, const double *x, const double *u) {
  (void)u;
  double invL = 1.0 / prm[0];
  double v = x[4];
  double sb, cb, tand, s_tb, c_tb, icd, dbeta;
  bicycle5_trig(prm, x, &sb, &cb, &tand, &s_tb, &c_tb, &icd, &dbeta);
  double domega_ddelta = (v * invL) * (-sb * tand * dbeta + cb * (icd * icd));
  double domega_dv = cb * tand * invL;
  const int n = 5;
  memset(jac, 0, sizeof(double) * 5 * 7);
#define J(i, j) jac[(i) + n * (j)]
  J(0, 2) = -v * s_tb;
  J(0, 3) = -v * s_tb * dbeta;
  J(0, 4) = c_tb;
  J(1, 2) = v * c_tb;
  J(1, 3) = v * c_tb * dbeta;
  J(1, 4) = s_tb;
  J(2, 3) = domega_ddelta;
  J(2, 4) = domega_dv;
  J(3, 6) = 1.0; /* d delta' / d u1 */
  J(4, 5) = 1.0; /* d v' / d u0 */
#undef J
}

/* chain: params n, m */
static void chain_xdot(const double *prm, double *xdot, const double *x, const double *u) {
  int n = (int)prm[0], m = (int)prm[1];
  int nq = n / 2;
  for (int i = 0; i < nq; ++i) {
    double ql = (i > 0) ? x[i - 1] : 0.0;
    double qr = (i < nq - 1) ? x[i + 1] : 0.0;
    double acc = -kChainGravity * sin(x[i]) - kChainDamping * x[nq + i] +
                 kChainCoupling * (ql - 2.0 * x[i] + qr);
    if (i < m) acc += u[i];
    xdot[i] = x[nq + i];
    xdot[nq + i] = acc;
  }
}

static void chain_cjac(const double *prm, double *jac, const double *x, const double *u) {
  (void)u;
  int n = (int)prm[0], m = (int)prm[1];
  int nq = n / 2;
  memset(jac, 0, sizeof(double) * n * (n + m));
#define J(i, j) jac[(i) + n * (j)]
  for (int i = 0; i < nq; ++i) {
    J(i, nq + i) = 1.0;
    J(nq + i, i) = -kChainGravity * cos(x[i]) - 2.0 * kChainCoupling;
    if (i > 0) J(nq + i, i - 1) = kChainCoupling;
    if (i < nq - 1) J(nq + i, i + 1) = kChainCoupling;
    J(nq + i, nq + i) = -kChainDamping;
    if (i < m) J(nq + i, n + i) = 1.0;
  }
#undef J
}

void oracle_model_continuous(int model_id, const double *params, double *xdot, const double *x,
                             const double *u) {
  switch (model_id) {
    case ORACLE_MODEL_PENDULUM:
      pendulum_xdot(xdot, x, u);
      break;
    case ORACLE_MODEL_BICYCLE4:
      bicycle4_xdot(params, xdot, x, u);
      break;
    case ORACLE_MODEL_BICYCLE5:
      bicycle5_xdot(params, xdot, x, u);
      break;
    case ORACLE_MODEL_CHAIN:
      chain_xdot(params, xdot, x, u);
      break;
    default:
      break;
  }
}

void oracle_model_continuous_jacobian(int model_id, const double *params, double *jac,
                                      const double *x, const double *u) {
  switch (model_id) {
    case ORACLE_MODEL_PENDULUM:
      pendulum_cjac(jac, x, u);
      break;
    case ORACLE_MODEL_BICYCLE4:
      bicycle4_cjac(params, jac, x, u);
      break;
    case ORACLE_MODEL_BICYCLE5:
      bicycle5_cjac(params, jac, x, u);
      break;
    case ORACLE_MODEL_CHAIN:
      chain_cjac(params, jac, x, u);
      break;
    default:
      break;
  }
}

/* test_utils.cpp:18-26 */
static void di_dynamics(int dim, double *xnext, const double *x, const double *u, float h) {
  double b = h * h / 2; /* float arithmetic, then widened (test_utils.cpp:20) */
  for (int i = 0; i < dim; ++i) {
    xnext[i] = x[i] + x[i + dim] * h + u[i] * b;
    xnext[i + dim] = x[i + dim] + u[i] * h;
  }
}

/* test_utils.cpp:28-41 */
static void di_jacobian(int dim, double *jac, float h) {
  int n = 2 * dim;
  memset(jac, 0, sizeof(double) * n * 3 * dim);
  double b = h * h / 2;
#define J(i, j) jac[(i) + n * (j)]
  for (int i = 0; i < dim; ++i) {
    J(i, i) = 1.0;
    J(i + dim, i + dim) = 1.0;
    J(i, i + dim) = h;
    J(i, 2 * dim + i) = b;
    J(i + dim, 2 * dim + i) = h;
  }
#undef J
}

/* test_utils.cpp:84-97 (MidpointDynamics) */
static void midpoint_dynamics(int model_id, const double *prm, int n, double *xn, const double *x,
                              const double *u, float h) {
  double xm[NMAX];
  double hh = h / 2; /* float division */
  oracle_model_continuous(model_id, prm, xm, x, u);
  for (int i = 0; i < n; ++i) xm[i] *= hh;
  for (int i = 0; i < n; ++i) xm[i] += x[i];
  oracle_model_continuous(model_id, prm, xn, xm, u);
  for (int i = 0; i < n; ++i) xn[i] = x[i] + h * xn[i];
}

/* test_utils.cpp:99-132 (MidpointJacobian) */
static void midpoint_jacobian(int model_id, const double *prm, int n, int m, double *jac,
                              const double *x, const double *u, float h) {
  double xm[NMAX], J0[NMAX * (NMAX + NMAX)], Jm[NMAX * (NMAX + NMAX)];
  double T[NMAX * NMAX], M[NMAX * NMAX];
  double hh = h / 2; /* float division */
  double hd = h;
  oracle_model_continuous(model_id, prm, xm, x, u);
  for (int i = 0; i < n; ++i) xm[i] = x[i] + hh * xm[i];
  oracle_model_continuous_jacobian(model_id, prm, J0, x, u);   /* A = J0[:, :n], B = J0[:, n:] */
  oracle_model_continuous_jacobian(model_id, prm, Jm, xm, u);  /* Am, Bm */
  const double *A = J0, *B = J0 + n * n, *Am = Jm, *Bm = Jm + n * n;
  /* A_d = I + h * Am * (I + h/2 * A) */
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) T[i + n * j] = (i == j ? 1.0 : 0.0) + hh * A[i + n * j];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int l = 0; l < n; ++l) s += Am[i + n * l] * T[l + n * j];
      M[i + n * j] = s;
    }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) jac[i + n * j] = (i == j ? 1.0 : 0.0) + hd * M[i + n * j];
  /* B_d = h * (Am * h/2 * B + Bm) */
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int l = 0; l < n; ++l) s += (Am[i + n * l] * hh) * B[l + n * j];
      jac[i + n * (n + j)] = hd * (s + Bm[i + n * j]);
    }
}

void oracle_model_dynamics(int model_id, const double *params, double *xn, const double *x,
                           const double *u, float h) {
  int n, m;
  oracle_model_dims(model_id, params, &n, &m);
  if (model_id == ORACLE_MODEL_DOUBLE_INTEGRATOR) {
    di_dynamics((int)params[0], xn, x, u, h);
  } else {
    midpoint_dynamics(model_id, params, n, xn, x, u, h);
  }
}

void oracle_model_jacobian(int model_id, const double *params, double *jac, const double *x,
                           const double *u, float h) {
  int n, m;
  oracle_model_dims(model_id, params, &n, &m);
  if (model_id == ORACLE_MODEL_DOUBLE_INTEGRATOR) {
    di_jacobian((int)params[0], jac, h);
  } else {
    midpoint_jacobian(model_id, params, n, m, jac, x, u, h);
  }
}
