// models.cuh -- device dynamics models.
//
// The reference takes dynamics as host std::function callbacks (typedefs.hpp:31-35), which
// cannot run on the device; the batched solver selects one of these compiled-in models by id
// instead (the one API extension, DESIGN.md "boundary").  Each model mirrors what the
// reference's tests plug in:
//   DoubleIntegrator<DIM>   test/test_utils.cpp:18-41  (closed-form discrete, `b = h*h/2` in float)
//   PendulumC               test/test_utils.cpp:43-82
//   Bicycle4C               test/test_utils.cpp:134-238 (centre-of-gravity frame)
//   Midpoint<Continuous>    test/test_utils.cpp:84-132  (explicit midpoint; `h/2` is a float op)
// plus the two models BASELINE.json needs that the reference does not have:
//   Bicycle5C     [x,y,theta,delta,v], u=[a,delta_dot]
//   ChainC<n,m>   coupled pendulum chain
// Model ids must match altro_b200/problems.py and include/altro_b200.h.
#pragma once
#include "linalg.cuh"
#include "fastmath.cuh"

namespace altro_b200 {

enum ModelId {
  MODEL_LINEAR = 0,
  MODEL_DOUBLE_INTEGRATOR = 1,
  MODEL_PENDULUM = 2,
  MODEL_BICYCLE4 = 3,
  MODEL_BICYCLE5 = 4,
  MODEL_CHAIN = 5,
};

// ------------------------------------------------------------------ double integrator
template <int DIM>
struct DoubleIntegrator {
  static constexpr int n = 2 * DIM;
  static constexpr int m = DIM;
  ALTRO_DEV static void dynamics(const double* prm, const double* x, const double* u, float h,
                                 double* xn) {
    (void)prm;
    const double b = h * h / 2;  // float arithmetic, then widened (test_utils.cpp:20)
    const double hd = h;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      xn[i] = x[i] + x[i + DIM] * hd + u[i] * b;
      xn[i + DIM] = x[i + DIM] + u[i] * hd;
    }
  }
  ALTRO_DEV static void jacobian(const double* prm, const double* x, const double* u, float h,
                                 double* A, double* B) {
    (void)prm;
    (void)x;
    (void)u;
    const double b = h * h / 2;
    const double hd = h;
#pragma unroll
    for (int i = 0; i < n * n; ++i) A[i] = 0.0;
#pragma unroll
    for (int i = 0; i < n * m; ++i) B[i] = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      A[i + n * i] = 1.0;
      A[(i + DIM) + n * (i + DIM)] = 1.0;
      A[i + n * (i + DIM)] = hd;
      B[i + n * i] = b;
      B[(i + DIM) + n * i] = hd;
    }
  }
};

// ------------------------------------------------------------------ continuous models
struct PendulumC {
  static constexpr int n = 2;
  static constexpr int m = 1;
  ALTRO_DEV static void xdot(const double* prm, const double* x, const double* u, double* f) {
    (void)prm;
    const double l = 0.5, g = 9.81, b = 0.1;
    const double mm_ = 1.0 * l * l;
    f[0] = x[1];
    f[1] = u[0] / mm_ - g * sin(x[0]) / l - b * x[1] / mm_;
  }
  // J = [A B], column-major n x (n+m)
  ALTRO_DEV static void jac(const double* prm, const double* x, const double* u, double* A,
                            double* B) {
    (void)prm;
    (void)u;
    const double l = 0.5, g = 9.81, b = 0.1;
    const double mm_ = 1.0 * l * l;
    A[0] = 0.0;
    A[1] = -g * cos(x[0]) / l;
    A[2] = 1.0;
    A[3] = -b / mm_;
    B[0] = 0.0;
    B[1] = 1 / mm_;
  }
};

struct Bicycle4C {  // prm[0]=L, prm[1]=lr
  static constexpr int n = 4;
  static constexpr int m = 2;
  ALTRO_DEV static void xdot(const double* prm, const double* x, const double* u, double* f) {
    const double L = prm[0], lr = prm[1];
    const double v = u[0], theta = x[2], delta = x[3];
    const double beta = atan2(lr * delta, L);
    const double omega = v * cos(beta) * tan(delta) / L;
    double s, c;
    sincos(theta + beta, &s, &c);
    f[0] = v * c;
    f[1] = v * s;
    f[2] = omega;
    f[3] = u[1];
  }
  ALTRO_DEV static void jac(const double* prm, const double* x, const double* u, double* A,
                            double* B) {
    const double L = prm[0], lr = prm[1];
    const double v = u[0], theta = x[2], delta = x[3];
    const double by = lr * delta, bx = L;
    const double beta = atan2(by, bx);
    const double dbeta_ddelta = bx / (bx * bx + by * by) * lr;
    double sb, cb, sd, cd, st, ct;
    sincos(beta, &sb, &cb);
    sincos(delta, &sd, &cd);
    sincos(theta + beta, &st, &ct);
    const double tand = tan(delta);
    const double domega_ddelta = v / L * (-sb * tand * dbeta_ddelta + cb / (cd * cd));
    const double domega_dv = cb * tand / L;
#pragma unroll
    for (int i = 0; i < n * n; ++i) A[i] = 0.0;
#pragma unroll
    for (int i = 0; i < n * m; ++i) B[i] = 0.0;
    A[0 + n * 2] = v * -st;
    A[0 + n * 3] = v * (-st * dbeta_ddelta);
    B[0 + n * 0] = ct;
    A[1 + n * 2] = v * ct;
    A[1 + n * 3] = v * (ct * dbeta_ddelta);
    B[1 + n * 0] = st;
    A[2 + n * 3] = domega_ddelta;
    B[2 + n * 0] = domega_dv;
    B[3 + n * 1] = 1.0;
  }
};

struct Bicycle5C {  // state [x,y,theta,delta,v], input [a, delta_dot]
  static constexpr int n = 5;
  static constexpr int m = 2;
  // Defined (here and in oracle/models.c, in exactly this operation order) in the algebraic form
  // that avoids atan2: with beta = atan(lr*delta/L): cos(beta) = L/hyp, sin(beta) = lr*delta/hyp,
  // hyp = sqrt(L^2 + (lr*delta)^2); sin/cos(theta+beta) by angle addition.  Two independent
  // sincos per evaluation instead of a chain of five transcendental calls, and reciprocals so that
  // an evaluation costs two IEEE divisions (1/hyp, 1/cos delta; 1/L is loop-invariant) instead of
  // six -- the rollout kernels are bound by instruction issue (ncu r01 v9) and a double division
  // is ~25 instructions.
  struct Trig {
    double sb, cb, tand, s_tb, c_tb, icd, dbeta;
  };
  ALTRO_DEV static Trig trig(const double* prm, const double* x) {
    const double L = prm[0], lr = prm[1];
    // sincos and 1/x through fastmath.cuh: the library's bits, but one range test per pair of calls
    // instead of a convergence region around each, so the four chains interleave
    double sd, cd, st, ct;
    if (sincos_in_range(x[3]) && sincos_in_range(x[2])) {
      sincos_inrange(x[3], &sd, &cd);
      sincos_inrange(x[2], &st, &ct);
    } else {
      sincos(x[3], &sd, &cd);
      sincos(x[2], &st, &ct);
    }
    const double by = lr * x[3];
    const double h2 = L * L + by * by;
    const double hyp = sqrt(h2);
    double inv, icd;
    if (rcp_in_range(hyp) && rcp_in_range(cd)) {
      inv = rcp_inrange(hyp);
      icd = rcp_inrange(cd);
    } else {
      inv = 1.0 / hyp;
      icd = 1.0 / cd;
    }
    Trig t;
    t.cb = L * inv;
    t.sb = by * inv;
    t.tand = sd * icd;
    t.s_tb = st * t.cb + ct * t.sb;
    t.c_tb = ct * t.cb - st * t.sb;
    t.icd = icd;
    t.dbeta = (L * lr) * (inv * inv);
    return t;
  }
  ALTRO_DEV static void xdot(const double* prm, const double* x, const double* u, double* f) {
    const double invL = 1.0 / prm[0];
    const double v = x[4];
    const Trig t = trig(prm, x);
    f[0] = v * t.c_tb;
    f[1] = v * t.s_tb;
    f[2] = v * t.cb * t.tand * invL;
    f[3] = u[1];
    f[4] = u[0];
  }
  // f(x,u) and its Jacobian at the same point from ONE trig evaluation (the midpoint Jacobian
  // needs both at x; same values as xdot() and jac())
  ALTRO_DEV static void xdot_jac(const double* prm, const double* x, const double* u, double* f,
                                 double* A, double* B) {
    const double invL = 1.0 / prm[0];
    const double v = x[4];
    const Trig t = trig(prm, x);
    f[0] = v * t.c_tb;
    f[1] = v * t.s_tb;
    f[2] = v * t.cb * t.tand * invL;
    f[3] = u[1];
    f[4] = u[0];
    fill_jac(t, v, invL, A, B);
  }
  ALTRO_DEV static void jac(const double* prm, const double* x, const double* u, double* A,
                            double* B) {
    (void)u;
    const double invL = 1.0 / prm[0];
    fill_jac(trig(prm, x), x[4], invL, A, B);
  }
  ALTRO_DEV static void fill_jac(const Trig& t, double v, double invL, double* A, double* B) {
    const double domega_ddelta = (v * invL) * (-t.sb * t.tand * t.dbeta + t.cb * (t.icd * t.icd));
    const double domega_dv = t.cb * t.tand * invL;
#pragma unroll
    for (int i = 0; i < n * n; ++i) A[i] = 0.0;
#pragma unroll
    for (int i = 0; i < n * m; ++i) B[i] = 0.0;
    A[0 + n * 2] = -v * t.s_tb;
    A[0 + n * 3] = -v * t.s_tb * t.dbeta;
    A[0 + n * 4] = t.c_tb;
    A[1 + n * 2] = v * t.c_tb;
    A[1 + n * 3] = v * t.c_tb * t.dbeta;
    A[1 + n * 4] = t.s_tb;
    A[2 + n * 3] = domega_ddelta;
    A[2 + n * 4] = domega_dv;
    B[3 + n * 1] = 1.0;
    B[4 + n * 0] = 1.0;
  }
};

template <int NS, int NI>
struct ChainC {  // q_i'' = -g sin q_i - b q_i' + kc (q_{i-1} - 2 q_i + q_{i+1}) + u_i [i < m]
  static constexpr int n = NS;
  static constexpr int m = NI;
  static constexpr int nq = NS / 2;
  ALTRO_DEV static void xdot(const double* prm, const double* x, const double* u, double* f) {
    (void)prm;
    const double g = 9.81, b = 0.1, kc = 1.0;
#pragma unroll
    for (int i = 0; i < nq; ++i) {
      const double ql = (i > 0) ? x[i - 1] : 0.0;
      const double qr = (i < nq - 1) ? x[i + 1] : 0.0;
      double acc = -g * sin(x[i]) - b * x[nq + i] + kc * (ql - 2.0 * x[i] + qr);
      if (i < m) acc += u[i];
      f[i] = x[nq + i];
      f[nq + i] = acc;
    }
  }
  ALTRO_DEV static void jac(const double* prm, const double* x, const double* u, double* A,
                            double* B) {
    (void)prm;
    (void)u;
    const double g = 9.81, b = 0.1, kc = 1.0;
    for (int i = 0; i < n * n; ++i) A[i] = 0.0;
    for (int i = 0; i < n * m; ++i) B[i] = 0.0;
#pragma unroll
    for (int i = 0; i < nq; ++i) {
      A[i + n * (nq + i)] = 1.0;
      A[(nq + i) + n * i] = -g * cos(x[i]) - 2.0 * kc;
      if (i > 0) A[(nq + i) + n * (i - 1)] = kc;
      if (i < nq - 1) A[(nq + i) + n * (i + 1)] = kc;
      A[(nq + i) + n * (nq + i)] = -b;
      if (i < m) B[(nq + i) + n * i] = 1.0;
    }
  }
};

// detects an optional fused CM::xdot_jac
template <class CM, class = void>
struct has_xdot_jac {
  static constexpr bool value = false;
};
template <class CM>
struct has_xdot_jac<CM, decltype(CM::xdot_jac(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), void())> {
  static constexpr bool value = true;
};

// detects an optional fused Model::dynamics_jacobian
template <class M, class = void>
struct has_dynamics_jacobian {
  static constexpr bool value = false;
};
template <class M>
struct has_dynamics_jacobian<M, decltype(M::dynamics_jacobian(nullptr, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr), void())> {
  static constexpr bool value = true;
};

// ------------------------------------------------------------------ explicit midpoint
template <class CM>
struct Midpoint {
  static constexpr int n = CM::n;
  static constexpr int m = CM::m;
  // test_utils.cpp:84-97
  ALTRO_DEV static void dynamics(const double* prm, const double* x, const double* u, float h,
                                 double* xn) {
    const double hh = h / 2;  // float division, then widened
    const double hd = h;
    double xm[n];
    CM::xdot(prm, x, u, xm);
#pragma unroll
    for (int i = 0; i < n; ++i) xm[i] = fma(xm[i], hh, x[i]);
    double f[n];
    CM::xdot(prm, xm, u, f);
#pragma unroll
    for (int i = 0; i < n; ++i) xn[i] = fma(hd, f[i], x[i]);
  }
  // dynamics() and jacobian() of the same point in one go: the two share the evaluation of the
  // continuous model at x and at the midpoint (for the bicycle: the trigonometry, by far the most
  // expensive part).  Same functions on the same inputs in the same order as the separate calls,
  // so xn, Ad, Bd carry the same bits.
  ALTRO_DEV static void dynamics_jacobian(const double* prm, const double* x, const double* u, float h,
                                          double* xn, double* Ad, double* Bd) {
    const double hh = h / 2;
    const double hd = h;
    double xm[n], f[n];
    double T[n * n], Bc[n * m], Am[n * n], Bm[n * m];
    if constexpr (has_xdot_jac<CM>::value) {
      CM::xdot_jac(prm, x, u, xm, T, Bc);
    } else {
      CM::xdot(prm, x, u, xm);
      CM::jac(prm, x, u, T, Bc);
    }
#pragma unroll
    for (int i = 0; i < n; ++i) xm[i] = fma(xm[i], hh, x[i]);
    if constexpr (has_xdot_jac<CM>::value) {
      CM::xdot_jac(prm, xm, u, f, Am, Bm);
    } else {
      CM::xdot(prm, xm, u, f);
      CM::jac(prm, xm, u, Am, Bm);
    }
#pragma unroll
    for (int i = 0; i < n; ++i) xn[i] = fma(hd, f[i], x[i]);
#pragma unroll
    for (int j = 0; j < n; ++j)
#pragma unroll
      for (int i = 0; i < n; ++i) T[i + n * j] = (i == j ? 1.0 : 0.0) + hh * T[i + n * j];
#pragma unroll
    for (int i = 0; i < n * m; ++i) Bc[i] *= hh;
    mm<n, n, n, false, false, 0>(Am, T, Ad);
#pragma unroll
    for (int j = 0; j < n; ++j)
#pragma unroll
      for (int i = 0; i < n; ++i) Ad[i + n * j] = (i == j ? 1.0 : 0.0) + hd * Ad[i + n * j];
    mm<n, m, n, false, false, 0>(Am, Bc, Bd);
#pragma unroll
    for (int i = 0; i < n * m; ++i) Bd[i] = hd * (Bd[i] + Bm[i]);
  }
  // test_utils.cpp:99-132:  A_d = I + h Am (I + h/2 A),  B_d = h (Am h/2 B + Bm)
  ALTRO_DEV static void jacobian(const double* prm, const double* x, const double* u, float h,
                                 double* Ad, double* Bd) {
    const double hh = h / 2;
    const double hd = h;
    double xm[n];
    double T[n * n], Bc[n * m], Am[n * n], Bm[n * m];
    if constexpr (has_xdot_jac<CM>::value) {
      CM::xdot_jac(prm, x, u, xm, T, Bc);
    } else {
      CM::xdot(prm, x, u, xm);
      CM::jac(prm, x, u, T, Bc);
    }
#pragma unroll
    for (int i = 0; i < n; ++i) xm[i] = fma(hh, xm[i], x[i]);
    CM::jac(prm, xm, u, Am, Bm);
    // T = I + h/2 A ; Bc = h/2 B
#pragma unroll
    for (int j = 0; j < n; ++j)
#pragma unroll
      for (int i = 0; i < n; ++i) T[i + n * j] = (i == j ? 1.0 : 0.0) + hh * T[i + n * j];
#pragma unroll
    for (int i = 0; i < n * m; ++i) Bc[i] *= hh;
    mm<n, n, n, false, false, 0>(Am, T, Ad);
#pragma unroll
    for (int j = 0; j < n; ++j)
#pragma unroll
      for (int i = 0; i < n; ++i) Ad[i + n * j] = (i == j ? 1.0 : 0.0) + hd * Ad[i + n * j];
    mm<n, m, n, false, false, 0>(Am, Bc, Bd);
#pragma unroll
    for (int i = 0; i < n * m; ++i) Bd[i] = hd * (Bd[i] + Bm[i]);
  }
};

using Pendulum = Midpoint<PendulumC>;
using Bicycle4 = Midpoint<Bicycle4C>;
using Bicycle5 = Midpoint<Bicycle5C>;
template <int NS, int NI>
using Chain = Midpoint<ChainC<NS, NI>>;

// ------------------------------------------------------------------ packed Jacobian storage
// What the kernels keep in HBM of the discrete expansion [A B] of one knot.  Generic models store
// all n*n + n*m entries.  Models whose midpoint Jacobian has entries that are the SAME constant
// (0, 1 or h) at every state store only the varying ones and re-create the constants on load:
// the dense blocks the arithmetic sees are bit-identical to the computed ones (a structural zero
// times a finite number is an exact zero), so nothing downstream changes -- only the bytes moved
// by every HBM-bound kernel (bicycle: 15 of 35 doubles).  The one observable difference: a
// trajectory that has already overflowed to inf/NaN would poison the "constant" entries in the
// dense computation (0 * inf) and does not here; such a solve has failed either way.
template <class Model>
struct JacPack {
  static constexpr int n = Model::n, m = Model::m;
  static constexpr int V = n * n + n * m;
  static constexpr bool packed = false;
  ALTRO_DEV static void pack(const double* A, const double* B, double* J) {
#pragma unroll
    for (int i = 0; i < n * n; ++i) J[i] = A[i];
#pragma unroll
    for (int i = 0; i < n * m; ++i) J[n * n + i] = B[i];
  }
  ALTRO_DEV static void unpack(const double* J, float h, double* A, double* B) {
    (void)h;
#pragma unroll
    for (int i = 0; i < n * n; ++i) A[i] = J[i];
#pragma unroll
    for (int i = 0; i < n * m; ++i) B[i] = J[n * n + i];
  }
};

// Bicycle5 (state [x,y,theta,delta,v], input [a,delta_dot]): the continuous Jacobian has rows
// 0..2 x columns 2..4 only and B_c = [e4 e3], so A_d = I + h Am (I + h/2 A) differs from the
// identity in rows 0..2 x columns 2..4 and B_d = h (Am h/2 B + Bm) has rows 0..2 varying,
// B_d[3,1] = B_d[4,0] = h.
template <>
struct JacPack<Midpoint<Bicycle5C>> {
  static constexpr int n = 5, m = 2;
  static constexpr int V = 15;
  static constexpr bool packed = true;
  ALTRO_DEV static void pack(const double* A, const double* B, double* J) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) J[r + 3 * c] = A[r + n * (2 + c)];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) J[9 + r + 3 * c] = B[r + n * c];
  }
  ALTRO_DEV static void unpack(const double* J, float h, double* A, double* B) {
    const double hd = h;
#pragma unroll
    for (int c = 0; c < n; ++c)
#pragma unroll
      for (int r = 0; r < n; ++r) A[r + n * c] = (r == c) ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) A[r + n * (2 + c)] = J[r + 3 * c];
#pragma unroll
    for (int i = 0; i < n * m; ++i) B[i] = 0.0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) B[r + n * c] = J[9 + r + 3 * c];
    B[3 + n * 1] = hd;
    B[4 + n * 0] = hd;
  }
};

// Bicycle4 (test/test_utils.cpp:134-238; state [x,y,theta,delta], input [v,delta_dot]): A_d
// differs from the identity in rows 0..2 x columns 2..3, B_d has rows 0..2 varying, B_d[3,1] = h.
template <>
struct JacPack<Midpoint<Bicycle4C>> {
  static constexpr int n = 4, m = 2;
  static constexpr int V = 12;
  static constexpr bool packed = true;
  ALTRO_DEV static void pack(const double* A, const double* B, double* J) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) J[r + 3 * c] = A[r + n * (2 + c)];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) J[6 + r + 3 * c] = B[r + n * c];
  }
  ALTRO_DEV static void unpack(const double* J, float h, double* A, double* B) {
    const double hd = h;
#pragma unroll
    for (int c = 0; c < n; ++c)
#pragma unroll
      for (int r = 0; r < n; ++r) A[r + n * c] = (r == c) ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) A[r + n * (2 + c)] = J[r + 3 * c];
#pragma unroll
    for (int i = 0; i < n * m; ++i) B[i] = 0.0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) B[r + n * c] = J[6 + r + 3 * c];
    B[3 + n * 1] = hd;
  }
};

}  // namespace altro_b200
