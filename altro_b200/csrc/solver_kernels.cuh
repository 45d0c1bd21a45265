// solver_kernels.cuh -- the batched AL-iLQR hot path, one thread per trajectory.
//
// What the reference runs per problem on one CPU thread (SolverImpl::Solve,
// src/altro/solver/solver.cpp:414-511, with tvlqr_BackwardPass src/tvlqr/tvlqr.cpp:65-195,
// MeritFunction solver.cpp:273-355, the AL terms of knotpoint_data.cpp:473-613 and the cones of
// cones.cpp) runs here as ONE persistent kernel: lane = trajectory, warp = 32 consecutive problems
// of the problem-fastest HBM layout (device_problem.h).  All n x n / n x m / m x m blocks of the
// current knot live in the registers of the owning thread; HBM is touched once per block per
// sweep with fully coalesced 256-byte rows.  Control flow (line-search branches, the
// `stat < sqrt(tol)` dual-update trigger, the `feas > tol` penalty trigger, quirks Q1-Q4 of
// SURVEY.md Appendix C) follows the reference decision for decision so iteration counts match.
#pragma once
#include <math.h>

#include "device_problem.h"
#include "linalg.cuh"
#include "linesearch.cuh"
#include "models.cuh"

namespace altro_b200 {

// ErrorCodes subset (exceptions.hpp:24-51)
enum DevErr { ERR_NONE = 0, ERR_LINESEARCH_FAILED = 20, ERR_MERIT_GRAD_TOO_SMALL = 21 };



// ---------------------------------------------------------------- linear model (shared A,B,f)
template <int NS, int NI>
struct LinearModel {
  static constexpr int n = NS;
  static constexpr int m = NI;
};

template <class Model>
struct ModelTraits {
  static constexpr bool is_linear = false;
};
template <int NS, int NI>
struct ModelTraits<LinearModel<NS, NI>> {
  static constexpr bool is_linear = true;
};

// ---------------------------------------------------------------- second-order cone pieces
// cones.cpp:13-39
ALTRO_DEV void soc_projection(int dim, const double* x, double* px) {
  const int nn = dim - 1;
  const double s = x[nn];
  double a = 0.0;
  for (int i = 0; i < nn; ++i) a += x[i] * x[i];
  a = sqrt(a);
  if (a <= -s) {
    for (int i = 0; i < dim; ++i) px[i] = 0.0;
  } else if (a <= s) {
    for (int i = 0; i < dim; ++i) px[i] = x[i];
  } else {
    const double c = 0.5 * (1 + s / a);
    for (int i = 0; i < nn; ++i) px[i] = c * x[i];
    px[nn] = c * a;
  }
}

// cones.cpp:41-77, column-major dim x dim
ALTRO_DEV void soc_jacobian(int dim, const double* x, double* J) {
  const int nn = dim - 1;
  const double s = x[nn];
  double a = 0.0;
  for (int i = 0; i < nn; ++i) a += x[i] * x[i];
  a = sqrt(a);
  for (int i = 0; i < dim * dim; ++i) J[i] = 0.0;
  if (a <= -s) {
    return;
  } else if (a <= s) {
    for (int i = 0; i < dim; ++i) J[i + dim * i] = 1.0;
  } else {
    const double c = 0.5 * (1 + s / a);
    for (int j = 0; j < nn; ++j)
      for (int i = 0; i < nn; ++i) {
        double v = -0.5 * s / (a * a * a) * x[i] * x[j];
        v += (i == j) ? c : 0;
        J[i + dim * j] = v;
      }
    for (int i = 0; i < nn; ++i) J[i + dim * nn] = 0.5 * x[i] / a;
    for (int j = 0; j < nn; ++j) J[nn + dim * j] = ((-0.5 * s / (a * a)) + c / a) * x[j];
    J[nn + dim * nn] = 0.5;
  }
}

// cones.cpp:79-123
ALTRO_DEV void soc_hessian(int dim, const double* x, const double* b, double* H) {
  const int nn = dim - 1;
  const double s = x[nn], bs = b[nn];
  double vbv = 0, a = 0;
  for (int i = 0; i < nn; ++i) {
    a += x[i] * x[i];
    vbv += x[i] * b[i];
  }
  a = sqrt(a);
  for (int i = 0; i < dim * dim; ++i) H[i] = 0.0;
  if (a <= -s || a <= s) return;
  for (int i = 0; i < nn; ++i) {
    double hi = 0;
    for (int j = 0; j < nn; ++j) {
      double Hij = -x[i] * x[j] / (a * a);
      Hij += (i == j) ? 1 : 0;
      hi += Hij * b[j];
    }
    H[i + dim * nn] = hi / (2 * a);
    H[nn + dim * i] = hi / (2 * a);
    for (int j = 0; j <= i; ++j) {
      const double vij = x[i] * x[j];
      const double H1 = hi * x[j] * (-s / (a * a * a));
      double H2 = vij * (2 * vbv) / (a * a * a * a) - x[i] * b[j] / (a * a);
      double H3 = -vij / (a * a);
      if (i == j) {
        H2 -= vbv / (a * a);
        H3 += 1;
      }
      H2 *= s / a;
      H3 *= bs / a;
      H[i + dim * j] = (H1 + H2 + H3) / 2.0;
      H[j + dim * i] = (H1 + H2 + H3) / 2.0;
    }
  }
  H[nn + dim * nn] = 0.0;
}

// ---------------------------------------------------------------- the per-trajectory solver
// CON: 0 unconstrained, 1 constraints with linear cones only (equality / identity / inequality:
// diagonal projections, no scratch arrays), 2 also second-order cones (the SOC projection,
// Jacobian and Hessian need local scratch that would otherwise bloat every constrained kernel).
template <class Model, int CON>
struct TrajSolver {
  static constexpr int n = Model::n;
  static constexpr int m = Model::m;
  static constexpr int NS_ = Model::n, NI_ = Model::m;
  static constexpr bool kLinear = ModelTraits<Model>::is_linear;
  static constexpr int kCon = CON;
  // TMA staging of the sequential sweeps (solver_phases.cuh) when the blocks are small enough to
  // live in registers (larger n keeps rolled loops over local-memory arrays, is bound by
  // arithmetic rather than by load latency, and a stage would not fit shared memory).
  static constexpr bool kStaged = (NS_ <= kUnrollDim);
  // first row of each field inside a knot record; order = device_problem.h / capi.cu
  static constexpr int rXbar = 0, rUbar = rXbar + NS_, rQ = rUbar + NI_, rR = rQ + NS_, rC = rR + NI_,
                       rK = rC + 1, rD = rK + NI_ * NS_, rX = rD + NI_, rU = rX + NS_, rA = rU + NI_,
                       rB = rA + NS_ * NS_, rLx = rB + NS_ * NI_, rLu = rLx + NS_, rY = rLu + NI_,
                       rP = rY + NS_, rPv = rP + NS_ * NS_, rUinit = rPv + NS_, kRecordRows = rUinit + NI_;

  const DeviceProblem& P;
  // diagonal cost weights of every knot (shared by the batch): HBM by default, the sweep kernels
  // point these at a shared-memory copy so the per-knot reads are LDS broadcasts, not L2 round trips
  const double* Qd;
  const double* Rd;
  const long S;   // knot-record stride (doubles) of the main record stream
  const long go;  // this problem's offset inside a field: group * group_stride + lane
  const int b;   // this thread's problem
  const int N;
  double rho;    // penalty (uniform over this trajectory's constraints)
  int merit_evals;

  ALTRO_DEV TrajSolver(const DeviceProblem& p, int b_)
      : P(p), Qd(p.Qd), Rd(p.Rd), S(p.R), go((long)(b_ >> 5) * p.GS + (b_ & 31)), b(b_), N(p.N), rho(1.0), merit_evals(0) {}

  // [A B] of knot k <-> the packed Jacobian rows of the record (models.cuh, JacPack)
  using JP = JacPack<Model>;
  static constexpr int kV = JP::V;
  ALTRO_DEV void store_jac(int k, const double* A, const double* Bm) const {
    double J[kV];
    JP::pack(A, Bm, J);
    store_block<kV>(P.A + go, S, k, J);
  }
  // time step of knot k (SetTimeStep(h, k_start, k_stop), altro_solver.cpp:49-63): one value for
  // the horizon unless the host installed a per-knot table
  ALTRO_DEV float hof(int k) const { return P.hk ? P.hk[k] : P.h; }
  ALTRO_DEV void load_jac(int k, double* A, double* Bm) const {
    double J[kV];
    load_block<kV>(P.A + go, S, k, J);
    JP::unpack(J, hof(k), A, Bm);
  }
  // the same from a landed stage of the TMA ring (J rows of knot k start at `row`)
  ALTRO_DEV void unstage_jac(const double* stage, int row, int lane, int k, double* A, double* Bm) const {
    double J[kV];
    unstage_block<kV>(stage, row, lane, J);
    JP::unpack(J, hof(k), A, Bm);
  }

  // rows per knot each sequential sweep stages through the TMA ring (solver_phases.cuh)
  static constexpr int kRowsBw = kV + NS_ + NI_;                          // [J] [lx lu]
  static constexpr int kRowsPhi = rA - rQ + kV;                           // [q r c K d x u J]
  static constexpr int kRowsRoll = rD + NI_;                              // [xbar ubar q r c K d]
  static constexpr int kRowsDphi = NI_ * NS_ + NI_ + kV + NS_ + NI_;      // [K d] [J] [lx lu]
  static constexpr int kRowsBackwardKernel = kRowsBw > kRowsPhi ? kRowsBw : kRowsPhi;

  // field pointer of this problem (knot 0); knot k is k * S further, rows are 32 doubles apart
  ALTRO_DEV double* F(double* field) const { return field + go; }
  ALTRO_DEV const double* F(const double* field) const { return field + go; }
  // per-group blocks that are not per knot: [group][rows][32]
  ALTRO_DEV const double* G(const double* base, int rows) const {
    return base + (long)(b >> 5) * rows * 32 + (b & 31);
  }
  // Staged copy of the CURRENT knot's dual record [z rows | z_est rows] for this lane (set by the
  // TMA-staged sweeps before they call a step; null: read HBM directly).  Row r of z is
  // zstage[r * 32], of z_est zstage[(zrows + r) * 32].
  const double* zstage = nullptr;
  // zrow = zoff(k, row0): HBM offset of the slot's first row at knot k; i = row inside the slot
  ALTRO_DEV double zread(long zrow, int row0, int i) const {
    return zstage ? zstage[(row0 + i) * 32] : P.z[zrow + i * 32];
  }
  ALTRO_DEV double zest_read(long zrow, int row0, int i) const {
    return zstage ? zstage[(P.zrows + row0 + i) * 32] : P.zest[zrow + i * 32];
  }
  // constraint duals live in their own record stream [group][knot][z rows | z_est rows][32]
  ALTRO_DEV long zoff(int k, int row0) const {
    return (long)(b >> 5) * P.GSz + (long)k * P.Rz + (long)row0 * 32 + (b & 31);
  }

  // ---- selector helpers (static indices only, so x/u stay in registers)
  // The selected variable is the same for every lane of the warp (the constraint table is shared by
  // the batch), so a switch is a uniform jump instead of a select chain over all n + m variables
  // (2 x 21 and 2 x 28 instructions per row in the rollout / follower paths of the constrained
  // kernels); the values are the same: x - 0.0 == x for the entries the chain left alone.
  template <int E>
  ALTRO_DEV static double pick_from(int e, const double* v) {
    if constexpr (E == 1) {
      return v[0];
    } else {
      switch (e) {
#define ALTRO_PICK_CASE(c) \
  case c:                  \
    if constexpr (c < E) return v[c]; else return 0.0;
        ALTRO_PICK_CASE(0) ALTRO_PICK_CASE(1) ALTRO_PICK_CASE(2) ALTRO_PICK_CASE(3)
        ALTRO_PICK_CASE(4) ALTRO_PICK_CASE(5) ALTRO_PICK_CASE(6) ALTRO_PICK_CASE(7)
        ALTRO_PICK_CASE(8) ALTRO_PICK_CASE(9) ALTRO_PICK_CASE(10) ALTRO_PICK_CASE(11)
        ALTRO_PICK_CASE(12) ALTRO_PICK_CASE(13) ALTRO_PICK_CASE(14) ALTRO_PICK_CASE(15)
#undef ALTRO_PICK_CASE
        default:
          return 0.0;
      }
    }
  }
  template <int E>
  ALTRO_DEV static void sub_at(int e, double val, double* v) {
    switch (e) {
#define ALTRO_SUB_CASE(c)            \
  case c:                            \
    if constexpr (c < E) v[c] -= val; \
    break;
      ALTRO_SUB_CASE(0) ALTRO_SUB_CASE(1) ALTRO_SUB_CASE(2) ALTRO_SUB_CASE(3)
      ALTRO_SUB_CASE(4) ALTRO_SUB_CASE(5) ALTRO_SUB_CASE(6) ALTRO_SUB_CASE(7)
      ALTRO_SUB_CASE(8) ALTRO_SUB_CASE(9) ALTRO_SUB_CASE(10) ALTRO_SUB_CASE(11)
      ALTRO_SUB_CASE(12) ALTRO_SUB_CASE(13) ALTRO_SUB_CASE(14) ALTRO_SUB_CASE(15)
#undef ALTRO_SUB_CASE
      default:
        break;
    }
  }
  static_assert(NS_ <= 16 && NI_ <= 16, "selector switch covers 16 variables per block");
  ALTRO_DEV static double pick(int id, const double* x, const double* u) {
    if (id < 0) return 0.0;
    return id < n ? pick_from<n>(id, x) : pick_from<m>(id - n, u);
  }
  ALTRO_DEV static void scatter_sub(int id, double val, double* lx, double* lu, bool terminal) {
    if (id < 0) return;
    if (id < n) {
      sub_at<n>(id, val, lx);
    } else if (!terminal) {
      sub_at<m>(id - n, val, lu);
    }
  }
  ALTRO_DEV double row_offset(const ConSlot& s, int i) const {
    return s.off_per_problem ? G(s.off_b, s.dim)[i * 32] : s.off[i];
  }
  ALTRO_DEV double row_value(const ConSlot& s, int i, const double* x, const double* u) const {
    const int id = s.idx[i];
    const double off = row_offset(s, i);
    return (id < 0) ? off : fma(s.scale[i], pick(id, x, u), off);
  }

  // ---- dynamics through the model (or the shared linear table)
  ALTRO_DEV void dynamics(int k, const double* x, const double* u, double* xn) const {
    if constexpr (kLinear) {
      const double* T = P.lin + (long)k * (n * n + n * m + n);
      // xnext = A x + B u + affine_term   (knotpoint_data.cpp:712-714)
      double t[n];
      mm<n, 1, n, false, false, 0>(T, x, t);
      mm<n, 1, m, false, false, 1>(T + n * n, u, t);
#pragma unroll
      for (int i = 0; i < n; ++i) xn[i] = t[i] + T[n * n + n * m + i];
    } else {
      Model::dynamics(P.model_params, x, u, hof(k), xn);
    }
  }
  ALTRO_DEV void jacobian(int k, const double* x, const double* u, double* A, double* B) const {
    if constexpr (kLinear) {
      const double* T = P.lin + (long)k * (n * n + n * m + n);
#pragma unroll
      for (int i = 0; i < n * n; ++i) A[i] = T[i];
#pragma unroll
      for (int i = 0; i < n * m; ++i) B[i] = T[n * n + i];
    } else {
      Model::jacobian(P.model_params, x, u, hof(k), A, B);
    }
  }

  // x+ = f(x,u) and [A B] = df/d[x;u] of the same point, sharing what the two have in common when
  // the model offers it (same values as dynamics() + jacobian())
  ALTRO_DEV void dynamics_jacobian(int k, const double* x, const double* u, double* xn, double* A,
                                   double* B) const {
    if constexpr (!kLinear && has_dynamics_jacobian<Model>::value) {
      Model::dynamics_jacobian(P.model_params, x, u, hof(k), xn, A, B);
    } else {
      dynamics(k, x, u, xn);
      jacobian(k, x, u, A, B);
    }
  }

  // ---- original (diagonal LQR) cost, knotpoint_data.cpp:636-645, :670-678
  ALTRO_DEV double stage_cost(int k, const double* x, const double* u, const double* q,
                              const double* r, bool terminal) const {
    return stage_cost(k, x, u, q, r, terminal, F(P.c)[(long)k * S]);
  }
  // cval: the constant term c_k of this problem (already loaded / staged)
  ALTRO_DEV double stage_cost(int k, const double* x, const double* u, const double* q,
                              const double* r, bool terminal, double cval) const {
    if constexpr (CON == 2) {
      if (P.Qf) {  // dense Q, R, H: knotpoint_data.cpp:620-634
        double t[n > m ? n : m];
        mm<n, 1, n, false, false, 0>(P.Qf + (long)k * n * n, x, t);
        double J = 0.5 * dot<n>(x, t);
        J += dot<n>(q, x);
        if (!terminal) {
          mm<m, 1, m, false, false, 0>(P.Rf + (long)k * m * m, u, t);
          J += 0.5 * dot<m>(u, t);
          J += dot<m>(r, u);
          mm<m, 1, n, false, false, 0>(P.Hf + (long)k * m * n, x, t);
          J += dot<m>(u, t);
        }
        J += cval;
        return J;
      }
    }
    double J = 0.0;
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) a += (0.5 * x[i]) * Qd[k * n + i] * x[i];
    J = a;
    J += dot<n>(q, x);
    if (!terminal) {
      double bb = 0.0;
#pragma unroll
      for (int i = 0; i < m; ++i) bb += (0.5 * u[i]) * Rd[k * m + i] * u[i];
      J += bb;
      J += dot<m>(r, u);
    }
    J += cval;
    return J;
  }
  ALTRO_DEV void stage_gradient(int k, const double* x, const double* u, const double* q,
                                const double* r, bool terminal, double* lx, double* lu) const {
    if constexpr (CON == 2) {
      if (P.Qf) {  // knotpoint_data.cpp:654-668
        mm<n, 1, n, false, false, 0>(P.Qf + (long)k * n * n, x, lx);
#pragma unroll
        for (int i = 0; i < n; ++i) lx[i] += q[i];
        if (!terminal) {
          mm<m, 1, m, false, false, 0>(P.Rf + (long)k * m * m, u, lu);
#pragma unroll
          for (int i = 0; i < m; ++i) lu[i] += r[i];
          mm<m, 1, n, false, false, 1>(P.Hf + (long)k * m * n, x, lu);
          mm<n, 1, m, true, false, 1>(P.Hf + (long)k * m * n, u, lx);
        }
        return;
      }
    }
#pragma unroll
    for (int i = 0; i < n; ++i) lx[i] = Qd[k * n + i] * x[i] + q[i];
    if (!terminal) {
#pragma unroll
      for (int i = 0; i < m; ++i) lu[i] = Rd[k * m + i] * u[i] + r[i];
    }
  }

  // ---- general constraint slots (CON == 2 only): any family, any cone, dense Jacobian.
  // A slot takes this path when its cone is the second-order cone or its family is not the
  // selector family; selector rows with a linear cone keep the diagonal fast path below.
  ALTRO_DEV static bool general_slot(const ConSlot& s) {
    return CON == 2 && (s.cone == CONE_SOC || s.family != CON_FAMILY_SELECTOR);
  }
  // c(x,u) [p] and, when J != null, dc/d[x;u] [p x (n+m)] column-major (typedefs.hpp:48-52)
  ALTRO_DEV void con_eval(const ConSlot& s, const double* x, const double* u, double* c, double* J) const {
    constexpr int nm = n + m;
    const int p = s.dim;
    double xu[nm];
#pragma unroll
    for (int i = 0; i < n; ++i) xu[i] = x[i];
#pragma unroll
    for (int i = 0; i < m; ++i) xu[n + i] = u[i];
    if (J)
      for (int i = 0; i < p * nm; ++i) J[i] = 0.0;
    if (s.family == CON_FAMILY_AFFINE) {
      for (int i = 0; i < p; ++i) {
        double v = row_offset(s, i);
        for (int j = 0; j < nm; ++j) v = fma(s.Jd[i + p * j], xu[j], v);
        c[i] = v;
      }
      if (J)
        for (int i = 0; i < p * nm; ++i) J[i] = s.Jd[i];
    } else if (s.family == CON_FAMILY_DISC) {
      const double cx = s.off_per_problem ? G(s.off_b, 3)[0] : s.off[0];
      const double cy = s.off_per_problem ? G(s.off_b, 3)[32] : s.off[1];
      const double rad = s.off_per_problem ? G(s.off_b, 3)[64] : s.off[2];
      const double da = xu[s.idx[0]] - cx, db = xu[s.idx[1]] - cy;
      c[0] = rad * rad - da * da - db * db;
      if (J) {
        J[p * s.idx[0]] = -2.0 * da;
        J[p * s.idx[1]] = -2.0 * db;
      }
    } else {  // selector rows
      for (int i = 0; i < p; ++i) {
        const int id = s.idx[i];
        const double off = row_offset(s, i);
        c[i] = (id < 0) ? off : fma(s.scale[i], xu[id], off);
        if (J && id >= 0) J[i + p * id] = s.scale[i];
      }
    }
  }
  // projection onto the DUAL cone of `cone` (cones.hpp:13-30, cones.cpp:125-150)
  ALTRO_DEV static void project_dual(int cone, int p, const double* zt, double* zp) {
    if (cone == CONE_SOC) {
      soc_projection(p, zt, zp);
    } else {
      for (int i = 0; i < p; ++i) {
        double v = 0.0;  // IDENTITY -> dual EQUALITY: {0}
        if (cone == CONE_EQUALITY) v = zt[i];
        if (cone == CONE_INEQUALITY) v = fmin(0.0, zt[i]);
        zp[i] = v;
      }
    }
  }
  // || Pi_K(c) - c ||_inf of one general slot (knotpoint_data.cpp:489-501)
  __device__ __noinline__ double al_violation_general(const ConSlot& s, const double* x, const double* u) const {
    double c[kMaxConDim], pc[kMaxConDim];
    con_eval(s, x, u, c, nullptr);
    if (s.cone == CONE_SOC) {
      soc_projection(s.dim, c, pc);
    } else {
      for (int i = 0; i < s.dim; ++i) {
        double v = 0.0;  // EQUALITY: projection onto {0}
        if (s.cone == CONE_IDENTITY) v = c[i];
        if (s.cone == CONE_INEQUALITY) v = fmin(0.0, c[i]);
        pc[i] = v;
      }
    }
    double viol = 0.0;
    for (int i = 0; i < s.dim; ++i) viol = fmax(viol, fabs(pc[i] - c[i]));
    return viol;
  }
  // AL cost and gradient of one general slot (knotpoint_data.cpp:523-547, :572-595)
  __device__ __noinline__ double al_terms_general(const ConSlot& s, long zrow, const double* x, const double* u,
                                    bool terminal, bool want_grad, double* lx, double* lu,
                                    bool store_zest) {
    constexpr int nm = n + m;
    const int p = s.dim;
    double c[kMaxConDim], zt[kMaxConDim], zp[kMaxConDim], v[kMaxConDim];
    double J[kMaxConDim * nm];
    con_eval(s, x, u, c, want_grad ? J : nullptr);
    for (int i = 0; i < p; ++i) {
      zt[i] = zread(zrow, s.row0, i) - rho * c[i];
      if (store_zest) P.zest[zrow + i * 32] = zt[i];
    }
    project_dual(s.cone, p, zt, zp);
    double nrm = 0.0;
    for (int i = 0; i < p; ++i) nrm += zp[i] * zp[i];
    if (want_grad) {
      if (s.cone == CONE_SOC) {  // v = dPi^T zp
        double Jp[kMaxSocDim * kMaxSocDim];
        soc_jacobian(p, zt, Jp);
        for (int i = 0; i < p; ++i) {
          double a = 0.0;
          for (int l = 0; l < p; ++l) a += Jp[l + p * i] * zp[l];
          v[i] = a;
        }
      } else {  // dPi is diagonal with entries {1 | zt<=0 | 0}: dPi^T zp == zp
        for (int i = 0; i < p; ++i) v[i] = zp[i];
      }
#pragma unroll
      for (int j = 0; j < n; ++j) {
        double a = 0.0;
        for (int i = 0; i < p; ++i) a = fma(J[i + p * j], v[i], a);
        lx[j] -= a;
      }
      if (!terminal) {
#pragma unroll
        for (int j = 0; j < m; ++j) {
          double a = 0.0;
          for (int i = 0; i < p; ++i) a = fma(J[i + p * (n + j)], v[i], a);
          lu[j] -= a;
        }
      }
    }
    return nrm / (2 * rho);
  }
  // Gauss-Newton AL Hessian of one general slot from the stored z_est, the current rho and the
  // constraint Jacobian at the accepted point (knotpoint_data.cpp:549-570, :597-613):
  // G = rho (dPi J)^T (dPi J)  [+ rho J^T (d/dz (dPi^T zp)) J for the second-order cone]
  __device__ __noinline__ void al_hessian_general(const ConSlot& s, int k, long zrow, bool terminal, double* lxx,
                                    double* luu, double* lux) const {
    constexpr int nm = n + m;
    const int p = s.dim;
    double zt[kMaxConDim], c[kMaxConDim];
    double J[kMaxConDim * nm], M[kMaxConDim * nm], Gm[nm * nm];
    {
      double x[n], u[m];
      if (s.family == CON_FAMILY_DISC) {  // state-dependent Jacobian: the accepted point
        load_block<n>(F(P.x), S, k, x);
        if (!terminal) load_block<m>(F(P.u), S, k, u);
      } else {
#pragma unroll
        for (int i = 0; i < n; ++i) x[i] = 0.0;
      }
      if (terminal || s.family != CON_FAMILY_DISC) {
#pragma unroll
        for (int i = 0; i < m; ++i) u[i] = 0.0;
      }
      con_eval(s, x, u, c, J);
    }
    for (int i = 0; i < p; ++i) zt[i] = zest_read(zrow, s.row0, i);
    if (s.cone == CONE_SOC) {
      double zp[kMaxSocDim], Jp[kMaxSocDim * kMaxSocDim], Hs[kMaxSocDim * kMaxSocDim];
      soc_projection(p, zt, zp);
      soc_jacobian(p, zt, Jp);
      soc_hessian(p, zt, zp, Hs);
      for (int j = 0; j < nm; ++j)
        for (int i = 0; i < p; ++i) {
          double a = 0.0;
          for (int l = 0; l < p; ++l) a = fma(Jp[i + p * l], J[l + p * j], a);
          M[i + p * j] = a;
        }
      for (int b = 0; b < nm; ++b)
        for (int a = 0; a < nm; ++a) {
          double g = 0.0;
          for (int i = 0; i < p; ++i) g = fma(M[i + p * a], M[i + p * b], g);
          Gm[a + nm * b] = rho * g;
        }
      // + rho J^T Hs J
      for (int j = 0; j < nm; ++j)
        for (int i = 0; i < p; ++i) {
          double a = 0.0;
          for (int l = 0; l < p; ++l) a = fma(Hs[i + p * l], J[l + p * j], a);
          M[i + p * j] = a;
        }
      for (int b = 0; b < nm; ++b)
        for (int a = 0; a < nm; ++a) {
          double g = 0.0;
          for (int i = 0; i < p; ++i) g = fma(J[i + p * a], M[i + p * b], g);
          Gm[a + nm * b] += rho * g;
        }
    } else {
      for (int i = 0; i < p; ++i) {
        double act = 0.0;  // diagonal of the dual-cone projection Jacobian (cones.cpp:160-171)
        if (s.cone == CONE_EQUALITY) act = 1.0;
        if (s.cone == CONE_INEQUALITY) act = (zt[i] <= 0) ? 1.0 : 0.0;
        for (int j = 0; j < nm; ++j) M[i + p * j] = act * J[i + p * j];
      }
      for (int b = 0; b < nm; ++b)
        for (int a = 0; a < nm; ++a) {
          double g = 0.0;
          for (int i = 0; i < p; ++i) g = fma(M[i + p * a], M[i + p * b], g);
          Gm[a + nm * b] = rho * g;
        }
    }
#pragma unroll
    for (int cc = 0; cc < n; ++cc)
#pragma unroll
      for (int r = 0; r < n; ++r) lxx[r + n * cc] += Gm[r + nm * cc];
    if (!terminal) {
#pragma unroll
      for (int cc = 0; cc < m; ++cc)
#pragma unroll
        for (int r = 0; r < m; ++r) luu[r + m * cc] += Gm[(n + r) + nm * (n + cc)];
#pragma unroll
      for (int cc = 0; cc < n; ++cc)
#pragma unroll
        for (int r = 0; r < m; ++r) lux[r + m * cc] += Gm[(n + r) + nm * cc];
    }
  }

  // ---- CalcOriginalCostHessian (knotpoint_data.cpp:683-708): lxx = Q, luu = R, lux = H | 0
  ALTRO_DEV void cost_hessian(int k, bool terminal, double* lxx, double* luu, double* lux) const {
    if constexpr (CON == 2) {
      if (P.Qf) {
        const double* Q = P.Qf + (long)k * n * n;
#pragma unroll
        for (int i = 0; i < n * n; ++i) lxx[i] = Q[i];
        if (!terminal) {
          const double* R = P.Rf + (long)k * m * m;
          const double* H = P.Hf + (long)k * m * n;
#pragma unroll
          for (int i = 0; i < m * m; ++i) luu[i] = R[i];
#pragma unroll
          for (int i = 0; i < m * n; ++i) lux[i] = H[i];
        }
        return;
      }
    }
#pragma unroll
    for (int i = 0; i < n * n; ++i) lxx[i] = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) lxx[i + n * i] = Qd[k * n + i];
    if (!terminal) {
#pragma unroll
      for (int i = 0; i < m * m; ++i) luu[i] = 0.0;
#pragma unroll
      for (int i = 0; i < m; ++i) luu[i + m * i] = Rd[k * m + i];
#pragma unroll
      for (int i = 0; i < m * n; ++i) lux[i] = 0.0;
    }
  }

  // ---- augmented-Lagrangian terms of knot k (knotpoint_data.cpp:473-595).
  // Evaluates c, z_est = z - rho c (stored), returns sum ||Pi(z_est)||^2 / (2 rho) and, if
  // want_grad, subtracts J^T dPi^T Pi(z_est) from lx, lu.
  ALTRO_DEV double al_terms(int k, const double* x, const double* u, bool terminal,
                            bool want_grad, double* lx, double* lu, bool store_zest = true) {
    if constexpr (!CON) {
      return 0.0;
    } else {
      const ConTable& T = P.contab;
      double cost = 0.0;
      for (int j = 0; j < T.ncon; ++j) {
        const ConSlot& s = T.slot[j];
        if (k < s.k_start || k >= s.k_stop) continue;
        const long zrow = zoff(k, s.row0);
        bool general = false;
        if constexpr (CON == 2) {
          general = general_slot(s);
          if (general) cost += al_terms_general(s, zrow, x, u, terminal, want_grad, lx, lu, store_zest);
        }
        if (!general) {
          double nrm = 0.0;
          for (int i = 0; i < s.dim; ++i) {
            const double c = row_value(s, i, x, u);
            const double zt = zread(zrow, s.row0, i) - rho * c;
            if (store_zest) P.zest[zrow + i * 32] = zt;
            // dual cones (cones.hpp:13-30): EQUALITY -> IDENTITY, INEQUALITY -> INEQUALITY,
            // IDENTITY -> EQUALITY (projection onto {0})
            double zp = 0.0;
            if (s.cone == CONE_EQUALITY) zp = zt;
            if (s.cone == CONE_INEQUALITY) zp = fmin(0.0, zt);
            nrm += zp * zp;
            // dPi is diagonal with entries {1 | zt<=0 | 0}; dPi^T zp == zp in all three cases
            if (want_grad) scatter_sub(s.idx[i], s.scale[i] * zp, lx, lu, terminal);
          }
          cost += nrm / (2 * rho);
        }
      }
      return cost;
    }
  }

  // ---- Gauss-Newton AL Hessian of knot k from the STORED z_est and the CURRENT rho
  // (knotpoint_data.cpp:549-570, :597-613).  Adds into lxx (n x n), luu (m x m), lux (m x n).
  ALTRO_DEV void al_hessian(int k, bool terminal, double* lxx, double* luu, double* lux) const {
    if constexpr (CON) {
      const ConTable& T = P.contab;
      for (int j = 0; j < T.ncon; ++j) {
        const ConSlot& s = T.slot[j];
        if (k < s.k_start || k >= s.k_stop) continue;
        const long zrow = zoff(k, s.row0);
        bool general = false;
        if constexpr (CON == 2) {
          general = general_slot(s);
          if (general) al_hessian_general(s, k, zrow, terminal, lxx, luu, lux);
        }
        if (!general) {
          for (int i = 0; i < s.dim; ++i) {
            const int id = s.idx[i];
            if (id < 0) continue;
            const double zt = zest_read(zrow, s.row0, i);
            double act = 0.0;  // diagonal of the dual-cone projection Jacobian (cones.cpp:160-171)
            if (s.cone == CONE_EQUALITY) act = 1.0;
            if (s.cone == CONE_INEQUALITY) act = (zt <= 0) ? 1.0 : 0.0;
            const double g = rho * ((act * s.scale[i]) * (act * s.scale[i]));
#pragma unroll
            for (int e = 0; e < n; ++e) lxx[e + n * e] += (id == e) ? g : 0.0;
            if (!terminal) {
#pragma unroll
              for (int e = 0; e < m; ++e) luu[e + m * e] += (id == n + e) ? g : 0.0;
            }
          }
        }
      }
    }
  }

  // ---- max_j || Pi_K(c_j) - c_j ||_inf at knot k (knotpoint_data.cpp:489-501)
  ALTRO_DEV double al_violation(int k, const double* x, const double* u) const {
    double viol = 0.0;
    if constexpr (CON) {
      const ConTable& T = P.contab;
      for (int j = 0; j < T.ncon; ++j) {
        const ConSlot& s = T.slot[j];
        if (k < s.k_start || k >= s.k_stop) continue;
        bool general = false;
        if constexpr (CON == 2) {
          general = general_slot(s);
          if (general) viol = fmax(viol, al_violation_general(s, x, u));
        }
        if (!general) {
          for (int i = 0; i < s.dim; ++i) {
            const double c = row_value(s, i, x, u);
            double pc = 0.0;  // EQUALITY: projection onto {0}
            if (s.cone == CONE_IDENTITY) pc = c;
            if (s.cone == CONE_INEQUALITY) pc = fmin(0.0, c);
            viol = fmax(viol, fabs(pc - c));
          }
        }
      }
    }
    return viol;
  }

  // ---- DualUpdate (z <- Pi(z_est), knotpoint_data.cpp:503-510) over the whole trajectory
  ALTRO_DEV void dual_update() {
    if constexpr (CON) {
      const ConTable& T = P.contab;
      for (int j = 0; j < T.ncon; ++j) {
        const ConSlot& s = T.slot[j];
        for (int k = s.k_start; k < s.k_stop; ++k) {
          const long zrow = zoff(k, s.row0);
          if (CON == 2 && s.cone == CONE_SOC) {
            double zt[kMaxSocDim], zp[kMaxSocDim];
            for (int i = 0; i < s.dim; ++i) zt[i] = P.zest[zrow + i * 32];
            soc_projection(s.dim, zt, zp);
            for (int i = 0; i < s.dim; ++i) P.z[zrow + i * 32] = zp[i];
          } else {
            for (int i = 0; i < s.dim; ++i) {
              const double zt = P.zest[zrow + i * 32];
              double zp = 0.0;
              if (s.cone == CONE_EQUALITY) zp = zt;
              if (s.cone == CONE_INEQUALITY) zp = fmin(0.0, zt);
              P.z[zrow + i * 32] = zp;
            }
          }
        }
      }
    }
  }

  // =================================================================== sweeps
  // Initial sweep of Solve (solver.cpp:422-430): open-loop rollout, copy to the reference
  // trajectory, constraint values + projected duals with the OLD penalty, expansions and
  // gradients, then the penalty reset (quirk Q3: gradient before SetPenalty).
  ALTRO_DEV void initial_sweep() {
    double x[n], u[m], xn[n], q[n], r[m], lx[n], lu[m], A[n * n], Bm[n * m];
    load_block<n>(G(P.x0, n), 0, 0, x);
    for (int k = 0; k < N; ++k) {
      load_block<m>(F(P.u), S, k, u);
      dynamics(k, x, u, xn);
      store_block<n>(F(P.x), S, k, x);
      store_block<n>(F(P.xbar), S, k, x);
      store_block<m>(F(P.ubar), S, k, u);
      load_block<n>(F(P.q), S, k, q);
      load_block<m>(F(P.r), S, k, r);
      jacobian(k, x, u, A, Bm);
      store_jac(k, A, Bm);
      stage_gradient(k, x, u, q, r, false, lx, lu);
      al_terms(k, x, u, false, true, lx, lu);
      store_block<n>(F(P.lx), S, k, lx);
      store_block<m>(F(P.lu), S, k, lu);
#pragma unroll
      for (int i = 0; i < n; ++i) x[i] = xn[i];
    }
    store_block<n>(F(P.x), S, N, x);
    store_block<n>(F(P.xbar), S, N, x);
    load_block<n>(F(P.q), S, N, q);
#pragma unroll
    for (int i = 0; i < m; ++i) u[i] = 0.0;  // terminal u_ is a zero m-vector (quirk Q8)
    stage_gradient(N, x, u, q, r, true, lx, lu);
    al_terms(N, x, u, true, true, lx, lu);
    store_block<n>(F(P.lx), S, N, lx);
    rho = P.opts.penalty_initial;
  }

  // Backward Riccati sweep = CalcExpansions + tvlqr_BackwardPass (solver.cpp:448-449,
  // tvlqr.cpp:65-195) with reg = 0, f = 0.  The cost Hessian is rebuilt per knot from the
  // diagonal weights and the AL Gauss-Newton terms instead of being stored.  Split into a
  // terminal part and a per-knot step so that the plain sweep below (direct loads) and the
  // TMA-staged group kernel (solver_phases.cuh) run the identical arithmetic.
  ALTRO_DEV void riccati_terminal(double* Pn, double* pn) {  // tvlqr.cpp:85-90
    cost_hessian(N, true, Pn, nullptr, nullptr);
    al_hessian(N, true, Pn, nullptr, nullptr);
    load_block<n>(F(P.lx), S, N, pn);
    store_block<n * n>(F(P.P), S, N, Pn);
    store_block<n>(F(P.p), S, N, pn);
  }

  // One knot (tvlqr.cpp:92-192).  A, Bm: dynamics expansion; Qx, Qu hold lx, lu on entry; Pn, pn
  // carry the cost-to-go.  Returns false when the Cholesky of Quu fails: tvlqr returns there
  // (:162-164) and Solve ignores it (quirk Q2), so this knot keeps the unsolved K = Qux,
  // d = -Qu and P_k, p_k and everything below stay stale -- the caller must stop the sweep.
  ALTRO_DEV bool riccati_step(int k, const double* A, const double* Bm, double* Qx, double* Qu,
                              double* Pn, double* pn) {
    double Qxx[n * n], Quu[m * m], Qux[m * n];
    cost_hessian(k, false, Qxx, Quu, Qux);
    al_hessian(k, false, Qxx, Quu, Qux);
    {
      double T1[n * n];
      mm<n, n, n, true, false, 0>(A, Pn, T1);    // A' P+            tvlqr.cpp:135
      mm<n, n, n, false, false, 1>(T1, A, Qxx);  // Qxx += (A'P+) A  :136
    }
    {
      double T2[m * n];
      mm<m, n, n, true, false, 0>(Bm, Pn, T2);    // B' P+            :139
      mm<m, m, n, false, false, 1>(T2, Bm, Quu);  // Quu += (B'P+) B  :140
      mm<m, n, n, false, false, 1>(T2, A, Qux);   // Qux += (B'P+) A  :143
    }
    mm<n, 1, n, true, false, 1>(A, pn, Qx);   // Qx = q + A' p+     :147-150 (f = 0)
    mm<m, 1, n, true, false, 1>(Bm, pn, Qu);  // Qu = r + B' p+     :151-152
    double K[m * n], d[m], L[m * m];
#pragma unroll
    for (int i = 0; i < m * n; ++i) K[i] = Qux[i];
#pragma unroll
    for (int i = 0; i < m; ++i) d[i] = -Qu[i];
#pragma unroll
    for (int i = 0; i < m * m; ++i) L[i] = Quu[i];
    const bool ok = cholesky<m>(L);  // :161
    if (!ok) {
      store_block<m * n>(F(P.K), S, k, K);
      store_block<m>(F(P.d), S, k, d);
      return false;
    }
    cholesky_solve<m, n>(L, K);
    cholesky_solve<m, 1>(L, d);
    store_block<m * n>(F(P.K), S, k, K);
    store_block<m>(F(P.d), S, k, d);
    // cost-to-go, :173-186
    double QuuK[m * n], KtQux[n * n];
    mm<m, n, m, false, false, 0>(Quu, K, QuuK);
    mm<n, n, m, true, false, 0>(K, Qux, KtQux);
#pragma unroll
    for (int i = 0; i < n * n; ++i) Pn[i] = Qxx[i];
    mm<n, n, m, true, false, 1>(QuuK, K, Pn);
#pragma unroll
    for (int c = 0; c < n; ++c)
#pragma unroll
      for (int rr = 0; rr < n; ++rr) Pn[rr + n * c] -= KtQux[rr + n * c];
#pragma unroll
    for (int c = 0; c < n; ++c)
#pragma unroll
      for (int rr = 0; rr < n; ++rr) Pn[rr + n * c] -= KtQux[c + n * rr];
#pragma unroll
    for (int i = 0; i < n; ++i) pn[i] = Qx[i];
    mm<n, 1, m, true, false, -1>(QuuK, d, pn);
    mm<n, 1, m, true, false, -1>(K, Qu, pn);
    mm<n, 1, m, true, false, 1>(Qux, d, pn);
    store_block<n * n>(F(P.P), S, k, Pn);
    store_block<n>(F(P.p), S, k, pn);
    return true;
  }

  // Out-of-line copy for the large-block models (n > kUnrollDim): their arrays live in local
  // memory anyway, and keeping the step a real function keeps ptxas' register allocation of the
  // surrounding kernel sane (13 KB of spills when inlined into k_phase_backward<Chain<12,4>>).
  __device__ __noinline__ bool riccati_step_call(int k, const double* A, const double* Bm, double* Qx,
                                                 double* Qu, double* Pn, double* pn) {
    return riccati_step(k, A, Bm, Qx, Qu, Pn, pn);
  }

  ALTRO_DEV void backward_sweep() {
    double Pn[n * n], pn[n];
    riccati_terminal(Pn, pn);
    for (int k = N - 1; k >= 0; --k) {
      if (k > 0) {
        prefetch_block<kV>(F(P.A), S, k - 1);
        prefetch_block<n>(F(P.lx), S, k - 1);
        prefetch_block<m>(F(P.lu), S, k - 1);
      }
      double A[n * n], Bm[n * m], Qx[n], Qu[m];
      load_jac(k, A, Bm);
      load_block<n>(F(P.lx), S, k, Qx);
      load_block<m>(F(P.lu), S, k, Qu);
      bool ok;
      if constexpr (n > kUnrollDim)
        ok = riccati_step_call(k, A, Bm, Qx, Qu, Pn, pn);
      else
        ok = riccati_step(k, A, Bm, Qx, Qu, Pn, pn);
      if (!ok) return;
    }
  }

  // MeritFunction (solver.cpp:273-355): nonlinear closed-loop rollout at step `alpha`.
  ALTRO_DEV void merit(double alpha, bool want, double* phi_out, double* dphi_out) {
    merit_evals += 1;
    double phi = 0.0, dphi = 0.0;
    double x[n], dxda[n];
    load_block<n>(G(P.x0, n), 0, 0, x);
#pragma unroll
    for (int i = 0; i < n; ++i) dxda[i] = 0.0;
    for (int k = 0; k < N; ++k) {
      double xb[n], ub[m], K[m * n], d[m], dx[n], u[m], y[n], xn[n], q[n], r[m];
      load_block<n>(F(P.xbar), S, k, xb);
      load_block<m>(F(P.ubar), S, k, ub);
      load_block<m * n>(F(P.K), S, k, K);
      load_block<m>(F(P.d), S, k, d);
#pragma unroll
      for (int i = 0; i < n; ++i) dx[i] = x[i] - xb[i];  // :290
      {
        double Kdx[m];
        mm<m, 1, n, false, false, 0>(K, dx, Kdx);
#pragma unroll
        for (int i = 0; i < m; ++i) u[i] = ub[i] + (-Kdx[i] + alpha * d[i]);  // :291-292
      }
      {
        double Pk[n * n];
        load_block<n * n>(F(P.P), S, k, Pk);
        load_block<n>(F(P.p), S, k, y);
        mm<n, 1, n, false, false, 1>(Pk, dx, y);  // y = P dx + p, :293
      }
      store_block<n>(F(P.x), S, k, x);
      store_block<m>(F(P.u), S, k, u);
      store_block<n>(F(P.y), S, k, y);
      dynamics(k, x, u, xn);  // :296
      load_block<n>(F(P.q), S, k, q);
      load_block<m>(F(P.r), S, k, r);
      double lx[n], lu[m];
      if (want) stage_gradient(k, x, u, q, r, false, lx, lu);
      phi += stage_cost(k, x, u, q, r, false) + al_terms(k, x, u, false, want, lx, lu);  // :299-301
      if (want) {
        double A[n * n], Bm[n * m], duda[m], dxn[n];
        jacobian(k, x, u, A, Bm);  // :305
        store_jac(k, A, Bm);
        {
          double Kd[m];
          mm<m, 1, n, false, false, 0>(K, dxda, Kd);
#pragma unroll
          for (int i = 0; i < m; ++i) duda[i] = -Kd[i] + d[i];  // :306
        }
        mm<n, 1, n, false, false, 0>(A, dxda, dxn);  // :307-308
        mm<n, 1, m, false, false, 1>(Bm, duda, dxn);
        store_block<n>(F(P.lx), S, k, lx);
        store_block<m>(F(P.lu), S, k, lu);
        dphi += dot<n>(lx, dxda);  // :313-314
        dphi += dot<m>(lu, duda);
#pragma unroll
        for (int i = 0; i < n; ++i) dxda[i] = dxn[i];
      }
#pragma unroll
      for (int i = 0; i < n; ++i) x[i] = xn[i];
    }
    {  // terminal knot, :319-332
      double xb[n], dx[n], y[n], q[n], u0[m], lx[n];
      load_block<n>(F(P.xbar), S, N, xb);
      load_block<n>(F(P.q), S, N, q);
#pragma unroll
      for (int i = 0; i < m; ++i) u0[i] = 0.0;
      if (want) stage_gradient(N, x, u0, q, nullptr, true, lx, nullptr);
      phi += stage_cost(N, x, u0, q, nullptr, true) + al_terms(N, x, u0, true, want, lx, nullptr);
#pragma unroll
      for (int i = 0; i < n; ++i) dx[i] = x[i] - xb[i];
      double Pk[n * n];
      load_block<n * n>(F(P.P), S, N, Pk);
      load_block<n>(F(P.p), S, N, y);
      mm<n, 1, n, false, false, 1>(Pk, dx, y);
      store_block<n>(F(P.x), S, N, x);
      store_block<n>(F(P.y), S, N, y);
      if (want) {
        store_block<n>(F(P.lx), S, N, lx);
        dphi += dot<n>(lx, dxda);
      }
    }
    *phi_out = phi;
    if (want) *dphi_out = dphi;
  }

  // Refresh A, B, lx, lu at the working trajectory.  with_dynamics: the post-line-search fix of
  // ForwardPass for backtracking (solver.cpp:256-262).  !with_dynamics: CalcProjectedDuals +
  // CalcCostGradient after a dual/penalty update (solver.cpp:483-486).
  ALTRO_DEV void refresh_sweep(bool with_dynamics) {
    for (int k = 0; k <= N; ++k) {
      const bool terminal = (k == N);
      double x[n], u[m], q[n], r[m], lx[n], lu[m];
      load_block<n>(F(P.x), S, k, x);
      load_block<n>(F(P.q), S, k, q);
      if (!terminal) {
        load_block<m>(F(P.u), S, k, u);
        load_block<m>(F(P.r), S, k, r);
        if (with_dynamics) {
          double A[n * n], Bm[n * m];
          jacobian(k, x, u, A, Bm);
          store_jac(k, A, Bm);
        }
      } else {
#pragma unroll
        for (int i = 0; i < m; ++i) u[i] = 0.0;
      }
      stage_gradient(k, x, u, q, r, terminal, lx, lu);
      al_terms(k, x, u, terminal, true, lx, lu);
      store_block<n>(F(P.lx), S, k, lx);
      if (!terminal) store_block<m>(F(P.lu), S, k, lu);
    }
  }

  // Stationarity (solver.cpp:207-222) + Feasibility (:224-231) + CopyTrajectory (:148-157)
  ALTRO_DEV void criteria_sweep(double* stat_out, double* feas_out) {
    double res_x = 0.0, res_u = 0.0, viol = 0.0;
    double y[n];
    load_block<n>(F(P.y), S, 0, y);
    for (int k = 0; k < N; ++k) {
      double yn[n], A[n * n], Bm[n * m], lx[n], lu[m], x[n], u[m];
      load_block<n>(F(P.y), S, k + 1, yn);
      load_jac(k, A, Bm);
      load_block<n>(F(P.lx), S, k, lx);
      load_block<m>(F(P.lu), S, k, lu);
      mm<n, 1, n, true, false, 1>(A, yn, lx);  // lx + A' y+
      mm<m, 1, n, true, false, 1>(Bm, yn, lu);
#pragma unroll
      for (int i = 0; i < n; ++i) res_x = fmax(res_x, fabs(lx[i] - y[i]));
#pragma unroll
      for (int i = 0; i < m; ++i) res_u = fmax(res_u, fabs(lu[i]));
      load_block<n>(F(P.x), S, k, x);
      load_block<m>(F(P.u), S, k, u);
      viol = fmax(viol, al_violation(k, x, u));
      store_block<n>(F(P.xbar), S, k, x);
      store_block<m>(F(P.ubar), S, k, u);
#pragma unroll
      for (int i = 0; i < n; ++i) y[i] = yn[i];
    }
    {
      double lx[n], x[n], u0[m];
      load_block<n>(F(P.lx), S, N, lx);
#pragma unroll
      for (int i = 0; i < n; ++i) res_x = fmax(res_x, fabs(lx[i] - y[i]));
      load_block<n>(F(P.x), S, N, x);
#pragma unroll
      for (int i = 0; i < m; ++i) u0[i] = 0.0;
      viol = fmax(viol, al_violation(N, x, u0));
      store_block<n>(F(P.xbar), S, N, x);
    }
    *stat_out = fmax(res_x, res_u);
    *feas_out = viol;
  }

  // ForwardPass (solver.cpp:237-271)
  ALTRO_DEV int forward_pass(double* alpha_out, double* phi_final) {
    double phi0, dphi0;
    merit(0.0, true, &phi0, &dphi0);
    *phi_final = phi0;
    if (fabs(dphi0) < P.opts.tol_meritfun_gradient) {
      *alpha_out = 0.0;
      return ERR_MERIT_GRAD_TOO_SMALL;
    }
    LsOptions lo;
    lo.try_cubic_first = true;  // solver.cpp:248
    lo.use_backtracking = P.opts.use_backtracking_linesearch != 0;
    lo.c1 = P.opts.ls_c1;
    lo.c2 = P.opts.ls_c2;
    LsMachine ls;
    ls.start(lo, 1.0, phi0, dphi0);
    while (!ls.done()) {
      double ph = 0.0, dph = 0.0;
      merit(ls.alpha, ls.want_derivative(), &ph, &dph);
      ls.update(lo, ph, dph);
    }
    const double alpha = ls.alpha;
    *alpha_out = alpha;
    if (ls.n_iters > 0) *phi_final = ls.phi;
    if (lo.use_backtracking && fabs(alpha - 1.0) > 0) refresh_sweep(true);
    if (isnan(alpha) || !(ls.status == LS_MINIMUM_FOUND || ls.status == LS_HIT_MAX_STEPSIZE)) {
      return ERR_LINESEARCH_FAILED;
    }
    return ERR_NONE;
  }

  // =================================================================== phase pieces
  // The same mathematics split so that (i) work that is independent per knot (dynamics
  // Jacobians, projected duals, cost gradients, costates, residuals) runs one thread per
  // (trajectory, knot), and (ii) the sequential sweeps carry as little as possible and prefetch
  // the next knot while they work on the current one.  Used by solver_phases.cuh.

  // knot whose [q r c] rows hold knot k's linear cost terms (DeviceProblem::qrc_uniform)
  ALTRO_DEV int kq(int k) const { return (P.qrc_uniform && k < N) ? 0 : k; }

  // working trajectory of candidate `slot` (-1: the main x_, u_ arrays)
  ALTRO_DEV double* xw(int slot) const {
    return slot < 0 ? F(P.x) : P.xs + slot_off(slot);
  }
  ALTRO_DEV double* uw(int slot) const { return slot < 0 ? F(P.u) : P.us + slot_off(slot); }
  // candidate slots: [slot][group][knot][x rows | u rows][32]
  ALTRO_DEV long slot_off(int slot) const {
    return ((long)slot * P.Gtot + (b >> 5)) * (long)(N + 1) * P.Rs + (b & 31);
  }
  ALTRO_DEV long sw(int slot) const { return slot < 0 ? S : P.Rs; }

  // OpenLoopRollout + CopyTrajectory (solver.cpp:422-423): x_, x, u from the working inputs.
  ALTRO_DEV void phase_init_rollout() {
    double x[n], u[m], xn[n];
    load_block<n>(G(P.x0, n), 0, 0, x);
    for (int k = 0; k < N; ++k) {
      if (k + 1 < N) prefetch_block<m>(F(P.u), S, k + 1);
      load_block<m>(F(P.u), S, k, u);
      dynamics(k, x, u, xn);
      store_block<n>(F(P.x), S, k, x);
      store_block<n>(F(P.xbar), S, k, x);
      store_block<m>(F(P.ubar), S, k, u);
#pragma unroll
      for (int i = 0; i < n; ++i) x[i] = xn[i];
    }
    store_block<n>(F(P.x), S, N, x);
    store_block<n>(F(P.xbar), S, N, x);
  }

  // z <- Pi(z_est) for the rows of knot k (KnotPointData::DualUpdate, knotpoint_data.cpp:503-510)
  ALTRO_DEV void dual_update_knot(int k) {
    if constexpr (CON) {
      const ConTable& T = P.contab;
      for (int j = 0; j < T.ncon; ++j) {
        const ConSlot& s = T.slot[j];
        if (k < s.k_start || k >= s.k_stop) continue;
        const long zrow = zoff(k, s.row0);
        if (CON == 2 && s.cone == CONE_SOC) {
          double zt[kMaxSocDim], zp[kMaxSocDim];
          for (int i = 0; i < s.dim; ++i) zt[i] = P.zest[zrow + i * 32];
          soc_projection(s.dim, zt, zp);
          for (int i = 0; i < s.dim; ++i) P.z[zrow + i * 32] = zp[i];
        } else {
          for (int i = 0; i < s.dim; ++i) {
            const double zt = P.zest[zrow + i * 32];
            double zp = 0.0;
            if (s.cone == CONE_EQUALITY) zp = zt;
            if (s.cone == CONE_INEQUALITY) zp = fmin(0.0, zt);
            P.z[zrow + i * 32] = zp;
          }
        }
      }
    }
  }

  // L2 prefetch of what phase_expand_knot(k, ., slot, .) will load (the knot-parallel phases are
  // bound by the latency of these first loads: ncu r02e, long scoreboard)
  ALTRO_DEV void phase_expand_prefetch(int k, int slot) const {
    prefetch_block<n>(xw(slot), sw(slot), k);
    if (!P.qrc_uniform) prefetch_block<n>(F(P.q), S, k);
    if (k < N) {
      prefetch_block<m>(uw(slot), sw(slot), k);
      if (!P.qrc_uniform) prefetch_block<m>(F(P.r), S, k);
    }
  }

  // Expansion of ONE knot: [A B] (if with_dyn), projected duals, cost gradient with AL terms
  // (CalcDynamicsExpansion, CalcConstraintJacobians, CalcProjectedDuals, CalcCostGradient;
  // knotpoint_data.cpp:406-437).  The trajectory is read from candidate `slot`; when that is not
  // the main copy it is also written there (the accepted candidate becomes x_, u_).
  ALTRO_DEV void phase_expand_knot(int k, bool with_dyn, int slot, bool dual_first) {
    const bool terminal = (k == N);
    double x[n], u[m], q[n], r[m], lx[n], lu[m];
    load_block<n>(xw(slot), sw(slot), k, x);
    load_block<n>(F(P.q), S, kq(k), q);
    if (slot >= 0) store_block<n>(F(P.x), S, k, x);
    if (!terminal) {
      load_block<m>(uw(slot), sw(slot), k, u);
      load_block<m>(F(P.r), S, kq(k), r);
      if (slot >= 0) store_block<m>(F(P.u), S, k, u);
      if (with_dyn) {
        double A[n * n], Bm[n * m];
        jacobian(k, x, u, A, Bm);
        store_jac(k, A, Bm);
      }
    } else {
#pragma unroll
      for (int i = 0; i < m; ++i) u[i] = 0.0;
    }
    if (dual_first) dual_update_knot(k);
    stage_gradient(k, x, u, q, r, terminal, lx, lu);
    al_terms(k, x, u, terminal, true, lx, lu);
    store_block<n>(F(P.lx), S, k, lx);
    if (!terminal) store_block<m>(F(P.lu), S, k, lu);
  }

  // merit(0, derivative) of ForwardPass (solver.cpp:241) WITHOUT re-simulating: at alpha = 0 the
  // closed-loop rollout reproduces the accepted trajectory bit for bit (dx = 0), so x_, u_, A, B
  // are already in HBM; what changes with the new gains / duals / penalty is the cost, the
  // projected duals, the gradients and the directional derivative, recomputed here in one
  // linear scan.  phi0_step handles knot k < N given its blocks in registers.
  ALTRO_DEV void phi0_step(int k, const double* x, const double* u, const double* q, const double* r,
                           double cval, const double* K, const double* d, const double* A,
                           const double* Bm, double* dxda, double& phi, double& dphi) {
    double lx[n], lu[m], duda[m], dxn[n];
    stage_gradient(k, x, u, q, r, false, lx, lu);
    phi += stage_cost(k, x, u, q, r, false, cval) + al_terms(k, x, u, false, true, lx, lu);
    store_block<n>(F(P.lx), S, k, lx);
    store_block<m>(F(P.lu), S, k, lu);
    {
      double Kd[m];
      mm<m, 1, n, false, false, 0>(K, dxda, Kd);
#pragma unroll
      for (int i = 0; i < m; ++i) duda[i] = -Kd[i] + d[i];
    }
    mm<n, 1, n, false, false, 0>(A, dxda, dxn);
    mm<n, 1, m, false, false, 1>(Bm, duda, dxn);
    dphi += dot<n>(lx, dxda);
    dphi += dot<m>(lu, duda);
#pragma unroll
    for (int i = 0; i < n; ++i) dxda[i] = dxn[i];
  }
  ALTRO_DEV void phi0_terminal(const double* dxda, double& phi, double& dphi) {
    double x[n], q[n], u0[m], lx[n];
    load_block<n>(F(P.x), S, N, x);
    load_block<n>(F(P.q), S, N, q);
#pragma unroll
    for (int i = 0; i < m; ++i) u0[i] = 0.0;
    stage_gradient(N, x, u0, q, nullptr, true, lx, nullptr);
    phi += stage_cost(N, x, u0, q, nullptr, true) + al_terms(N, x, u0, true, true, lx, nullptr);
    store_block<n>(F(P.lx), S, N, lx);
    dphi += dot<n>(lx, dxda);
  }
  ALTRO_DEV void phase_phi0_scan(double* phi_out, double* dphi_out) {
    double phi = 0.0, dphi = 0.0;
    double dxda[n];
#pragma unroll
    for (int i = 0; i < n; ++i) dxda[i] = 0.0;
    for (int k = 0; k < N; ++k) {
      {
        const int kn = k + 1;
        prefetch_block<n>(F(P.x), S, kn);
        prefetch_block<n>(F(P.q), S, kn);
        if (kn < N) {
          prefetch_block<m>(F(P.u), S, kn);
          prefetch_block<m>(F(P.r), S, kn);
          prefetch_block<m * n>(F(P.K), S, kn);
          prefetch_block<m>(F(P.d), S, kn);
          prefetch_block<kV>(F(P.A), S, kn);
        }
      }
      double x[n], u[m], q[n], r[m], K[m * n], d[m], A[n * n], Bm[n * m];
      load_block<n>(F(P.x), S, k, x);
      load_block<m>(F(P.u), S, k, u);
      load_block<n>(F(P.q), S, k, q);
      load_block<m>(F(P.r), S, k, r);
      load_block<m * n>(F(P.K), S, k, K);
      load_block<m>(F(P.d), S, k, d);
      load_jac(k, A, Bm);
      phi0_step(k, x, u, q, r, F(P.c)[(long)k * S], K, d, A, Bm, dxda, phi, dphi);
    }
    phi0_terminal(dxda, phi, dphi);
    *phi_out = phi;
    *dphi_out = dphi;
  }

  // The simulation half of MeritFunction (solver.cpp:285-301, :319-322): closed-loop rollout at
  // step alpha, returning the merit value.  Derivative work is left to phase_expand_knot
  // (parallel over knots) + the d(phi) scan; y_ is produced for the accepted point only
  // (post_chunk).  z_est is not stored here: every candidate that can be accepted is
  // expanded afterwards, which stores it.  rollout_step advances x from knot k to k + 1; xo/uo
  // (may be null: merit value only) receive the trial trajectory with knot stride `so`.
  // the control of knot k on the closed-loop rollout at step alpha (solver.cpp:287-291)
  ALTRO_DEV void rollout_control(double alpha, const double* xb, const double* ub, const double* K,
                                 const double* d, const double* x, double* u) const {
    double dx[n];
#pragma unroll
    for (int i = 0; i < n; ++i) dx[i] = x[i] - xb[i];
    double Kdx[m];
    mm<m, 1, n, false, false, 0>(K, dx, Kdx);
#pragma unroll
    for (int i = 0; i < m; ++i) u[i] = ub[i] + (-Kdx[i] + alpha * d[i]);
  }
  // ... and the rest of the step: stores, x_{k+1}, the knot's share of the merit value
  ALTRO_DEV void rollout_advance(int k, const double* u, const double* q, const double* r, double cval,
                                 double* x, double* xo, double* uo, long so, double& phi) {
    double xn[n];
    if (xo) {
      store_block<n>(xo, so, k, x);
      store_block<m>(uo, so, k, u);
    }
    dynamics(k, x, u, xn);
    phi += stage_cost(k, x, u, q, r, false, cval) +
           al_terms(k, x, u, false, false, nullptr, nullptr, false);
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = xn[i];
  }
  ALTRO_DEV void rollout_step(int k, double alpha, const double* xb, const double* ub,
                              const double* K, const double* d, const double* q, const double* r,
                              double cval, double* x, double* xo, double* uo, long so, double& phi) {
    double u[m];
    rollout_control(alpha, xb, ub, K, d, x, u);
    rollout_advance(k, u, q, r, cval, x, xo, uo, so, phi);
  }
  // The derivative half of MeritFunction (solver.cpp:303-315) for knot k of a trial trajectory, by
  // the FOLLOWER warp of k_phase_forward right behind the rollout warp, which hands it x_k, u_k
  // through shared memory: [A B] (stored), cost gradient with AL terms (stored, z_est stored) and
  // the phi' recurrence -- the trial point's expansion and d(phi) scan without re-reading x, u, [J],
  // lx, lu from HBM and off the critical path of the state recursion.
  ALTRO_DEV void follow_step(int k, const double* x, const double* u, const double* q, const double* r,
                             const double* K, const double* d, double* dxda, double& dphi) {
    double A[n * n], Bm[n * m], lx[n], lu[m];
    jacobian(k, x, u, A, Bm);
    {
      double J[kV];  // through the packed form, like every other reader of [A B]
      JP::pack(A, Bm, J);
      store_block<kV>(P.A + go, S, k, J);
      JP::unpack(J, hof(k), A, Bm);
    }
    stage_gradient(k, x, u, q, r, false, lx, lu);
    al_terms(k, x, u, false, true, lx, lu);
    store_block<n>(F(P.lx), S, k, lx);
    store_block<m>(F(P.lu), S, k, lu);
    dphi_step(K, d, A, Bm, lx, lu, dxda, dphi);
  }
  // terminal knot of the same (solver.cpp:327-331)
  ALTRO_DEV void follow_terminal(const double* x, const double* dxda, double& dphi) {
    double q[n], u0[m], lx[n];
    load_block<n>(F(P.q), S, N, q);
#pragma unroll
    for (int i = 0; i < m; ++i) u0[i] = 0.0;
    stage_gradient(N, x, u0, q, nullptr, true, lx, nullptr);
    al_terms(N, x, u0, true, true, lx, nullptr);
    store_block<n>(F(P.lx), S, N, lx);
    dphi += dot<n>(lx, dxda);
  }
  ALTRO_DEV void rollout_terminal(const double* x, double* xo, long so, double& phi) {
    double q[n], u0[m];
    load_block<n>(F(P.q), S, N, q);
#pragma unroll
    for (int i = 0; i < m; ++i) u0[i] = 0.0;
    phi += stage_cost(N, x, u0, q, nullptr, true) + al_terms(N, x, u0, true, false, nullptr, nullptr, false);
    if (xo) store_block<n>(xo, so, N, x);
  }
  ALTRO_DEV double phase_rollout(double alpha, double* xo, double* uo, long so) {
    double phi = 0.0;
    double x[n];
    load_block<n>(G(P.x0, n), 0, 0, x);
    for (int k = 0; k < N; ++k) {
      {
        const int kn = k + 1;
        prefetch_block<n>(F(P.xbar), S, kn);
        if (!P.qrc_uniform || kn == N) prefetch_block<n>(F(P.q), S, kn);
        if (kn < N) {
          prefetch_block<m>(F(P.ubar), S, kn);
          prefetch_block<m * n>(F(P.K), S, kn);
          prefetch_block<m>(F(P.d), S, kn);
          if (!P.qrc_uniform) prefetch_block<m>(F(P.r), S, kn);
        }
      }
      double xb[n], ub[m], K[m * n], d[m], q[n], r[m];
      load_block<n>(F(P.xbar), S, k, xb);
      load_block<m>(F(P.ubar), S, k, ub);
      load_block<m * n>(F(P.K), S, k, K);
      load_block<m>(F(P.d), S, k, d);
      load_block<n>(F(P.q), S, kq(k), q);
      load_block<m>(F(P.r), S, kq(k), r);
      rollout_step(k, alpha, xb, ub, K, d, q, r, F(P.c)[(long)kq(k) * S], x, xo, uo, so, phi);
    }
    rollout_terminal(x, xo, so, phi);
    return phi;
  }

  // The derivative half of MeritFunction (solver.cpp:303-315, :327-331) once A, B, lx, lu of the
  // trial trajectory are in HBM.
  ALTRO_DEV void dphi_step(const double* K, const double* d, const double* A, const double* Bm,
                           const double* lx, const double* lu, double* dxda, double& dphi) {
    double duda[m], dxn[n];
    {
      double Kd[m];
      mm<m, 1, n, false, false, 0>(K, dxda, Kd);
#pragma unroll
      for (int i = 0; i < m; ++i) duda[i] = -Kd[i] + d[i];
    }
    mm<n, 1, n, false, false, 0>(A, dxda, dxn);
    mm<n, 1, m, false, false, 1>(Bm, duda, dxn);
    dphi += dot<n>(lx, dxda);
    dphi += dot<m>(lu, duda);
#pragma unroll
    for (int i = 0; i < n; ++i) dxda[i] = dxn[i];
  }
  ALTRO_DEV double dphi_terminal(const double* dxda, double dphi) {
    double lx[n];
    load_block<n>(F(P.lx), S, N, lx);
    dphi += dot<n>(lx, dxda);
    return dphi;
  }
  ALTRO_DEV double phase_dphi_scan() {
    double dphi = 0.0;
    double dxda[n];
#pragma unroll
    for (int i = 0; i < n; ++i) dxda[i] = 0.0;
    for (int k = 0; k < N; ++k) {
      {
        const int kn = k + 1;
        prefetch_block<n>(F(P.lx), S, kn);
        if (kn < N) {
          prefetch_block<m * n>(F(P.K), S, kn);
          prefetch_block<m>(F(P.d), S, kn);
          prefetch_block<kV>(F(P.A), S, kn);
          prefetch_block<m>(F(P.lu), S, kn);
        }
      }
      double K[m * n], d[m], A[n * n], Bm[n * m], lx[n], lu[m];
      load_block<m * n>(F(P.K), S, k, K);
      load_block<m>(F(P.d), S, k, d);
      load_jac(k, A, Bm);
      load_block<n>(F(P.lx), S, k, lx);
      load_block<m>(F(P.lu), S, k, lu);
      dphi_step(K, d, A, Bm, lx, lu, dxda, dphi);
    }
    return dphi_terminal(dxda, dphi);
  }

  // ---- everything between the line search and the convergence decision, ONE pass over the knots:
  // the post-search expansion of an accepted backtracking step (solver.cpp:256-262, `refresh`; the
  // accepted candidate may still sit in candidate slot `slot`), the costates y_k (:293, :324),
  // Stationarity (:207-222), Feasibility (:224-231) and CopyTrajectory (:148-157).  A thread walks
  // its chunk of knots [k0, k1) DOWNWARDS carrying y_{k+1} in registers, so every block of the
  // record is read once and y, [A B], lx, lu never make the round trip through HBM that the
  // separate expand / costate / residual passes paid for.  post_boundary computes y_{k1} (from the
  // old xbar_{k1}, which the neighbouring chunk overwrites at its very end: the caller puts a CTA
  // barrier between post_boundary and post_chunk).
  ALTRO_DEV void costate_of(int k, const double* x, double* y) const {
    double xb[n], dx[n], Pk[n * n];
    load_block<n>(F(P.xbar), S, k, xb);
#pragma unroll
    for (int i = 0; i < n; ++i) dx[i] = x[i] - xb[i];
    load_block<n * n>(F(P.P), S, k, Pk);
    load_block<n>(F(P.p), S, k, y);
    mm<n, 1, n, false, false, 1>(Pk, dx, y);
  }
  ALTRO_DEV void post_boundary(int k, int slot, double* y) const {
    double x[n];
    load_block<n>(xw(slot), sw(slot), k, x);
    costate_of(k, x, y);
  }
  // L2 prefetch of what post_chunk reads at knot k, spread over the lanes of the warp (128-byte
  // lines of the group's record: [xbar ubar q r c], [x u J], [lx lu], [P p])
  ALTRO_DEV void post_prefetch(int k) const {
    const char* rec = reinterpret_cast<const char*>(P.xbar + (long)(b >> 5) * P.GS + (long)k * S);
    const int lane = b & 31;
    auto range = [&](int row0, int rows) {
      for (int i = lane; i < rows * 2; i += 32)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + (long)row0 * 256 + i * 128));
    };
    range(rXbar, (P.qrc_uniform ? rQ : rK) - rXbar);
    range(rX, rA + kV - rX);
    range(rLx, rY - rLx);
    range(rP, rUinit - rP);
  }
  ALTRO_DEV void post_chunk(int k0, int k1, bool refresh, int slot, double* yn) {
    double res = 0.0, viol = 0.0;
    for (int k = k1 - 1; k >= k0; --k) {
      if (k > k0) post_prefetch(k - 1);
      const bool terminal = (k == N);
      double x[n], u[m], lx[n], lu[m], y[n], A[n * n], Bm[n * m];
      load_block<n>(xw(slot), sw(slot), k, x);
      if (!terminal) {
        load_block<m>(uw(slot), sw(slot), k, u);
      } else {
#pragma unroll
        for (int i = 0; i < m; ++i) u[i] = 0.0;
      }
      if (refresh) {
        double q[n], r[m];
        load_block<n>(F(P.q), S, kq(k), q);
        if (slot >= 0) store_block<n>(F(P.x), S, k, x);
        if (!terminal) {
          load_block<m>(F(P.r), S, kq(k), r);
          if (slot >= 0) store_block<m>(F(P.u), S, k, u);
          jacobian(k, x, u, A, Bm);
          // through the packed form, like every other reader of [A B]
          double J[kV];
          JP::pack(A, Bm, J);
          store_block<kV>(P.A + go, S, k, J);
          JP::unpack(J, hof(k), A, Bm);
        }
        stage_gradient(k, x, u, q, r, terminal, lx, lu);
        al_terms(k, x, u, terminal, true, lx, lu);
        store_block<n>(F(P.lx), S, k, lx);
        if (!terminal) store_block<m>(F(P.lu), S, k, lu);
      } else {
        load_block<n>(F(P.lx), S, k, lx);
        if (!terminal) {
          load_block<m>(F(P.lu), S, k, lu);
          load_jac(k, A, Bm);
        }
      }
      costate_of(k, x, y);
      store_block<n>(F(P.y), S, k, y);
      if (!terminal) {
        mm<n, 1, n, true, false, 1>(A, yn, lx);  // lx + A' y+
        mm<m, 1, n, true, false, 1>(Bm, yn, lu);
#pragma unroll
        for (int i = 0; i < m; ++i) res = fmax(res, fabs(lu[i]));
        store_block<m>(F(P.ubar), S, k, u);
      }
#pragma unroll
      for (int i = 0; i < n; ++i) res = fmax(res, fabs(lx[i] - y[i]));
      viol = fmax(viol, al_violation(k, x, u));
      store_block<n>(F(P.xbar), S, k, x);
#pragma unroll
      for (int i = 0; i < n; ++i) yn[i] = y[i];
    }
    if (res > 0.0) atomicMax(P.stat_acc + b, (unsigned long long)__double_as_longlong(res));
    if (CON && viol > 0.0) atomicMax(P.feas_acc + b, (unsigned long long)__double_as_longlong(viol));
  }

  // SolverImpl::Solve (solver.cpp:414-511)
  ALTRO_DEV void solve() {
    const DevOptions& o = P.opts;
    rho = CON ? P.rho[b] : 1.0;  // penalties persist between solves (MPC warm start)
    initial_sweep();
    bool is_converged = false, stop_iterating = false;
    int status = SOLVE_UNSOLVED;
    double stationarity = 0.0, feasibility = 0.0, phi = 0.0;
    int ls_fail = 0;
    int iter;
    for (iter = 0; iter < o.iterations_max; ++iter) {
      backward_sweep();
      double alpha;
      const int err = forward_pass(&alpha, &phi);
      if (!(err == ERR_NONE || err == ERR_MERIT_GRAD_TOO_SMALL)) {
        stop_iterating = true;
        ls_fail = 1;
      }
      criteria_sweep(&stationarity, &feasibility);
      if (fabs(stationarity) < o.tol_stationarity && feasibility < o.tol_primal_feasibility) {
        is_converged = true;
        stop_iterating = true;
        status = SOLVE_SUCCESS;
      }
      if (stationarity < sqrt(o.tol_stationarity)) {  // :474-489
        if constexpr (CON) {
          dual_update();
          if (feasibility > o.tol_primal_feasibility)
            rho = fmin(rho * o.penalty_scaling, o.penalty_max);
        }
        refresh_sweep(false);
      }
      if (stop_iterating) break;
    }
    if (!is_converged && iter == o.iterations_max) status = SOLVE_MAX_ITERATIONS;
    P.status[b] = status;
    P.iters[b] = iter + 1;  // quirk Q4
    P.merit_evals[b] = merit_evals;
    P.ls_fail[b] = ls_fail;
    P.phi[b] = phi;
    P.stat[b] = stationarity;
    P.feas[b] = feasibility;
    if constexpr (CON) P.rho[b] = rho;
  }
};

template <class Model, int CON>
__global__ void __launch_bounds__(32) solve_kernel(const __grid_constant__ DeviceProblem P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  TrajSolver<Model, CON> s(P, b);
  s.solve();
}

}  // namespace altro_b200
