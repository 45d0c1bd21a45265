/*
 * altro_oracle.c -- ORACLE (test infrastructure only, see altro_oracle.h).
 *
 * Plain-C, single-problem restatement of the reference's AL-iLQR solve path.
 *   tvlqr                 src/tvlqr/tvlqr.cpp:18-248
 *   cones                 src/altro/solver/cones.cpp:13-202, cones.hpp:13-49
 *   knot-point math       src/altro/solver/knotpoint_data.cpp:229-719
 *   solver loop           src/altro/solver/solver.cpp:116-511
 *   LQR cost / MPC helpers src/altro/altro_solver.cpp:138-172, 266-293
 * All matrices column-major, FP64, `float` time step (typedefs.hpp:31-35).
 * State/input dimensions are uniform over the horizon (every reference test is).
 */
#include "altro_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ============================================================== small dense helpers */
static double *dalloc(int n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }
static void vcopy(int n, const double *a, double *b) { memcpy(b, a, sizeof(double) * (size_t)n); }
static void vzero(int n, double *a) { memset(a, 0, sizeof(double) * (size_t)n); }
static double vdot(int n, const double *a, const double *b) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
static double vinfnorm(int n, const double *a) {
  double s = 0;
  for (int i = 0; i < n; ++i) s = fmax(s, fabs(a[i]));
  return s;
}
/* C(ra x cb) (+)= op(A) * op(B); all column-major. ta/tb: transpose flags.
 * acc: 0 assign, +1 add, -1 subtract. A is (ta ? k x ra : ra x k), B is (tb ? cb x k : k x cb) */
static void gemm(int ra, int cb, int k, const double *A, int ta, const double *B, int tb, double *C,
                 int acc) {
  int lda = ta ? k : ra;
  int ldb = tb ? cb : k;
  for (int j = 0; j < cb; ++j)
    for (int i = 0; i < ra; ++i) {
      double s = 0;
      for (int l = 0; l < k; ++l) {
        double a = ta ? A[l + lda * i] : A[i + lda * l];
        double b = tb ? B[j + ldb * l] : B[l + ldb * j];
        s += a * b;
      }
      if (acc == 0)
        C[i + ra * j] = s;
      else if (acc > 0)
        C[i + ra * j] += s;
      else
        C[i + ra * j] -= s;
    }
}

/* In-place lower Cholesky of the m x m matrix M (column-major).  Mirrors Eigen's unblocked
 * LLT (used by tvlqr.cpp:161): fails when a pivot is <= 0.  Returns 0 on success. */
static int chol_lower(int m, double *M) {
  for (int j = 0; j < m; ++j) {
    double x = M[j + m * j];
    for (int l = 0; l < j; ++l) x -= M[j + m * l] * M[j + m * l];
    if (x <= 0.0) return 1;
    x = sqrt(x);
    M[j + m * j] = x;
    for (int i = j + 1; i < m; ++i) {
      double s = M[i + m * j];
      for (int l = 0; l < j; ++l) s -= M[i + m * l] * M[j + m * l];
      M[i + m * j] = s / x;
    }
  }
  return 0;
}
/* X <- (L L^T)^-1 X for X m x c */
static void chol_solve(int m, const double *L, int c, double *X) {
  for (int col = 0; col < c; ++col) {
    double *x = X + m * col;
    for (int i = 0; i < m; ++i) {
      double s = x[i];
      for (int l = 0; l < i; ++l) s -= L[i + m * l] * x[l];
      x[i] = s / L[i + m * i];
    }
    for (int i = m - 1; i >= 0; --i) {
      double s = x[i];
      for (int l = i + 1; l < m; ++l) s -= L[l + m * i] * x[l];
      x[i] = s / L[i + m * i];
    }
  }
}

void oracle_default_options(oracle_options *o) { /* solver_options.hpp:18-37 */
  o->iterations_max = 200;
  o->tol_primal_feasibility = 1e-4;
  o->tol_stationarity = 1e-4;
  o->tol_meritfun_gradient = 1e-8;
  o->penalty_initial = 1.0;
  o->penalty_scaling = 10.0;
  o->penalty_max = 1e8;
  o->use_backtracking_linesearch = 0;
  o->ls_c1 = 1e-4;
  o->ls_c2 = 0.9;
}

/* ============================================================== cones (cones.cpp) */
int oracle_dual_cone(int cone) { /* cones.hpp:13-30 */
  switch (cone) {
    case ORACLE_EQUALITY:
      return ORACLE_IDENTITY;
    case ORACLE_INEQUALITY:
      return ORACLE_INEQUALITY;
    case ORACLE_SOC:
      return ORACLE_SOC;
    case ORACLE_IDENTITY:
      return ORACLE_EQUALITY;
  }
  return ORACLE_IDENTITY;
}
static int cone_projection_is_linear(int cone) { return cone != ORACLE_SOC; } /* cones.hpp:32-49 */

static void soc_projection(int dim, const double *x, double *px) { /* cones.cpp:13-39 */
  int n = dim - 1;
  double s = x[n], a = 0.0;
  for (int i = 0; i < n; ++i) a += x[i] * x[i];
  a = sqrt(a);
  if (a <= -s) {
    for (int i = 0; i < dim; ++i) px[i] = 0.0;
  } else if (a <= s) {
    for (int i = 0; i < dim; ++i) px[i] = x[i];
  } else {
    double c = 0.5 * (1 + s / a);
    for (int i = 0; i < n; ++i) px[i] = c * x[i];
    px[n] = c * a;
  }
}

static void soc_jacobian(int dim, const double *x, double *J) { /* cones.cpp:41-77 */
  int n = dim - 1;
  double s = x[n], a = 0.0;
  for (int i = 0; i < n; ++i) a += x[i] * x[i];
  a = sqrt(a);
  vzero(dim * dim, J);
  if (a <= -s) {
    return;
  } else if (a <= s) {
    for (int i = 0; i < dim; ++i) J[i + dim * i] = 1.0;
  } else {
    double c = 0.5 * (1 + s / a);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        J[i + dim * j] = -0.5 * s / (a * a * a) * x[i] * x[j];
        J[i + dim * j] += (i == j) ? c : 0;
      }
    for (int i = 0; i < n; ++i) J[i + dim * n] = 0.5 * x[i] / a;
    for (int j = 0; j < n; ++j) J[n + dim * j] = ((-0.5 * s / (a * a)) + c / a) * x[j];
    J[n + dim * n] = 0.5;
  }
}

static void soc_hessian(int dim, const double *x, const double *b, double *H) { /* :79-123 */
  int n = dim - 1;
  double s = x[n], bs = b[n], vbv = 0, a = 0;
  for (int i = 0; i < n; ++i) {
    a += x[i] * x[i];
    vbv += x[i] * b[i];
  }
  a = sqrt(a);
  vzero(dim * dim, H);
  if (a <= -s || a <= s) return;
  for (int i = 0; i < n; ++i) {
    double hi = 0;
    for (int j = 0; j < n; ++j) {
      double Hij = -x[i] * x[j] / (a * a);
      Hij += (i == j) ? 1 : 0;
      hi += Hij * b[j];
    }
    H[i + dim * n] = hi / (2 * a);
    H[n + dim * i] = hi / (2 * a);
    for (int j = 0; j <= i; ++j) {
      double vij = x[i] * x[j];
      double H1 = hi * x[j] * (-s / (a * a * a));
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
  H[n + dim * n] = 0.0;
}

void oracle_conic_projection(int cone, int dim, const double *x, double *px) { /* :125-150 */
  switch (cone) {
    case ORACLE_EQUALITY:
      for (int i = 0; i < dim; ++i) px[i] = 0;
      break;
    case ORACLE_IDENTITY:
      for (int i = 0; i < dim; ++i) px[i] = x[i];
      break;
    case ORACLE_INEQUALITY:
      for (int i = 0; i < dim; ++i) px[i] = fmin(0.0, x[i]);
      break;
    case ORACLE_SOC:
      soc_projection(dim, x, px);
      break;
  }
}

void oracle_conic_projection_jacobian(int cone, int dim, const double *x, double *J) { /* :152-177 */
  switch (cone) {
    case ORACLE_EQUALITY:
      vzero(dim * dim, J);
      break;
    case ORACLE_IDENTITY:
      vzero(dim * dim, J);
      for (int i = 0; i < dim; ++i) J[i + dim * i] = 1.0;
      break;
    case ORACLE_INEQUALITY:
      vzero(dim * dim, J);
      for (int i = 0; i < dim; ++i) J[i + dim * i] = (x[i] <= 0) ? 1 : 0;
      break;
    case ORACLE_SOC:
      soc_jacobian(dim, x, J);
      break;
  }
}

void oracle_conic_projection_hessian(int cone, int dim, const double *x, const double *b,
                                     double *H) { /* :179-202 */
  if (cone == ORACLE_SOC)
    soc_hessian(dim, x, b, H);
  else
    vzero(dim * dim, H);
}

/* ============================================================== tvlqr (tvlqr.cpp) */
int oracle_tvlqr_total_mem_size(const int *nx, const int *nu, int N, int is_diag) { /* :18-63 */
  if (!nx) return 0;
  if (!nu) return 0;
  int mem = 0;
  for (int k = 0; k <= N; ++k) {
    int n = nx[k];
    mem += is_diag ? n : n * n; /* Q */
    mem += n + n * n + n + n + n; /* q P p x y */
    if (k < N) {
      int m = nu[k];
      mem += n * n + n * m + n;        /* A B f */
      mem += is_diag ? m : m * m;      /* R */
      mem += is_diag ? 0 : m * n;      /* H */
      mem += m;                        /* r */
      mem += m * n + m;                /* K d */
      mem += 2 * (n * n + m * m + m * n + n + m); /* Qxx..Qu and _tmp */
      mem += m;                        /* u */
    }
  }
  mem += 2;
  return mem * (int)sizeof(double);
}

int oracle_tvlqr_backward_pass(const int *nx, const int *nu, int N, const double *const *A,
                               const double *const *B, const double *const *f,
                               const double *const *Q, const double *const *R,
                               const double *const *H, const double *const *q,
                               const double *const *r, double reg, double **K, double **d,
                               double **P, double **p, double *delta_V, double **Qxx, double **Quu,
                               double **Qux, double **Qx, double **Qu, double **Qxx_tmp,
                               double **Quu_tmp, double **Qux_tmp, double **Qx_tmp,
                               double **Qu_tmp, int linear_only_update, int is_diag) {
  (void)linear_only_update; /* tvlqr.cpp:78 */
  {                         /* terminal cost-to-go, :82-90 */
    int n = nx[N];
    delta_V[0] = 0;
    delta_V[1] = 0;
    if (is_diag) {
      vzero(n * n, P[N]);
      for (int i = 0; i < n; ++i) P[N][i + n * i] = Q[N][i];
    } else {
      vcopy(n * n, Q[N], P[N]);
    }
    vcopy(n, q[N], p[N]);
  }
  for (int k = N - 1; k >= 0; --k) { /* :92-192 */
    int n = nx[k], m = nu[k], n2 = nx[k + 1];
    const double *Pn = P[k + 1], *pn = p[k + 1];
    double *Qxx_k = Qxx[k], *Quu_k = Quu[k], *Qux_k = Qux[k], *Qx_k = Qx[k], *Qu_k = Qu[k];
    double *Qxx_ = Qxx_tmp[k], *Quu_ = Quu_tmp[k], *Qux_ = Qux_tmp[k], *Qx_ = Qx_tmp[k],
           *Qu_ = Qu_tmp[k];
    const double *A_k = A[k], *B_k = B[k], *f_k = f[k];
    if (is_diag) { /* :125-128 */
      vzero(n * n, Qxx_k);
      for (int i = 0; i < n; ++i) Qxx_k[i + n * i] = Q[k][i];
      vzero(m * m, Quu_k);
      for (int i = 0; i < m; ++i) Quu_k[i + m * i] = R[k][i];
      vzero(m * n, Qux_k);
    } else { /* :129-133 */
      vcopy(n * n, Q[k], Qxx_k);
      vcopy(m * m, R[k], Quu_k);
      vcopy(m * n, H[k], Qux_k);
    }
    gemm(n, n2, n2, A_k, 1, Pn, 0, Qxx_, 0);   /* Qxx_ = A' P+          :135 */
    gemm(n, n, n2, Qxx_, 0, A_k, 0, Qxx_k, 1); /* Qxx += Qxx_ A         :136 */
    gemm(m, n2, n2, B_k, 1, Pn, 0, Qux_, 0);   /* Qux_ = B' P+          :139 */
    gemm(m, m, n2, Qux_, 0, B_k, 0, Quu_k, 1); /* Quu += Qux_ B         :140 */
    gemm(m, n, n2, Qux_, 0, A_k, 0, Qux_k, 1); /* Qux += Qux_ A         :143 */
    vcopy(n2, pn, Qx_);                        /* Qx_ = p+ + P+ f       :147-148 */
    gemm(n2, 1, n2, Pn, 0, f_k, 0, Qx_, 1);
    vcopy(n, q[k], Qx_k);                      /* Qx = q + A' Qx_       :149-150 */
    gemm(n, 1, n2, A_k, 1, Qx_, 0, Qx_k, 1);
    vcopy(m, r[k], Qu_k);                      /* Qu = r + B' Qx_       :151-152 */
    gemm(m, 1, n2, B_k, 1, Qx_, 0, Qu_k, 1);

    double *K_k = K[k], *d_k = d[k];
    vcopy(m * n, Qux_k, K_k); /* :157 */
    for (int i = 0; i < m; ++i) d_k[i] = -Qu_k[i];
    vcopy(m * m, Quu_k, Quu_);
    for (int i = 0; i < m; ++i) Quu_[i + m * i] += reg;
    if (chol_lower(m, Quu_)) return k; /* :161-164 */
    chol_solve(m, Quu_, n, K_k);
    chol_solve(m, Quu_, 1, d_k);

    double *P_k = P[k], *p_k = p[k];
    vcopy(n * n, Qxx_k, P_k);                  /* :173 */
    gemm(m, n, m, Quu_k, 0, K_k, 0, Qux_, 0);  /* Qux_ = Quu K          :174 */
    gemm(n, n, m, K_k, 1, Qux_k, 0, Qxx_, 0);  /* Qxx_ = K' Qux         :175 */
    gemm(n, 1, m, K_k, 1, Qu_k, 0, Qx_, 0);    /* Qx_  = K' Qu          :176 */
    gemm(n, n, m, Qux_, 1, K_k, 0, P_k, 1);    /* P += (Quu K)' K       :177 */
    for (int j = 0; j < n; ++j)                /* P -= Qxx_ ; P -= Qxx_':178-179 */
      for (int i = 0; i < n; ++i) P_k[i + n * j] -= Qxx_[i + n * j];
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) P_k[i + n * j] -= Qxx_[j + n * i];
    vcopy(n, Qx_k, p_k);                       /* :183 */
    gemm(n, 1, m, Qux_, 1, d_k, 0, p_k, -1);   /* p -= (Quu K)' d       :184 */
    gemm(n, 1, m, K_k, 1, Qu_k, 0, p_k, -1);   /* p -= K' Qu            :185 */
    gemm(n, 1, m, Qux_k, 1, d_k, 0, p_k, 1);   /* p += Qux' d           :186 */
    gemm(m, 1, m, Quu_k, 0, d_k, 0, Qu_, 0);   /* :189-191 */
    delta_V[0] += vdot(m, d_k, Qu_k);
    delta_V[1] += 0.5 * vdot(m, d_k, Qu_);
  }
  return -1; /* TVLQR_SUCCESS, tvlqr.h:11 */
}

int oracle_tvlqr_forward_pass(const int *nx, const int *nu, int N, const double *const *A,
                              const double *const *B, const double *const *f,
                              const double *const *K, const double *const *d,
                              const double *const *P, const double *const *p, const double *x0,
                              double **x, double **u, double **y) { /* tvlqr.cpp:197-248 */
  vcopy(nx[0], x0, x[0]);
  for (int k = 0; k < N; ++k) {
    int n = nx[k], m = nu[k], n2 = nx[k + 1];
    vcopy(m, d[k], u[k]);
    gemm(m, 1, n, K[k], 0, x[k], 0, u[k], -1);
    vcopy(n2, f[k], x[k + 1]);
    gemm(n2, 1, n, A[k], 0, x[k], 0, x[k + 1], 1);
    gemm(n2, 1, m, B[k], 0, u[k], 0, x[k + 1], 1);
    if (y) {
      gemm(n, 1, n, P[k], 0, x[k], 0, y[k], 0);
      for (int i = 0; i < n; ++i) y[k][i] += p[k][i];
    }
  }
  if (y) {
    int n = nx[N];
    gemm(n, 1, n, P[N], 0, x[N], 0, y[N], 0);
    for (int i = 0; i < n; ++i) y[N][i] += p[N][i];
  }
  return -1;
}

/* ============================================================== knot point (knotpoint_data.*) */
typedef struct {
  int cone, dim;
  oracle_con_fn con, jac; /* callbacks, or NULL for the built-in selector-affine rows */
  void *ud;
  int idx[ORACLE_MAX_CON_DIM];
  double scale[ORACLE_MAX_CON_DIM], off[ORACLE_MAX_CON_DIM];
  /* knotpoint_data.hpp:187-198 */
  double *constraint_val_, *constraint_jac_, *constraint_hess_, *v_, *z_, *z_est_, *z_proj_,
      *proj_jvp_, *proj_jac_, *proj_hess_, *jac_tmp_;
  double rho_;
} con_t;

typedef struct {
  int index, is_terminal, n, m;
  float h;
  int cost_type;
  double *Q_, *R_, *H_, *q_, *r_;
  double c_;
  int dynamics_are_linear, dynamics_is_set, cost_is_set;
  double *affine_term_;
  int ncon;
  con_t con[ORACLE_MAX_CON];
  /* knotpoint_data.hpp:160-233 */
  double *x, *u, *y, *x_, *u_, *y_, *dynamics_jac_;
  double *lxx_, *luu_, *lux_, *lx_, *lu_, *A_, *B_, *f_;
  double *Qxx_, *Quu_, *Qux_, *Qx_, *Qu_, *Qxx_tmp_, *Quu_tmp_, *Qux_tmp_, *Qx_tmp_, *Qu_tmp_;
  double *K_, *d_, *P_, *p_, *dx_da_, *du_da_;
} knot_t;

struct oracle_solver {
  int N, n, m;
  knot_t *data_;
  double *initial_state_;
  oracle_options opts;
  int model_id;
  double model_params[8];
  oracle_dyn_fn dyn_cb, jac_cb;
  void *dyn_ud;
  int is_initialized;
  oracle_linesearch ls_;
  double phi0_, dphi0_, phi_, dphi_, rho_;
  int ls_iters_;
  int status, iterations;
  long merit_evals;
  double delta_V_[2];
  /* pointer tables for tvlqr (solver.hpp:91-121) */
  int *nx_, *nu_;
  double **tab[26];
};

enum { T_x, T_u, T_y, T_A, T_B, T_f, T_lxx, T_luu, T_lux, T_lx, T_lu, T_K, T_d, T_P, T_p,
       T_Qxx, T_Quu, T_Qux, T_Qx, T_Qu, T_Qxx_tmp, T_Quu_tmp, T_Qux_tmp, T_Qx_tmp, T_Qu_tmp };

oracle_solver *oracle_create(int N, int n, int m) {
  oracle_solver *s = (oracle_solver *)calloc(1, sizeof(oracle_solver));
  s->N = N;
  s->n = n;
  s->m = m;
  s->data_ = (knot_t *)calloc((size_t)N + 1, sizeof(knot_t));
  for (int k = 0; k <= N; ++k) {
    knot_t *z = &s->data_[k];
    z->index = k;
    z->is_terminal = (k == N);
    z->n = n;
    z->m = m;
    z->Q_ = dalloc(n * n);
    z->R_ = dalloc(m * m);
    z->H_ = dalloc(m * n);
    z->q_ = dalloc(n);
    z->r_ = dalloc(m);
    z->affine_term_ = dalloc(n);
    z->A_ = dalloc(n * n);
    z->B_ = dalloc(n * m);
  }
  s->initial_state_ = dalloc(n);
  oracle_default_options(&s->opts);
  s->model_id = ORACLE_MODEL_CALLBACK;
  oracle_ls_init(&s->ls_);
  s->status = ORACLE_STATUS_UNSOLVED;
  s->nx_ = (int *)calloc((size_t)N + 1, sizeof(int));
  s->nu_ = (int *)calloc((size_t)N + 1, sizeof(int));
  for (int t = 0; t < 26; ++t) s->tab[t] = (double **)calloc((size_t)N + 1, sizeof(double *));
  return s;
}

static void free_con(con_t *c) {
  free(c->constraint_val_);
  free(c->constraint_jac_);
  free(c->constraint_hess_);
  free(c->v_);
  free(c->z_);
  free(c->z_est_);
  free(c->z_proj_);
  free(c->proj_jvp_);
  free(c->proj_jac_);
  free(c->proj_hess_);
  free(c->jac_tmp_);
}

void oracle_destroy(oracle_solver *s) {
  if (!s) return;
  for (int k = 0; k <= s->N; ++k) {
    knot_t *z = &s->data_[k];
    double *all[] = {z->Q_, z->R_, z->H_, z->q_, z->r_, z->affine_term_, z->x, z->u, z->y, z->x_,
                     z->u_, z->y_, z->dynamics_jac_, z->lxx_, z->luu_, z->lux_, z->lx_, z->lu_,
                     z->A_, z->B_, z->f_, z->Qxx_, z->Quu_, z->Qux_, z->Qx_, z->Qu_, z->Qxx_tmp_,
                     z->Quu_tmp_, z->Qux_tmp_, z->Qx_tmp_, z->Qu_tmp_, z->K_, z->d_, z->P_, z->p_,
                     z->dx_da_, z->du_da_};
    for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) free(all[i]);
    for (int j = 0; j < z->ncon; ++j) free_con(&z->con[j]);
  }
  free(s->data_);
  free(s->initial_state_);
  free(s->nx_);
  free(s->nu_);
  for (int t = 0; t < 26; ++t) free(s->tab[t]);
  free(s);
}

void oracle_set_options(oracle_solver *s, const oracle_options *o) { s->opts = *o; }

void oracle_set_time_step(oracle_solver *s, float h) {
  for (int k = 0; k < s->N; ++k) s->data_[k].h = h;
}
/* SetTimeStep(h, k_start, k_stop), altro_solver.cpp:49-63 */
void oracle_set_time_step_range(oracle_solver *s, float h, int k_start, int k_stop) {
  for (int k = k_start; k < k_stop && k < s->N; ++k)
    if (k >= 0) s->data_[k].h = h;
}

void oracle_set_model(oracle_solver *s, int model_id, const double *params, int nparams) {
  s->model_id = model_id;
  for (int i = 0; i < 8; ++i) s->model_params[i] = (params && i < nparams) ? params[i] : 0.0;
  for (int k = 0; k < s->N; ++k) {
    s->data_[k].dynamics_is_set = 1;
    s->data_[k].dynamics_are_linear = 0;
  }
}

void oracle_set_dynamics_callback(oracle_solver *s, oracle_dyn_fn dyn, oracle_dyn_fn jac,
                                  void *ud) {
  s->model_id = ORACLE_MODEL_CALLBACK;
  s->dyn_cb = dyn;
  s->jac_cb = jac;
  s->dyn_ud = ud;
  for (int k = 0; k < s->N; ++k) {
    s->data_[k].dynamics_is_set = 1;
    s->data_[k].dynamics_are_linear = 0;
  }
}

/* knotpoint_data.cpp:123-142 */
void oracle_set_linear_dynamics(oracle_solver *s, int k, const double *A, const double *B,
                                const double *f) {
  knot_t *z = &s->data_[k];
  vcopy(z->n * z->n, A, z->A_);
  vcopy(z->n * z->m, B, z->B_);
  if (f) vcopy(z->n, f, z->affine_term_);
  z->dynamics_is_set = 1;
  z->dynamics_are_linear = 1;
}

/* knotpoint_data.cpp:87-110 */
void oracle_set_diagonal_cost(oracle_solver *s, int k, const double *Qd, const double *Rd,
                              const double *q, const double *r, double c) {
  knot_t *z = &s->data_[k];
  int n = z->n, m = z->m;
  vzero(n * n, z->Q_);
  vcopy(n, Qd, z->Q_);
  vzero(m * n, z->H_);
  vcopy(n, q, z->q_);
  z->c_ = c;
  if (!z->is_terminal) {
    vzero(m * m, z->R_);
    vcopy(m, Rd, z->R_);
    vcopy(m, r, z->r_);
  }
  z->cost_is_set = 1;
  z->cost_type = ORACLE_COST_DIAGONAL;
}

/* knotpoint_data.cpp:64-85 */
void oracle_set_quadratic_cost(oracle_solver *s, int k, const double *Q, const double *R,
                               const double *H, const double *q, const double *r, double c) {
  knot_t *z = &s->data_[k];
  int n = z->n, m = z->m;
  vcopy(n * n, Q, z->Q_);
  vcopy(m * m, R, z->R_);
  vcopy(m * n, H, z->H_);
  vcopy(n, q, z->q_);
  vcopy(m, r, z->r_);
  z->c_ = c;
  z->cost_is_set = 1;
  z->cost_type = ORACLE_COST_QUADRATIC;
}

/* altro_solver.cpp:159-169 */
void oracle_set_lqr_cost(oracle_solver *s, int k, const double *Qd, const double *Rd,
                         const double *xref, const double *uref) {
  int n = s->n, m = s->m;
  double q[64], r[64];
  double c = 0;
  for (int i = 0; i < n; ++i) q[i] = -(Qd[i] * xref[i]);
  for (int i = 0; i < m; ++i) r[i] = -(Rd[i] * uref[i]);
  for (int i = 0; i < n; ++i) c += (0.5 * xref[i]) * Qd[i] * xref[i];
  if (k != s->N) {
    double cu = 0;
    for (int i = 0; i < m; ++i) cu += (0.5 * uref[i]) * Rd[i] * uref[i];
    c += cu;
  }
  oracle_set_diagonal_cost(s, k, Qd, Rd, q, r, c);
}

static con_t *add_con(oracle_solver *s, int k, int cone, int dim) {
  knot_t *z = &s->data_[k];
  if (z->ncon >= ORACLE_MAX_CON || dim <= 0) return NULL;
  con_t *c = &z->con[z->ncon++];
  memset(c, 0, sizeof(*c));
  c->cone = cone;
  c->dim = dim;
  return c;
}

int oracle_add_constraint_selector(oracle_solver *s, int k, int cone, int dim, const int *idx,
                                   const double *scale, const double *off) {
  if (dim > ORACLE_MAX_CON_DIM) return -1;
  con_t *c = add_con(s, k, cone, dim);
  if (!c) return -1;
  for (int i = 0; i < dim; ++i) {
    c->idx[i] = idx[i];
    c->scale[i] = scale[i];
    c->off[i] = off[i];
  }
  return s->data_[k].ncon - 1;
}

int oracle_add_constraint_callback(oracle_solver *s, int k, int cone, int dim, oracle_con_fn con,
                                   oracle_con_fn jac, void *ud) {
  con_t *c = add_con(s, k, cone, dim);
  if (!c) return -1;
  c->con = con;
  c->jac = jac;
  c->ud = ud;
  return s->data_[k].ncon - 1;
}

void oracle_set_initial_state(oracle_solver *s, const double *x0) {
  vcopy(s->n, x0, s->initial_state_);
}

static void calc_original_cost_hessian(knot_t *z);

/* knotpoint_data.cpp:229-400 */
static int knot_initialize(knot_t *z) {
  int n = z->n, m = z->m;
  if (n <= 0) return 1;
  if (!z->is_terminal) {
    if (m <= 0) return 2;
    if (!(z->h > 0.0f)) return 10;
    if (!z->dynamics_is_set) return 12;
  }
  if (!z->cost_is_set) return 11;
  z->x = dalloc(n);
  z->u = dalloc(m);
  z->y = dalloc(n);
  z->x_ = dalloc(n);
  z->u_ = dalloc(m);
  z->y_ = dalloc(n);
  z->f_ = dalloc(n);
  z->dynamics_jac_ = dalloc(n * (n + m));
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    int p = c->dim;
    c->constraint_val_ = dalloc(p);
    c->constraint_jac_ = dalloc(p * (n + m));
    c->constraint_hess_ = dalloc((n + m) * (n + m));
    c->v_ = dalloc(p);
    c->z_ = dalloc(p);
    c->z_est_ = dalloc(p);
    c->z_proj_ = dalloc(p);
    c->proj_jvp_ = dalloc(p);
    c->proj_jac_ = dalloc(p * p);
    c->proj_hess_ = dalloc(p * p);
    c->jac_tmp_ = dalloc(p * (n + m));
    c->rho_ = 1.0; /* :343 */
  }
  z->lxx_ = dalloc(n * n);
  z->luu_ = dalloc(m * m);
  z->lux_ = dalloc(m * n);
  z->lx_ = dalloc(n);
  z->lu_ = dalloc(m);
  if (!z->dynamics_are_linear) { /* :353-357 */
    vzero(n * n, z->A_);
    vzero(n * m, z->B_);
  }
  z->Qxx_ = dalloc(n * n);
  z->Quu_ = dalloc(m * m);
  z->Qux_ = dalloc(m * n);
  z->Qx_ = dalloc(n);
  z->Qu_ = dalloc(m);
  z->Qxx_tmp_ = dalloc(n * n);
  z->Quu_tmp_ = dalloc(m * m);
  z->Qux_tmp_ = dalloc(m * n);
  z->Qx_tmp_ = dalloc(n);
  z->Qu_tmp_ = dalloc(m);
  z->K_ = dalloc(m * n);
  z->d_ = dalloc(m);
  z->P_ = dalloc(n * n);
  z->p_ = dalloc(n);
  z->dx_da_ = dalloc(n);
  z->du_da_ = dalloc(m);
  calc_original_cost_hessian(z); /* :383-385 */
  if (z->is_terminal) vcopy(n, z->q_, z->lx_); /* :389-391 */
  if (!z->is_terminal && z->dynamics_are_linear) { /* :392-396 */
    vcopy(n, z->q_, z->lx_);
    vcopy(m, z->r_, z->lu_);
    vcopy(n, z->affine_term_, z->f_);
  }
  return 0;
}

/* solver.cpp:53-110 */
int oracle_initialize(oracle_solver *s) {
  int N = s->N;
  for (int k = 0; k <= N; ++k) {
    int err = knot_initialize(&s->data_[k]);
    if (err) return err;
  }
  for (int k = 0; k <= N; ++k) {
    knot_t *z = &s->data_[k];
    s->nx_[k] = z->n;
    s->nu_[k] = z->m;
    double *ptrs[] = {z->x_, z->u_, z->y_, z->A_, z->B_, z->f_, z->lxx_, z->luu_, z->lux_, z->lx_,
                      z->lu_, z->K_, z->d_, z->P_, z->p_, z->Qxx_, z->Quu_, z->Qux_, z->Qx_,
                      z->Qu_, z->Qxx_tmp_, z->Quu_tmp_, z->Qux_tmp_, z->Qx_tmp_, z->Qu_tmp_};
    for (int t = 0; t < 25; ++t) s->tab[t][k] = ptrs[t];
  }
  s->is_initialized = 1;
  return 0;
}

void oracle_set_state(oracle_solver *s, int k, const double *x) { vcopy(s->n, x, s->data_[k].x_); }
void oracle_set_input(oracle_solver *s, int k, const double *u) { vcopy(s->m, u, s->data_[k].u_); }
void oracle_set_dual(oracle_solver *s, int k, int j, const double *zz) {
  con_t *c = &s->data_[k].con[j];
  vcopy(c->dim, zz, c->z_);
}
/* knotpoint_data.cpp:180-191 over all knots */
void oracle_set_penalty(oracle_solver *s, double rho) {
  for (int k = 0; k <= s->N; ++k)
    for (int j = 0; j < s->data_[k].ncon; ++j) s->data_[k].con[j].rho_ = rho;
}
/* knotpoint_data.cpp:193-224 */
void oracle_update_linear_costs(oracle_solver *s, int k, const double *q, const double *r,
                                double c) {
  knot_t *z = &s->data_[k];
  if (q) vcopy(z->n, q, z->q_);
  if (r && !z->is_terminal) vcopy(z->m, r, z->r_);
  z->c_ = c;
}
/* altro_solver.cpp:283-293 */
void oracle_shift_trajectory(oracle_solver *s) {
  int N = s->N;
  for (int k = 0; k < N; ++k) {
    vcopy(s->n, s->data_[k + 1].x_, s->data_[k].x_);
    if (k < N - 1) vcopy(s->m, s->data_[k + 1].u_, s->data_[k].u_);
  }
}

/* ------------------------------------------------------------ KnotPointData::Calc* */
static oracle_solver *g_unused;

/* knotpoint_data.cpp:710-719 */
static void calc_dynamics(oracle_solver *s, knot_t *z, double *xnext) {
  int n = z->n, m = z->m;
  if (z->dynamics_are_linear) {
    double tmp[64];
    gemm(n, 1, n, z->A_, 0, z->x_, 0, tmp, 0);
    gemm(n, 1, m, z->B_, 0, z->u_, 0, tmp, 1);
    for (int i = 0; i < n; ++i) xnext[i] = tmp[i] + z->affine_term_[i];
  } else if (s->model_id == ORACLE_MODEL_CALLBACK) {
    s->dyn_cb(s->dyn_ud, xnext, z->x_, z->u_, z->h);
  } else {
    oracle_model_dynamics(s->model_id, s->model_params, xnext, z->x_, z->u_, z->h);
  }
}

/* knotpoint_data.cpp:406-419 */
static void calc_dynamics_expansion(oracle_solver *s, knot_t *z) {
  if (z->is_terminal) return;
  int n = z->n, m = z->m;
  if (!z->dynamics_are_linear) {
    if (s->model_id == ORACLE_MODEL_CALLBACK)
      s->jac_cb(s->dyn_ud, z->dynamics_jac_, z->x_, z->u_, z->h);
    else
      oracle_model_jacobian(s->model_id, s->model_params, z->dynamics_jac_, z->x_, z->u_, z->h);
    vcopy(n * n, z->dynamics_jac_, z->A_);
    vcopy(n * m, z->dynamics_jac_ + n * n, z->B_);
  } else {
    vzero(n, z->f_);
  }
}

/* knotpoint_data.cpp:616-648 */
static double calc_original_cost(knot_t *z) {
  int n = z->n, m = z->m;
  double J = 0.0;
  if (z->cost_type == ORACLE_COST_QUADRATIC) {
    double t[64];
    gemm(n, 1, n, z->Q_, 0, z->x_, 0, t, 0);
    J = 0.5 * vdot(n, z->x_, t);
    J += vdot(n, z->q_, z->x_);
    if (!z->is_terminal) {
      gemm(m, 1, m, z->R_, 0, z->u_, 0, t, 0);
      J += 0.5 * vdot(m, z->u_, t);
      J += vdot(m, z->r_, z->u_);
      gemm(m, 1, n, z->H_, 0, z->x_, 0, t, 0);
      J += vdot(m, z->u_, t);
    }
    J += z->c_;
  } else {
    double a = 0;
    for (int i = 0; i < n; ++i) a += (0.5 * z->x_[i]) * z->Q_[i] * z->x_[i];
    J = a;
    J += vdot(n, z->q_, z->x_);
    if (!z->is_terminal) {
      double b = 0;
      for (int i = 0; i < m; ++i) b += (0.5 * z->u_[i]) * z->R_[i] * z->u_[i];
      J += b;
      J += vdot(m, z->r_, z->u_);
    }
    J += z->c_;
  }
  return J;
}

/* knotpoint_data.cpp:650-681 */
static void calc_original_cost_gradient(knot_t *z) {
  int n = z->n, m = z->m;
  if (z->cost_type == ORACLE_COST_QUADRATIC) {
    gemm(n, 1, n, z->Q_, 0, z->x_, 0, z->lx_, 0);
    for (int i = 0; i < n; ++i) z->lx_[i] += z->q_[i];
    if (!z->is_terminal) {
      gemm(m, 1, m, z->R_, 0, z->u_, 0, z->lu_, 0);
      for (int i = 0; i < m; ++i) z->lu_[i] += z->r_[i];
      gemm(m, 1, n, z->H_, 0, z->x_, 0, z->lu_, 1);
      gemm(n, 1, m, z->H_, 1, z->u_, 0, z->lx_, 1);
    }
  } else {
    for (int i = 0; i < n; ++i) z->lx_[i] = z->Q_[i] * z->x_[i];
    for (int i = 0; i < n; ++i) z->lx_[i] += z->q_[i];
    if (!z->is_terminal) {
      for (int i = 0; i < m; ++i) z->lu_[i] = z->R_[i] * z->u_[i];
      for (int i = 0; i < m; ++i) z->lu_[i] += z->r_[i];
    }
  }
}

/* knotpoint_data.cpp:683-708 */
static void calc_original_cost_hessian(knot_t *z) {
  int n = z->n, m = z->m;
  if (z->cost_type == ORACLE_COST_QUADRATIC) {
    vcopy(n * n, z->Q_, z->lxx_);
    if (!z->is_terminal) {
      vcopy(m * m, z->R_, z->luu_);
      vcopy(m * n, z->H_, z->lux_);
    }
  } else {
    vzero(n * n, z->lxx_);
    for (int i = 0; i < n; ++i) z->lxx_[i + n * i] = z->Q_[i];
    if (!z->is_terminal) {
      vzero(m * m, z->luu_);
      for (int i = 0; i < m; ++i) z->luu_[i + m * i] = z->R_[i];
      vzero(m * n, z->lux_);
    }
  }
}

static void selector_eval(const knot_t *z, const con_t *c, double *val) {
  for (int i = 0; i < c->dim; ++i) {
    int id = c->idx[i];
    double v = (id < 0) ? 0.0 : (id < z->n ? z->x_[id] : z->u_[id - z->n]);
    val[i] = (id < 0) ? c->off[i] : c->scale[i] * v + c->off[i];
  }
}
static void selector_jac(const knot_t *z, const con_t *c, double *jac) {
  int p = c->dim;
  vzero(p * (z->n + z->m), jac);
  for (int i = 0; i < p; ++i)
    if (c->idx[i] >= 0) jac[i + p * c->idx[i]] = c->scale[i];
}

/* knotpoint_data.cpp:473-479 */
static void calc_constraints(knot_t *z) {
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    if (c->con)
      c->con(c->ud, c->constraint_val_, z->x_, z->u_);
    else
      selector_eval(z, c, c->constraint_val_);
  }
}
/* knotpoint_data.cpp:481-487 */
static void calc_constraint_jacobians(knot_t *z) {
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    if (c->jac)
      c->jac(c->ud, c->constraint_jac_, z->x_, z->u_);
    else if (!c->con)
      selector_jac(z, c, c->constraint_jac_);
  }
}
/* knotpoint_data.cpp:489-501 */
static double calc_violations(knot_t *z) {
  double viol = 0.0;
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    oracle_conic_projection(c->cone, c->dim, c->constraint_val_, c->v_);
    for (int i = 0; i < c->dim; ++i) c->v_[i] -= c->constraint_val_[i];
    viol = fmax(viol, vinfnorm(c->dim, c->v_));
  }
  return viol;
}
/* knotpoint_data.cpp:503-510 */
static void knot_dual_update(knot_t *z) {
  for (int j = 0; j < z->ncon; ++j) vcopy(z->con[j].dim, z->con[j].z_proj_, z->con[j].z_);
}
/* knotpoint_data.cpp:512-517 */
static void knot_penalty_update(knot_t *z, double scaling, double penalty_max) {
  for (int j = 0; j < z->ncon; ++j) z->con[j].rho_ = fmin(z->con[j].rho_ * scaling, penalty_max);
}
/* knotpoint_data.cpp:523-535 */
static void calc_projected_duals(knot_t *z) {
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    int dual = oracle_dual_cone(c->cone);
    for (int i = 0; i < c->dim; ++i) c->z_est_[i] = c->z_[i] - c->rho_ * c->constraint_val_[i];
    oracle_conic_projection(dual, c->dim, c->z_est_, c->z_proj_);
  }
}
/* knotpoint_data.cpp:537-547 */
static void calc_conic_jacobians(knot_t *z) {
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    int dual = oracle_dual_cone(c->cone);
    oracle_conic_projection_jacobian(dual, c->dim, c->z_est_, c->proj_jac_);
    gemm(c->dim, 1, c->dim, c->proj_jac_, 1, c->z_proj_, 0, c->proj_jvp_, 0);
  }
}
/* knotpoint_data.cpp:549-570 */
static void calc_conic_hessians(knot_t *z) {
  int nm = z->n + z->m;
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    int dual = oracle_dual_cone(c->cone);
    int p = c->dim;
    gemm(p, nm, p, c->proj_jac_, 0, c->constraint_jac_, 0, c->jac_tmp_, 0);
    gemm(nm, nm, p, c->jac_tmp_, 1, c->jac_tmp_, 0, c->constraint_hess_, 0);
    for (int i = 0; i < nm * nm; ++i) c->constraint_hess_[i] *= c->rho_;
    if (!cone_projection_is_linear(dual)) {
      oracle_conic_projection_hessian(dual, p, c->z_est_, c->z_proj_, c->proj_hess_);
      gemm(p, nm, p, c->proj_hess_, 0, c->constraint_jac_, 0, c->jac_tmp_, 0);
      double *tmp = dalloc(nm * nm);
      gemm(nm, nm, p, c->constraint_jac_, 1, c->jac_tmp_, 0, tmp, 0);
      for (int i = 0; i < nm * nm; ++i) c->constraint_hess_[i] += c->rho_ * tmp[i];
      free(tmp);
    }
  }
}
/* knotpoint_data.cpp:572-581 */
static double calc_constraint_costs(knot_t *z) {
  double cost = 0;
  calc_projected_duals(z);
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    cost += vdot(c->dim, c->z_proj_, c->z_proj_) / (2 * c->rho_);
  }
  return cost;
}
/* knotpoint_data.cpp:583-595 */
static void calc_constraint_cost_gradients(knot_t *z) {
  int n = z->n, m = z->m;
  calc_conic_jacobians(z);
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    int p = c->dim;
    gemm(n, 1, p, c->constraint_jac_, 1, c->proj_jvp_, 0, z->lx_, -1);
    if (!z->is_terminal) gemm(m, 1, p, c->constraint_jac_ + p * n, 1, c->proj_jvp_, 0, z->lu_, -1);
  }
}
/* knotpoint_data.cpp:597-613 */
static void calc_constraint_cost_hessians(knot_t *z) {
  int n = z->n, m = z->m, nm = z->n + z->m;
  calc_conic_hessians(z);
  for (int j = 0; j < z->ncon; ++j) {
    con_t *c = &z->con[j];
    for (int jj = 0; jj < n; ++jj)
      for (int i = 0; i < n; ++i) z->lxx_[i + n * jj] += c->constraint_hess_[i + nm * jj];
    if (!z->is_terminal) {
      for (int jj = 0; jj < m; ++jj)
        for (int i = 0; i < m; ++i)
          z->luu_[i + m * jj] += c->constraint_hess_[(n + i) + nm * (n + jj)];
      for (int jj = 0; jj < n; ++jj)
        for (int i = 0; i < m; ++i) z->lux_[i + m * jj] += c->constraint_hess_[(n + i) + nm * jj];
    }
  }
}
/* knotpoint_data.cpp:421-428 */
static double knot_calc_cost(knot_t *z) {
  double cost = calc_original_cost(z);
  double al_cost = calc_constraint_costs(z);
  return cost + al_cost;
}
/* knotpoint_data.cpp:430-437 */
static void knot_calc_cost_gradient(knot_t *z) {
  calc_original_cost_gradient(z);
  calc_constraint_cost_gradients(z);
}
/* knotpoint_data.cpp:439-448 */
static void knot_calc_cost_hessian(knot_t *z) {
  calc_original_cost_hessian(z);
  calc_constraint_cost_hessians(z);
}

void oracle_knot_op(oracle_solver *s, int k, int op) {
  knot_t *z = &s->data_[k];
  (void)g_unused;
  switch (op) {
    case ORACLE_OP_CALC_CONSTRAINTS: calc_constraints(z); break;
    case ORACLE_OP_CALC_CONSTRAINT_JACOBIANS: calc_constraint_jacobians(z); break;
    case ORACLE_OP_CALC_PROJECTED_DUALS: calc_projected_duals(z); break;
    case ORACLE_OP_CALC_CONIC_JACOBIANS: calc_conic_jacobians(z); break;
    case ORACLE_OP_CALC_CONIC_HESSIANS: calc_conic_hessians(z); break;
    case ORACLE_OP_CALC_COST_GRADIENT: knot_calc_cost_gradient(z); break;
    case ORACLE_OP_CALC_COST_HESSIAN: knot_calc_cost_hessian(z); break;
    case ORACLE_OP_CALC_DYNAMICS_EXPANSION: calc_dynamics_expansion(s, z); break;
    case ORACLE_OP_CALC_CONSTRAINT_COST_GRADIENTS: calc_constraint_cost_gradients(z); break;
    case ORACLE_OP_CALC_CONSTRAINT_COST_HESSIANS: calc_constraint_cost_hessians(z); break;
    case ORACLE_OP_CALC_ORIGINAL_COST_GRADIENT: calc_original_cost_gradient(z); break;
    case ORACLE_OP_CALC_ORIGINAL_COST_HESSIAN: calc_original_cost_hessian(z); break;
  }
}
double oracle_knot_calc_cost(oracle_solver *s, int k) { return knot_calc_cost(&s->data_[k]); }
double oracle_knot_calc_constraint_costs(oracle_solver *s, int k) {
  return calc_constraint_costs(&s->data_[k]);
}
double oracle_knot_calc_violations(oracle_solver *s, int k) {
  return calc_violations(&s->data_[k]);
}

/* ============================================================== SolverImpl (solver.cpp) */
/* solver.cpp:116-131 */
void oracle_open_loop_rollout(oracle_solver *s) {
  vcopy(s->n, s->initial_state_, s->data_[0].x_);
  for (int k = 0; k < s->N; ++k) calc_dynamics(s, &s->data_[k], s->data_[k + 1].x_);
}
/* solver.cpp:133-146 */
void oracle_linear_rollout(oracle_solver *s) {
  oracle_tvlqr_forward_pass(s->nx_, s->nu_, s->N, (const double *const *)s->tab[T_A],
                            (const double *const *)s->tab[T_B], (const double *const *)s->tab[T_f],
                            (const double *const *)s->tab[T_K], (const double *const *)s->tab[T_d],
                            (const double *const *)s->tab[T_P], (const double *const *)s->tab[T_p],
                            s->initial_state_, s->tab[T_x], s->tab[T_u], s->tab[T_y]);
}
/* solver.cpp:148-157 */
void oracle_copy_trajectory(oracle_solver *s) {
  for (int k = 0; k <= s->N; ++k) {
    knot_t *z = &s->data_[k];
    vcopy(z->n, z->x_, z->x);
    vcopy(z->n, z->y_, z->y);
    if (k < s->N) vcopy(z->m, z->u_, z->u);
  }
}
/* solver.cpp:163-174 */
double oracle_calc_cost(oracle_solver *s) {
  double cost = 0.0;
  for (int k = 0; k <= s->N; ++k) {
    calc_constraints(&s->data_[k]);
    cost += knot_calc_cost(&s->data_[k]);
  }
  return cost;
}
/* solver.cpp:176-187 */
void oracle_calc_cost_gradient(oracle_solver *s) {
  for (int k = 0; k <= s->N; ++k) knot_calc_cost_gradient(&s->data_[k]);
}
/* solver.cpp:189-201 */
void oracle_calc_expansions(oracle_solver *s) {
  for (int k = 0; k <= s->N; ++k) knot_calc_cost_hessian(&s->data_[k]);
}
/* solver.cpp:207-222 */
double oracle_stationarity(oracle_solver *s) {
  int N = s->N;
  double res_x = 0, res_u = 0;
  double t[64];
  for (int k = 0; k < N; ++k) {
    knot_t *z = &s->data_[k], *zn = &s->data_[k + 1];
    int n = z->n, m = z->m;
    gemm(n, 1, n, z->A_, 1, zn->y_, 0, t, 0);
    for (int i = 0; i < n; ++i) t[i] = z->lx_[i] + t[i] - z->y_[i];
    res_x = fmax(res_x, vinfnorm(n, t));
    gemm(m, 1, n, z->B_, 1, zn->y_, 0, t, 0);
    for (int i = 0; i < m; ++i) t[i] = z->lu_[i] + t[i];
    res_u = fmax(res_u, vinfnorm(m, t));
  }
  knot_t *z = &s->data_[N];
  for (int i = 0; i < z->n; ++i) t[i] = z->lx_[i] - z->y_[i];
  res_x = fmax(res_x, vinfnorm(z->n, t));
  return fmax(res_x, res_u);
}
/* solver.cpp:224-231 */
double oracle_feasibility(oracle_solver *s) {
  double viol = 0;
  for (int k = 0; k <= s->N; ++k) viol = fmax(viol, calc_violations(&s->data_[k]));
  return viol;
}

/* solver.cpp:273-355 */
void oracle_merit_function(oracle_solver *s, double alpha, double *phi, double *dphi) {
  int N = s->N;
  int calc_derivative = dphi != NULL;
  s->merit_evals += 1;
  s->phi_ = 0;
  s->dphi_ = 0;
  vcopy(s->n, s->initial_state_, s->data_[0].x_);
  vzero(s->n, s->data_[0].dx_da_);
  double dx[64], du[64];
  for (int k = 0; k < N; ++k) {
    knot_t *z = &s->data_[k], *zn = &s->data_[k + 1];
    int n = z->n, m = z->m;
    for (int i = 0; i < n; ++i) dx[i] = z->x_[i] - z->x[i];                 /* :290 */
    gemm(m, 1, n, z->K_, 0, dx, 0, du, 0);                                    /* :291 */
    for (int i = 0; i < m; ++i) du[i] = -du[i] + alpha * z->d_[i];
    for (int i = 0; i < m; ++i) z->u_[i] = z->u[i] + du[i];                   /* :292 */
    gemm(n, 1, n, z->P_, 0, dx, 0, z->y_, 0);                                 /* :293 */
    for (int i = 0; i < n; ++i) z->y_[i] += z->p_[i];
    calc_dynamics(s, z, zn->x_);                                              /* :296 */
    calc_constraints(z);                                                      /* :299 */
    s->phi_ += knot_calc_cost(z);                                             /* :300-301 */
    if (calc_derivative) {
      calc_dynamics_expansion(s, z);                                          /* :305 */
      gemm(m, 1, n, z->K_, 0, z->dx_da_, 0, z->du_da_, 0);                    /* :306 */
      for (int i = 0; i < m; ++i) z->du_da_[i] = -z->du_da_[i] + z->d_[i];
      gemm(n, 1, n, z->A_, 0, z->dx_da_, 0, zn->dx_da_, 0);                   /* :307-308 */
      gemm(n, 1, m, z->B_, 0, z->du_da_, 0, zn->dx_da_, 1);
      calc_constraint_jacobians(z);                                           /* :311 */
      knot_calc_cost_gradient(z);                                             /* :312 */
      s->dphi_ += vdot(n, z->lx_, z->dx_da_);                                 /* :313-314 */
      s->dphi_ += vdot(m, z->lu_, z->du_da_);
    }
  }
  knot_t *z = &s->data_[N]; /* :319-332 */
  int n = z->n;
  calc_constraints(z);
  s->phi_ += knot_calc_cost(z);
  for (int i = 0; i < n; ++i) dx[i] = z->x_[i] - z->x[i];
  gemm(n, 1, n, z->P_, 0, dx, 0, z->y_, 0);
  for (int i = 0; i < n; ++i) z->y_[i] += z->p_[i];
  *phi = s->phi_;
  if (calc_derivative) {
    calc_constraint_jacobians(z);
    knot_calc_cost_gradient(z);
    s->dphi_ += vdot(n, z->lx_, z->dx_da_);
    *dphi = s->dphi_;
  }
}

static void merit_trampoline(void *ctx, double alpha, double *phi, double *dphi) {
  oracle_merit_function((oracle_solver *)ctx, alpha, phi, dphi);
}

/* solver.cpp:237-271 */
int oracle_forward_pass(oracle_solver *s, double *alpha) {
  oracle_merit_function(s, 0.0, &s->phi0_, &s->dphi0_);
  if (fabs(s->dphi0_) < s->opts.tol_meritfun_gradient) { /* :242, std::fabs (SURVEY 8c hazard) */
    *alpha = 0.0;
    return ORACLE_MERIT_GRAD_TOO_SMALL;
  }
  s->ls_.try_cubic_first = 1; /* :248 */
  s->ls_.c1 = s->opts.ls_c1;
  s->ls_.c2 = s->opts.ls_c2;
  *alpha = oracle_ls_run(&s->ls_, merit_trampoline, s, 1.0, s->phi0_, s->dphi0_);
  s->phi_ = s->ls_.phi; /* GetFinalMeritValues, :250 */
  s->dphi_ = s->ls_.dphi;
  int res = s->ls_.return_code;
  s->ls_iters_ = s->ls_.n_iters;
  if (s->opts.use_backtracking_linesearch && (fabs(*alpha - 1.0) > 0)) { /* :256-262 */
    for (int k = 0; k <= s->N; ++k) {
      calc_dynamics_expansion(s, &s->data_[k]);
      calc_constraint_jacobians(&s->data_[k]);
      knot_calc_cost_gradient(&s->data_[k]);
    }
  }
  if (isnan(*alpha) || !(res == 1 /*MINIMUM_FOUND*/ || res == 7 /*HIT_MAX_STEPSIZE*/)) {
    return ORACLE_LINESEARCH_FAILED;
  }
  return ORACLE_NOERROR;
}

/* solver.cpp:360-378 */
int oracle_backward_pass(oracle_solver *s) {
  int res = oracle_tvlqr_backward_pass(
      s->nx_, s->nu_, s->N, (const double *const *)s->tab[T_A], (const double *const *)s->tab[T_B],
      (const double *const *)s->tab[T_f], (const double *const *)s->tab[T_lxx],
      (const double *const *)s->tab[T_luu], (const double *const *)s->tab[T_lux],
      (const double *const *)s->tab[T_lx], (const double *const *)s->tab[T_lu], 0.0, s->tab[T_K],
      s->tab[T_d], s->tab[T_P], s->tab[T_p], s->delta_V_, s->tab[T_Qxx], s->tab[T_Quu],
      s->tab[T_Qux], s->tab[T_Qx], s->tab[T_Qu], s->tab[T_Qxx_tmp], s->tab[T_Quu_tmp],
      s->tab[T_Qux_tmp], s->tab[T_Qx_tmp], s->tab[T_Qu_tmp], 0, 0);
  return res != -1 ? ORACLE_BACKWARD_PASS_FAILED : ORACLE_NOERROR;
}

/* solver.cpp:383-395 */
void oracle_dual_update(oracle_solver *s) {
  for (int k = 0; k <= s->N; ++k) knot_dual_update(&s->data_[k]);
}
/* solver.cpp:397-409 */
void oracle_penalty_update(oracle_solver *s) {
  for (int k = 0; k <= s->N; ++k)
    knot_penalty_update(&s->data_[k], s->opts.penalty_scaling, s->opts.penalty_max);
  s->rho_ = fmin(s->rho_ * s->opts.penalty_scaling, s->opts.penalty_max);
}

/* solver.cpp:414-511 */
int oracle_solve(oracle_solver *s) {
  int N = s->N;
  s->ls_.use_backtracking_linesearch = s->opts.use_backtracking_linesearch;
  s->rho_ = s->opts.penalty_initial;
  s->merit_evals = 0;

  oracle_open_loop_rollout(s);
  oracle_copy_trajectory(s);
  (void)oracle_calc_cost(s);
  for (int k = 0; k <= N; ++k) { /* :425-430 -- gradient BEFORE the penalty reset (quirk Q3) */
    calc_dynamics_expansion(s, &s->data_[k]);
    calc_constraint_jacobians(&s->data_[k]);
    knot_calc_cost_gradient(&s->data_[k]);
    for (int j = 0; j < s->data_[k].ncon; ++j) s->data_[k].con[j].rho_ = s->opts.penalty_initial;
  }
  double alpha;
  int is_converged = 0, stop_iterating = 0;
  s->status = ORACLE_STATUS_UNSOLVED;
  int iter;
  for (iter = 0; iter < s->opts.iterations_max; ++iter) {
    oracle_calc_expansions(s);
    (void)oracle_backward_pass(s); /* result ignored, :449 (quirk Q2) */
    int err = oracle_forward_pass(s, &alpha);
    if (!(err == ORACLE_NOERROR || err == ORACLE_MERIT_GRAD_TOO_SMALL)) stop_iterating = 1;

    double stationarity = oracle_stationarity(s);
    double feasibility = oracle_feasibility(s);
    oracle_copy_trajectory(s);

    if (fabs(stationarity) < s->opts.tol_stationarity &&
        feasibility < s->opts.tol_primal_feasibility) {
      is_converged = 1;
      stop_iterating = 1;
      s->status = ORACLE_STATUS_SUCCESS;
    }
    if (stationarity < sqrt(s->opts.tol_stationarity)) { /* :474-489 */
      oracle_dual_update(s);
      if (feasibility > s->opts.tol_primal_feasibility) oracle_penalty_update(s);
      for (int k = 0; k <= N; ++k) {
        calc_projected_duals(&s->data_[k]);
        knot_calc_cost_gradient(&s->data_[k]);
      }
    }
    if (stop_iterating) break;
  }
  if (!is_converged && iter == s->opts.iterations_max) s->status = ORACLE_STATUS_MAX_ITERATIONS;
  s->iterations = iter + 1; /* :506 (quirk Q4) */
  return ORACLE_NOERROR;
}

int oracle_get_status(const oracle_solver *s) { return s->status; }
int oracle_get_iterations(const oracle_solver *s) { return s->iterations; }
long oracle_get_merit_evals(const oracle_solver *s) { return s->merit_evals; }
double oracle_get_final_phi(const oracle_solver *s) { return s->phi_; }
int oracle_ls_iters(const oracle_solver *s) { return s->ls_iters_; }
void oracle_get_merit_values(const oracle_solver *s, double *phi0, double *dphi0, double *phi,
                             double *dphi) {
  *phi0 = s->phi0_;
  *dphi0 = s->dphi0_;
  *phi = s->phi_;
  *dphi = s->dphi_;
}

/* ------------------------------------------------------------ field access */
static double *field_ptr(const oracle_solver *s, int k, const char *name, int *len) {
  const knot_t *z = &s->data_[k];
  int n = z->n, m = z->m;
#define F(nm, p, l)              \
  if (strcmp(name, nm) == 0) {   \
    *len = (l);                  \
    return (p);                  \
  }
  F("x", z->x, n) F("u", z->u, m) F("y", z->y, n) F("x_", z->x_, n) F("u_", z->u_, m)
  F("y_", z->y_, n) F("A_", z->A_, n * n) F("B_", z->B_, n * m) F("f_", z->f_, n)
  F("lxx_", z->lxx_, n * n) F("luu_", z->luu_, m * m) F("lux_", z->lux_, m * n)
  F("lx_", z->lx_, n) F("lu_", z->lu_, m) F("K_", z->K_, m * n) F("d_", z->d_, m)
  F("P_", z->P_, n * n) F("p_", z->p_, n) F("Qxx_", z->Qxx_, n * n) F("Quu_", z->Quu_, m * m)
  F("Qux_", z->Qux_, m * n) F("Qx_", z->Qx_, n) F("Qu_", z->Qu_, m) F("dx_da_", z->dx_da_, n)
  F("du_da_", z->du_da_, m) F("q_", z->q_, n) F("r_", z->r_, m) F("Q_", z->Q_, n * n)
  F("R_", z->R_, m * m) F("H_", z->H_, m * n)
  F("dynamics_jac_", z->dynamics_jac_, n * (n + m))
#undef F
  /* per-constraint members: "<member>_<j>" */
  const char *us = strrchr(name, '_');
  if (us && us[1] >= '0' && us[1] <= '9') {
    int j = atoi(us + 1);
    if (j < 0 || j >= z->ncon) return NULL;
    const con_t *c = &z->con[j];
    int p = c->dim, nm = n + m;
    size_t base = (size_t)(us - name);
#define G(nmm, ptr, l)                                         \
  if (strlen(nmm) == base && strncmp(name, nmm, base) == 0) {  \
    *len = (l);                                                \
    return (ptr);                                              \
  }
    G("constraint_val", c->constraint_val_, p) G("constraint_jac", c->constraint_jac_, p * nm)
    G("constraint_hess", c->constraint_hess_, nm * nm) G("v", c->v_, p) G("z", c->z_, p)
    G("z_est", c->z_est_, p) G("z_proj", c->z_proj_, p) G("proj_jvp", c->proj_jvp_, p)
    G("proj_jac", c->proj_jac_, p * p) G("proj_hess", c->proj_hess_, p * p)
    G("rho", (double *)&c->rho_, 1)
#undef G
  }
  return NULL;
}

int oracle_get_field(const oracle_solver *s, int k, const char *name, double *out) {
  int len = 0;
  if (strcmp(name, "c_") == 0) {
    out[0] = s->data_[k].c_;
    return 1;
  }
  double *p = field_ptr(s, k, name, &len);
  if (!p) return -1;
  vcopy(len, p, out);
  return len;
}
int oracle_set_field(oracle_solver *s, int k, const char *name, const double *in) {
  int len = 0;
  double *p = field_ptr(s, k, name, &len);
  if (!p) return -1;
  vcopy(len, in, p);
  return len;
}
