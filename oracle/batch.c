/*
 * batch.c -- ORACLE (test infrastructure only, see altro_oracle.h).
 *
 * Drives the single-problem oracle (altro_oracle.c) over a batch of independent problems, one
 * problem per OpenMP thread.  Used (i) as the checker for the GPU parity tests and (ii) as the
 * timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`), SURVEY.md 8d "CPU
 * baseline timing".  Problem setup mirrors what a user of the reference would write with B
 * separate altro::ALTROSolver objects (test/pendulum_test.cpp:69-98, test/bicycle_test.cpp:
 * 144-224): SetDimension / SetTimeStep / SetExplicitDynamics / SetLQRCost / SetConstraint /
 * SetInitialState / Initialize / SetInput / Solve.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "altro_oracle.h"

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

oracle_solver *oracle_batch_make_solver(const oracle_batch_spec *sp, int b) {
  int N = sp->N, n = sp->n, m = sp->m;
  oracle_solver *s = oracle_create(N, n, m);
  oracle_set_options(s, &sp->opts);
  oracle_set_time_step(s, sp->h);
  oracle_set_model(s, sp->model_id, sp->model_params, 8);
  double zero_u[ORACLE_MAX_CON_DIM * 4];
  memset(zero_u, 0, sizeof(zero_u));
  for (int k = 0; k <= N; ++k) {
    const double *Qd = sp->Qd + (size_t)k * n;
    const double *Rd = (k < N) ? sp->Rd + (size_t)k * m : sp->Rd + (size_t)(N - 1) * m;
    switch (sp->ref_mode) {
      case 0: { /* shared q, r, c */
        const double *q = sp->q + (size_t)k * n;
        const double *r = (k < N) ? sp->r + (size_t)k * m : zero_u;
        oracle_set_diagonal_cost(s, k, Qd, Rd, q, r, sp->c[k]);
        break;
      }
      case 1: { /* per-problem q, r, c */
        const double *q = sp->q + ((size_t)b * (N + 1) + k) * n;
        const double *r = (k < N) ? sp->r + ((size_t)b * N + k) * m : zero_u;
        oracle_set_diagonal_cost(s, k, Qd, Rd, q, r, sp->c[(size_t)b * (N + 1) + k]);
        break;
      }
      case 2: /* per-problem goal */
        oracle_set_lqr_cost(s, k, Qd, Rd, sp->xref + (size_t)b * n, sp->uref + (size_t)b * m);
        break;
      case 3: { /* window into the shared reference table */
        int row = sp->offsets[b] + k;
        int urow = row; /* the table holds T rows of both xref and uref */
        oracle_set_lqr_cost(s, k, Qd, Rd, sp->xref + (size_t)row * n,
                            sp->uref + (size_t)urow * m);
        break;
      }
    }
  }
  for (int j = 0; j < sp->ncon; ++j) {
    const oracle_con_spec *c = &sp->con[j];
    const double *off = c->off_b ? c->off_b + (size_t)b * c->dim : c->off;
    for (int k = c->k_start; k < c->k_stop; ++k)
      oracle_add_constraint_selector(s, k, c->cone, c->dim, c->idx, c->scale, off);
  }
  oracle_set_initial_state(s, sp->x0 + (size_t)b * n);
  oracle_initialize(s);
  for (int k = 0; k < N; ++k) {
    const double *u = sp->U0_per_problem ? sp->U0 + ((size_t)b * N + k) * m : sp->U0 + (size_t)k * m;
    oracle_set_input(s, k, u);
  }
  return s;
}

double oracle_batch_solve(const oracle_batch_spec *sp, int b0, int b1, int nthreads,
                          oracle_batch_result *res) {
  int N = sp->N, n = sp->n, m = sp->m;
  if (nthreads < 1) nthreads = 1;
  int nb = b1 - b0;
  oracle_solver **solvers = (oracle_solver **)calloc((size_t)(nb > 0 ? nb : 1), sizeof(*solvers));
  /* phase 1 (untimed): problem definition, as the user of the reference would do it */
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
  for (int b = b0; b < b1; ++b) solvers[b - b0] = oracle_batch_make_solver(sp, b);
  /* phase 2 (timed): Solve() only */
  double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int b = b0; b < b1; ++b) oracle_solve(solvers[b - b0]);
  double elapsed = now_s() - t0;
  /* phase 3 (untimed): gather */
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
  for (int b = b0; b < b1; ++b) {
    oracle_solver *s = solvers[b - b0];
    size_t o = (size_t)(b - b0);
    for (int k = 0; k <= N; ++k) {
      if (res->X) oracle_get_field(s, k, "x_", res->X + (o * (N + 1) + k) * n);
      if (res->Y) oracle_get_field(s, k, "y_", res->Y + (o * (N + 1) + k) * n);
      if (k < N && res->U) oracle_get_field(s, k, "u_", res->U + (o * N + k) * m);
    }
    if (res->status) res->status[o] = oracle_get_status(s);
    if (res->iters) res->iters[o] = oracle_get_iterations(s);
    if (res->merit_evals) res->merit_evals[o] = oracle_get_merit_evals(s);
    if (res->cost) res->cost[o] = oracle_get_final_phi(s);
    if (res->stat) res->stat[o] = oracle_stationarity(s);
    if (res->feas) res->feas[o] = oracle_feasibility(s);
    oracle_destroy(s);
  }
  free(solvers);
  return elapsed;
}
