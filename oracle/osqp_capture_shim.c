/*
 * oracle/osqp_capture_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * libosqp.so stand-in for the SHIPPED reference binaries (src/libtrp.so, src/libcub.so):
 * provides the five symbols they import (csc_matrix, osqp_set_default_settings,
 * osqp_setup, osqp_solve, osqp_cleanup).  osqp_setup dumps the
 * QP the reference assembled (FormulateProblem, solve_3d.cc:1143-1229) to the file named by
 * $SPECTRAL_QP_DUMP with %.17g, then:
 *   - if $SPECTRAL_QP_SOLUTION names a file holding n doubles (binary), osqp_solve returns
 *     them as the solution with status SOLVED, so the shipped binary goes on to its own
 *     Bezier sampling, cost and output-file code with a solution of our choosing;
 *   - otherwise osqp_solve reports status -10 (unsolved) and find_traj returns 1e11.
 */
#include <stdio.h>
#include "shim_include/osqp/osqp.h"

csc *csc_matrix(c_int m, c_int n, c_int nzmax, c_float *x, c_int *i, c_int *p) {
  csc *M = (csc *)malloc(sizeof(csc));
  M->m = m; M->n = n; M->nz = -1; M->nzmax = nzmax; M->x = x; M->i = i; M->p = p;
  return M;
}

void osqp_set_default_settings(OSQPSettings *s) {
  s->rho = 0.1; s->sigma = 1e-6; s->scaling = 10; s->adaptive_rho = 1;
  s->adaptive_rho_interval = 0; s->adaptive_rho_tolerance = 5; s->adaptive_rho_fraction = 0.4;
  s->max_iter = 4000; s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4;
  s->eps_dual_inf = 1e-4; s->alpha = 1.6; s->linsys_solver = QDLDL_SOLVER; s->delta = 1e-6;
  s->polish = 0; s->polish_refine_iter = 3; s->verbose = 1; s->scaled_termination = 0;
  s->check_termination = 25; s->warm_start = 1; s->time_limit = 0;
}

static void dump_vec(FILE *f, const char *name, const c_float *v, c_int n) {
  fprintf(f, "%s %lld\n", name, n);
  for (c_int k = 0; k < n; k++) fprintf(f, "%.17g\n", v[k]);
}
static void dump_ivec(FILE *f, const char *name, const c_int *v, c_int n) {
  fprintf(f, "%s %lld\n", name, n);
  for (c_int k = 0; k < n; k++) fprintf(f, "%lld\n", v[k]);
}

OSQPWorkspace *osqp_setup(const OSQPData *d, OSQPSettings *s) {
  const char *path = getenv("SPECTRAL_QP_DUMP");
  if (path) {
    FILE *f = fopen(path, "w");
    if (f) {
      fprintf(f, "n %lld\nm %lld\n", d->n, d->m);
      fprintf(f, "settings rho %.17g sigma %.17g scaling %lld adaptive_rho %lld max_iter %lld "
                 "eps_abs %.17g eps_rel %.17g eps_prim_inf %.17g eps_dual_inf %.17g alpha %.17g "
                 "polish %lld verbose %lld scaled_termination %lld check_termination %lld\n",
              s->rho, s->sigma, s->scaling, s->adaptive_rho, s->max_iter, s->eps_abs, s->eps_rel,
              s->eps_prim_inf, s->eps_dual_inf, s->alpha, s->polish, s->verbose,
              s->scaled_termination, s->check_termination);
      dump_vec(f, "q", d->q, d->n);
      dump_vec(f, "l", d->l, d->m);
      dump_vec(f, "u", d->u, d->m);
      dump_ivec(f, "P_p", d->P->p, d->P->n + 1);
      dump_ivec(f, "P_i", d->P->i, d->P->p[d->P->n]);
      dump_vec(f, "P_x", d->P->x, d->P->p[d->P->n]);
      dump_ivec(f, "A_p", d->A->p, d->A->n + 1);
      dump_ivec(f, "A_i", d->A->i, d->A->p[d->A->n]);
      dump_vec(f, "A_x", d->A->x, d->A->p[d->A->n]);
      fclose(f);
    }
  }
  OSQPWorkspace *w = (OSQPWorkspace *)calloc(1, sizeof(OSQPWorkspace));
  w->settings = s;
  w->info = (OSQPInfo *)calloc(1, sizeof(OSQPInfo));
  w->solution = (OSQPSolution *)calloc(1, sizeof(OSQPSolution));
  w->solution->x = (c_float *)calloc(d->n, sizeof(c_float));
  w->solution->y = (c_float *)calloc(d->m, sizeof(c_float));
  w->info->status_val = OSQP_UNSOLVED;
  strcpy(w->info->status, "unsolved");
  const char *sol = getenv("SPECTRAL_QP_SOLUTION");
  if (sol) {
    FILE *f = fopen(sol, "rb");
    if (f) {
      if (fread(w->solution->x, sizeof(c_float), d->n, f) == (size_t)d->n) {
        w->info->status_val = OSQP_SOLVED;
        strcpy(w->info->status, "solved");
        /* objective of the injected point, 0.5 x'Px + q'x, P upper triangular */
        double obj = 0.0;
        for (c_int j = 0; j < d->n; j++) {
          obj += d->q[j] * w->solution->x[j];
          for (c_int p = d->P->p[j]; p < d->P->p[j + 1]; p++) {
            c_int i = d->P->i[p];
            double v = d->P->x[p] * w->solution->x[i] * w->solution->x[j];
            obj += (i == j) ? 0.5 * v : v;
          }
        }
        w->info->obj_val = obj;
      }
      fclose(f);
    }
  }
  return w;
}

c_int osqp_solve(OSQPWorkspace *w) { (void)w; return 0; }

c_int osqp_cleanup(OSQPWorkspace *w) {
  if (w) {
    free(w->solution->x); free(w->solution->y); free(w->solution); free(w->info); free(w);
  }
  return 0;
}
