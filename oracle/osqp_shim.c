/*
 * oracle/osqp_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Solve-mode implementation of the five OSQP 0.5.0 entry points the reference links
 * against (solve_3d.cc:1211-1249,1451; makefile:2 -losqp), backed by the restatement in
 * osqp_restate.c.  Linked into oracle/_ref/libref_{trp,cub}.so together with the
 * reference's own sources.  Extras (not OSQP API):
 *   spectral_shim_last()         per-thread copy of the last solution / info, so the in-memory
 *                                driver can read the control points, which the reference keeps
 *                                only inside the OSQP workspace (solve_3d.cc:1343-1344);
 *   spectral_shim_set_override() per-thread settings override used for the CONVERGED oracle
 *                                (tight eps + polish); default is "exactly what the reference set".
 */
#include <stdio.h>
#include "shim_include/osqp/osqp.h"
#include "osqp_restate.h"
#include "osqp_shim.h"

static __thread SpectralShimLast g_last;
static __thread SpectralShimOverride g_ovr;

SpectralShimLast *spectral_shim_last(void) { return &g_last; }
void spectral_shim_set_override(const SpectralShimOverride *o) {
  if (o) g_ovr = *o; else memset(&g_ovr, 0, sizeof(g_ovr));
}

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

typedef struct { const OSQPData *data; } ShimState;

OSQPWorkspace *osqp_setup(const OSQPData *d, OSQPSettings *s) {
  OSQPWorkspace *w = (OSQPWorkspace *)calloc(1, sizeof(OSQPWorkspace));
  w->settings = s;
  w->info = (OSQPInfo *)calloc(1, sizeof(OSQPInfo));
  w->solution = (OSQPSolution *)calloc(1, sizeof(OSQPSolution));
  w->solution->x = (c_float *)calloc(d->n + 1, sizeof(c_float));
  w->solution->y = (c_float *)calloc(d->m + 1, sizeof(c_float));
  w->info->status_val = OSQP_UNSOLVED;
  ShimState *st = (ShimState *)calloc(1, sizeof(ShimState));
  st->data = d;
  w->shim_state = st;
  return w;
}

c_int osqp_solve(OSQPWorkspace *w) {
  const OSQPData *d = ((ShimState *)w->shim_state)->data;
  const OSQPSettings *s = w->settings;
  OsqpRestateSettings r;
  r.rho = s->rho; r.sigma = s->sigma; r.scaling = s->scaling; r.adaptive_rho = s->adaptive_rho;
  r.adaptive_rho_interval = s->adaptive_rho_interval; r.adaptive_rho_tolerance = s->adaptive_rho_tolerance;
  r.max_iter = s->max_iter; r.eps_abs = s->eps_abs; r.eps_rel = s->eps_rel;
  r.eps_prim_inf = s->eps_prim_inf; r.eps_dual_inf = s->eps_dual_inf; r.alpha = s->alpha;
  r.delta = s->delta; r.polish = s->polish; r.polish_refine_iter = s->polish_refine_iter;
  r.scaled_termination = s->scaled_termination; r.check_termination = s->check_termination;
  r.polish_rounds = 1;
  if (g_ovr.active) {
    if (g_ovr.eps > 0) { r.eps_abs = g_ovr.eps; r.eps_rel = g_ovr.eps; }
    if (g_ovr.max_iter > 0) r.max_iter = g_ovr.max_iter;
    r.polish = g_ovr.polish;
    if (g_ovr.delta > 0) r.delta = g_ovr.delta;
    if (g_ovr.polish_refine_iter > 0) r.polish_refine_iter = g_ovr.polish_refine_iter;
    if (g_ovr.polish_rounds > 0) r.polish_rounds = g_ovr.polish_rounds;
  }
  OsqpRestateInfo info;
  osqp_restate_solve(d->n, d->m, d->P->p, d->P->i, d->P->x, d->q, d->A->p, d->A->i, d->A->x,
                     d->l, d->u, &r, w->solution->x, w->solution->y, &info);
  w->info->iter = info.iter;
  w->info->status_val = info.status;
  w->info->obj_val = info.obj_val;
  w->info->pri_res = info.pri_res;
  w->info->dua_res = info.dua_res;
  w->info->rho_updates = info.rho_updates;
  w->info->rho_estimate = info.rho_final;
  w->info->status_polish = info.polish_status;
  snprintf(w->info->status, sizeof(w->info->status), "%s",
           info.status == 1 ? "solved" : info.status == 2 ? "solved inaccurate" : "not solved");
  g_last.n = (int)d->n; g_last.m = (int)d->m;
  g_last.status = info.status; g_last.iter = info.iter; g_last.polish_status = info.polish_status;
  g_last.rho_updates = info.rho_updates; g_last.obj_val = info.obj_val;
  g_last.pri_res = info.pri_res; g_last.dua_res = info.dua_res;
  int nn = d->n < SPECTRAL_SHIM_MAX_N ? (int)d->n : SPECTRAL_SHIM_MAX_N;
  memcpy(g_last.x, w->solution->x, nn * sizeof(double));
  return 0;
}

c_int osqp_cleanup(OSQPWorkspace *w) {
  if (w) {
    free(w->solution->x); free(w->solution->y); free(w->solution); free(w->info);
    free(w->shim_state); free(w);
  }
  return 0;
}
