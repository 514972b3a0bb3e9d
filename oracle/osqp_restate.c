/*
 * oracle/osqp_restate.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * CPU restatement of the OSQP 0.5.0 algorithm, the third-party solver the reference calls
 * at solve_3d.cc:1246-1249 / cuboid_3d.cc:1105-1108 (osqp_setup + osqp_solve) with the
 * settings of solve_3d.cc:1235-1243,1446-1462.  OSQP is NOT vendored under /root/reference
 * (makefile:2 links -losqp; the only version pin is the comment "osqp-0.4.1, 0.5.0" at
 * solve_3d.cc:1246 plus the 0.5.0 struct layout recovered from the shipped binaries,
 * SURVEY.md Appendix C).  What follows restates the published algorithm (Stellato et al.,
 * "OSQP: an operator splitting solver for quadratic programs"; OSQP 0.5.0 sources as
 * publicly documented):
 *   - modified Ruiz equilibration of [P A'; A 0] with cost scaling (`scaling` passes),
 *   - per-constraint rho (equality rows x1e3), sigma regularisation,
 *   - quasi-definite KKT matrix [P+sigma I, A'; A, -diag(1/rho)] factorised by an
 *     up-looking sparse LDL' (the algorithm of QDLDL / T. Davis' LDL), constraints ordered
 *     first so the fill is that of the banded reduced system,
 *   - the ADMM iteration with relaxation alpha, termination / infeasibility checks every
 *     `check_termination` iterations in the scaled space when scaled_termination=1,
 *   - adaptive rho.  OSQP 0.5.0 picks the adaptation interval from WALL-CLOCK time
 *     (adaptive_rho_fraction of the setup time), which makes its iterate sequence
 *     non-reproducible; we use OSQP's own no-profiling rule instead (a fixed interval of
 *     4 x check_termination = 100 iterations).
 *   - optional solution polishing (active-set KKT solve + iterative refinement).
 * PARITY: OSQP's own iterates are NOT pinned by anything under /root/reference (3-decimal
 * output files only).  The QP optimum is unique (P is positive definite on every fixture),
 * so parity on the QP solution is anchored on the converged optimum: this file with
 * tight tolerances + polish, independently certified by a KKT-residual check and by a
 * HiGHS cross-check in tests/ ("parity unpinned beyond 1e-3").
 */
#include "osqp_restate.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OSQP_INFTY 1e20
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_TOL 1e-4
#define RHO_EQ_OVER_RHO_INEQ 1e3

typedef long long ll;

/* ------------------------------------------------------------------ small helpers */
static double vmax_abs(const double *v, ll n) {
  double r = 0.0;
  for (ll i = 0; i < n; i++) { double a = fabs(v[i]); if (a > r) r = a; }
  return r;
}
static double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  v = v > MAX_SCALING ? MAX_SCALING : v;
  return v;
}

/* CSC matrix (owned copy) */
typedef struct { ll m, n, nnz; ll *p, *i; double *x; } Csc;

static Csc csc_copy(ll m, ll n, const ll *p, const ll *i, const double *x) {
  Csc M; M.m = m; M.n = n; M.nnz = p[n];
  M.p = (ll *)malloc((n + 1) * sizeof(ll)); memcpy(M.p, p, (n + 1) * sizeof(ll));
  M.i = (ll *)malloc((M.nnz + 1) * sizeof(ll)); memcpy(M.i, i, M.nnz * sizeof(ll));
  M.x = (double *)malloc((M.nnz + 1) * sizeof(double)); memcpy(M.x, x, M.nnz * sizeof(double));
  return M;
}
static void csc_free(Csc *M) { free(M->p); free(M->i); free(M->x); }

/* y (+)= A x */
static void mat_vec(const Csc *A, const double *x, double *y, int accumulate) {
  if (!accumulate) for (ll i = 0; i < A->m; i++) y[i] = 0.0;
  for (ll j = 0; j < A->n; j++)
    for (ll p = A->p[j]; p < A->p[j + 1]; p++) y[A->i[p]] += A->x[p] * x[j];
}
/* y (+)= A' x */
static void mat_tpose_vec(const Csc *A, const double *x, double *y, int accumulate) {
  for (ll j = 0; j < A->n; j++) {
    double s = accumulate ? y[j] : 0.0;
    for (ll p = A->p[j]; p < A->p[j + 1]; p++) s += A->x[p] * x[A->i[p]];
    y[j] = s;
  }
}
/* y = P x, P symmetric stored as upper triangle */
static void sym_mat_vec(const Csc *P, const double *x, double *y) {
  for (ll i = 0; i < P->n; i++) y[i] = 0.0;
  for (ll j = 0; j < P->n; j++)
    for (ll p = P->p[j]; p < P->p[j + 1]; p++) {
      ll i = P->i[p];
      y[i] += P->x[p] * x[j];
      if (i != j) y[j] += P->x[p] * x[i];
    }
}
static double quad_form(const Csc *P, const double *x) {
  double q = 0.0;
  for (ll j = 0; j < P->n; j++)
    for (ll p = P->p[j]; p < P->p[j + 1]; p++) {
      ll i = P->i[p];
      if (i == j) q += 0.5 * P->x[p] * x[i] * x[i];
      else if (i < j) q += P->x[p] * x[i] * x[j];
    }
  return q;
}

/* ------------------------------------------------------------------ sparse LDL' (up-looking)
 * Factorises a symmetric quasi-definite matrix given as upper-triangular CSC. */
typedef struct {
  ll n; ll *Lp, *Li; double *Lx, *D, *Dinv;
  ll *parent, *Lnz, *flag, *pattern; double *Y;
} Ldl;

static void ldl_free(Ldl *F) {
  free(F->Lp); free(F->Li); free(F->Lx); free(F->D); free(F->Dinv);
  free(F->parent); free(F->Lnz); free(F->flag); free(F->pattern); free(F->Y);
  memset(F, 0, sizeof(*F));
}

static int ldl_symbolic(Ldl *F, ll n, const ll *Ap, const ll *Ai) {
  F->n = n;
  F->Lp = (ll *)malloc((n + 1) * sizeof(ll));
  F->parent = (ll *)malloc(n * sizeof(ll));
  F->Lnz = (ll *)malloc(n * sizeof(ll));
  F->flag = (ll *)malloc(n * sizeof(ll));
  F->pattern = (ll *)malloc(n * sizeof(ll));
  F->Y = (double *)malloc(n * sizeof(double));
  F->D = (double *)malloc(n * sizeof(double));
  F->Dinv = (double *)malloc(n * sizeof(double));
  for (ll k = 0; k < n; k++) {
    F->parent[k] = -1; F->flag[k] = k; F->Lnz[k] = 0;
    for (ll p = Ap[k]; p < Ap[k + 1]; p++) {
      ll i = Ai[p];
      if (i < k) {
        for (; F->flag[i] != k; i = F->parent[i]) {
          if (F->parent[i] == -1) F->parent[i] = k;
          F->Lnz[i]++;
          F->flag[i] = k;
        }
      }
    }
  }
  F->Lp[0] = 0;
  for (ll k = 0; k < n; k++) F->Lp[k + 1] = F->Lp[k] + F->Lnz[k];
  F->Li = (ll *)malloc((F->Lp[n] + 1) * sizeof(ll));
  F->Lx = (double *)malloc((F->Lp[n] + 1) * sizeof(double));
  return 0;
}

static int ldl_numeric(Ldl *F, const ll *Ap, const ll *Ai, const double *Ax) {
  ll n = F->n;
  for (ll k = 0; k < n; k++) {
    F->Y[k] = 0.0;
    ll top = n;
    F->flag[k] = k;
    F->Lnz[k] = 0;
    for (ll p = Ap[k]; p < Ap[k + 1]; p++) {
      ll i = Ai[p];
      if (i <= k) {
        F->Y[i] += Ax[p];
        ll len = 0;
        for (; F->flag[i] != k; i = F->parent[i]) { F->pattern[len++] = i; F->flag[i] = k; }
        while (len > 0) F->pattern[--top] = F->pattern[--len];
      }
    }
    F->D[k] = F->Y[k];
    F->Y[k] = 0.0;
    for (; top < n; top++) {
      ll i = F->pattern[top];
      double yi = F->Y[i];
      F->Y[i] = 0.0;
      ll p2 = F->Lp[i] + F->Lnz[i];
      for (ll p = F->Lp[i]; p < p2; p++) F->Y[F->Li[p]] -= F->Lx[p] * yi;
      double lki = yi / F->D[i];
      F->D[k] -= lki * yi;
      F->Li[p2] = k;
      F->Lx[p2] = lki;
      F->Lnz[i]++;
    }
    if (F->D[k] == 0.0) return -1;
    F->Dinv[k] = 1.0 / F->D[k];
  }
  return 0;
}

static void ldl_solve(const Ldl *F, double *x) {
  ll n = F->n;
  for (ll j = 0; j < n; j++)
    for (ll p = F->Lp[j]; p < F->Lp[j + 1]; p++) x[F->Li[p]] -= F->Lx[p] * x[j];
  for (ll j = 0; j < n; j++) x[j] *= F->Dinv[j];
  for (ll j = n - 1; j >= 0; j--)
    for (ll p = F->Lp[j]; p < F->Lp[j + 1]; p++) x[j] -= F->Lx[p] * x[F->Li[p]];
}

/* ------------------------------------------------------------------ KKT assembly
 * Ordering: the m constraint nodes first (index i), then the n variables (index m+j); with
 * this ordering the fill-in is exactly that of P + sigma I + A' diag(rho) A.
 * Upper-triangular CSC: column i<m holds only its diagonal; column m+j holds A(:,j)
 * (rows = constraint ids), then P(0..j, j) shifted by m with sigma added on the diagonal. */
typedef struct { ll dim; ll *p, *i; double *x; ll *diag_c; /* position of -1/rho_i */ } Kkt;

static void kkt_free(Kkt *K) { free(K->p); free(K->i); free(K->x); free(K->diag_c); }

static void kkt_build(Kkt *K, const Csc *P, const Csc *A, double sigma, const double *rho_inv_neg,
                      ll n, ll m) {
  ll dim = n + m;
  K->dim = dim;
  K->p = (ll *)malloc((dim + 1) * sizeof(ll));
  ll cap = m + A->nnz + P->nnz + n + 1;
  K->i = (ll *)malloc(cap * sizeof(ll));
  K->x = (double *)malloc(cap * sizeof(double));
  K->diag_c = (ll *)malloc((m + 1) * sizeof(ll));
  ll nz = 0;
  for (ll c = 0; c < m; c++) {
    K->p[c] = nz;
    K->diag_c[c] = nz;
    K->i[nz] = c; K->x[nz] = rho_inv_neg[c]; nz++;
  }
  for (ll j = 0; j < n; j++) {
    K->p[m + j] = nz;
    for (ll q = A->p[j]; q < A->p[j + 1]; q++) { K->i[nz] = A->i[q]; K->x[nz] = A->x[q]; nz++; }
    int have_diag = 0;
    for (ll q = P->p[j]; q < P->p[j + 1]; q++) {
      ll r = P->i[q];
      if (r > j) continue; /* only the upper triangle is meaningful */
      K->i[nz] = m + r;
      K->x[nz] = P->x[q] + (r == j ? sigma : 0.0);
      if (r == j) have_diag = 1;
      nz++;
    }
    if (!have_diag) { K->i[nz] = m + j; K->x[nz] = sigma; nz++; }
  }
  K->p[dim] = nz;
}

/* ------------------------------------------------------------------ the solver */
void osqp_restate_default_settings(OsqpRestateSettings *s) {
  s->rho = 0.1; s->sigma = 1e-6; s->scaling = 10; s->adaptive_rho = 1;
  s->adaptive_rho_interval = 0; s->adaptive_rho_tolerance = 5.0; s->max_iter = 4000;
  s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->alpha = 1.6; s->delta = 1e-6; s->polish = 0; s->polish_refine_iter = 3;
  s->scaled_termination = 0; s->check_termination = 25; s->polish_rounds = 1;
}

typedef struct {
  ll n, m;
  Csc P, A;           /* scaled */
  double *q, *l, *u;  /* scaled */
  double *D, *E, *Dinv, *Einv; double c, cinv;
  const OsqpRestateSettings *s;
  double rho; double *rho_vec, *rho_inv_neg; int *constr_type;
  Kkt K; Ldl F;
  double *x, *z, *y, *x_prev, *z_prev, *xz_tilde, *delta_y, *delta_x;
  double *Ax, *Px, *Aty, *tmp_n, *tmp_m;
} Work;

static void scale_data(Work *w) {
  ll n = w->n, m = w->m;
  double *Dt = (double *)malloc(n * sizeof(double));
  double *Et = (double *)malloc((m + 1) * sizeof(double));
  for (ll j = 0; j < n; j++) w->D[j] = 1.0;
  for (ll i = 0; i < m; i++) w->E[i] = 1.0;
  w->c = 1.0;
  for (ll it = 0; it < w->s->scaling; it++) {
    /* inf-norms of the columns of [P A'; A 0] */
    for (ll j = 0; j < n; j++) Dt[j] = 0.0;
    for (ll i = 0; i < m; i++) Et[i] = 0.0;
    for (ll j = 0; j < n; j++)
      for (ll p = w->P.p[j]; p < w->P.p[j + 1]; p++) {
        ll i = w->P.i[p]; double a = fabs(w->P.x[p]);
        if (a > Dt[j]) Dt[j] = a;
        if (i != j && a > Dt[i]) Dt[i] = a;
      }
    for (ll j = 0; j < n; j++)
      for (ll p = w->A.p[j]; p < w->A.p[j + 1]; p++) {
        ll i = w->A.i[p]; double a = fabs(w->A.x[p]);
        if (a > Dt[j]) Dt[j] = a;
        if (a > Et[i]) Et[i] = a;
      }
    for (ll j = 0; j < n; j++) Dt[j] = 1.0 / sqrt(limit_scaling(Dt[j]));
    for (ll i = 0; i < m; i++) Et[i] = 1.0 / sqrt(limit_scaling(Et[i]));
    /* equilibrate */
    for (ll j = 0; j < n; j++)
      for (ll p = w->P.p[j]; p < w->P.p[j + 1]; p++) w->P.x[p] *= Dt[w->P.i[p]] * Dt[j];
    for (ll j = 0; j < n; j++)
      for (ll p = w->A.p[j]; p < w->A.p[j + 1]; p++) w->A.x[p] *= Et[w->A.i[p]] * Dt[j];
    for (ll j = 0; j < n; j++) { w->q[j] *= Dt[j]; w->D[j] *= Dt[j]; }
    for (ll i = 0; i < m; i++) w->E[i] *= Et[i];
    /* cost normalisation */
    for (ll j = 0; j < n; j++) Dt[j] = 0.0;
    for (ll j = 0; j < n; j++)
      for (ll p = w->P.p[j]; p < w->P.p[j + 1]; p++) {
        ll i = w->P.i[p]; double a = fabs(w->P.x[p]);
        if (a > Dt[j]) Dt[j] = a;
        if (i != j && a > Dt[i]) Dt[i] = a;
      }
    double ct = 0.0;
    for (ll j = 0; j < n; j++) ct += Dt[j];
    ct /= (double)n;
    double qn = limit_scaling(vmax_abs(w->q, n));
    ct = ct > qn ? ct : qn;
    ct = limit_scaling(ct);
    ct = 1.0 / ct;
    for (ll p = 0; p < w->P.nnz; p++) w->P.x[p] *= ct;
    for (ll j = 0; j < n; j++) w->q[j] *= ct;
    w->c *= ct;
  }
  w->cinv = 1.0 / w->c;
  for (ll j = 0; j < n; j++) w->Dinv[j] = 1.0 / w->D[j];
  for (ll i = 0; i < m; i++) { w->Einv[i] = 1.0 / w->E[i]; w->l[i] *= w->E[i]; w->u[i] *= w->E[i]; }
  free(Dt); free(Et);
}

static void set_rho_vec(Work *w) {
  double rho = w->rho;
  rho = rho < RHO_MIN ? RHO_MIN : (rho > RHO_MAX ? RHO_MAX : rho);
  w->rho = rho;
  for (ll i = 0; i < w->m; i++) {
    if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) {
      w->constr_type[i] = -1; w->rho_vec[i] = RHO_MIN;
    } else if (w->u[i] - w->l[i] < RHO_TOL) {
      w->constr_type[i] = 1; w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * rho;
    } else {
      w->constr_type[i] = 0; w->rho_vec[i] = rho;
    }
    w->rho_inv_neg[i] = -1.0 / w->rho_vec[i];
  }
}

static int refactor(Work *w) {
  for (ll i = 0; i < w->m; i++) w->K.x[w->K.diag_c[i]] = w->rho_inv_neg[i];
  return ldl_numeric(&w->F, w->K.p, w->K.i, w->K.x);
}

static double compute_obj(Work *w, const double *x) {
  double o = quad_form(&w->P, x);
  for (ll j = 0; j < w->n; j++) o += w->q[j] * x[j];
  return o * w->cinv;
}

/* residuals in the space selected by scaled_termination */
static void compute_res(Work *w, double *pri, double *dua, double *eps_pri_scale, double *eps_dua_scale) {
  ll n = w->n, m = w->m;
  int unscale = (w->s->scaling && !w->s->scaled_termination);
  mat_vec(&w->A, w->x, w->Ax, 0);
  double pr = 0.0, nz = 0.0, nax = 0.0;
  for (ll i = 0; i < m; i++) {
    double f = unscale ? w->Einv[i] : 1.0;
    double r = fabs(f * (w->Ax[i] - w->z[i])); if (r > pr) pr = r;
    double a = fabs(f * w->z[i]); if (a > nz) nz = a;
    a = fabs(f * w->Ax[i]); if (a > nax) nax = a;
  }
  sym_mat_vec(&w->P, w->x, w->Px);
  mat_tpose_vec(&w->A, w->y, w->Aty, 0);
  double dr = 0.0, nq = 0.0, npx = 0.0, naty = 0.0;
  for (ll j = 0; j < n; j++) {
    double f = unscale ? w->cinv * w->Dinv[j] : 1.0;
    double r = fabs(f * (w->Px[j] + w->q[j] + w->Aty[j])); if (r > dr) dr = r;
    double a = fabs(f * w->q[j]); if (a > nq) nq = a;
    a = fabs(f * w->Px[j]); if (a > npx) npx = a;
    a = fabs(f * w->Aty[j]); if (a > naty) naty = a;
  }
  *pri = pr; *dua = dr;
  *eps_pri_scale = nz > nax ? nz : nax;
  double t = nq > npx ? nq : npx;
  *eps_dua_scale = t > naty ? t : naty;
}

static int is_primal_infeasible(Work *w, double eps) {
  ll n = w->n, m = w->m;
  int unscale = (w->s->scaling && !w->s->scaled_termination);
  double nd = 0.0;
  for (ll i = 0; i < m; i++) {
    double v = unscale ? w->E[i] * w->delta_y[i] : w->delta_y[i];
    v = fabs(v); if (v > nd) nd = v;
  }
  if (nd > eps) {
    double lhs = 0.0;
    for (ll i = 0; i < m; i++) {
      double dy = w->delta_y[i];
      if (w->u[i] > OSQP_INFTY * MIN_SCALING) {
        if (w->l[i] < -OSQP_INFTY * MIN_SCALING) dy = 0.0; else dy = dy < 0.0 ? dy : 0.0;
      } else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) dy = dy > 0.0 ? dy : 0.0;
      w->tmp_m[i] = dy;
      lhs += w->u[i] * (dy > 0.0 ? dy : 0.0) + w->l[i] * (dy < 0.0 ? dy : 0.0);
    }
    if (lhs < -eps * nd) {
      mat_tpose_vec(&w->A, w->tmp_m, w->tmp_n, 0);
      double na = 0.0;
      for (ll j = 0; j < n; j++) {
        double v = unscale ? w->Dinv[j] * w->tmp_n[j] : w->tmp_n[j];
        v = fabs(v); if (v > na) na = v;
      }
      return na < eps * nd;
    }
  }
  return 0;
}

static int is_dual_infeasible(Work *w, double eps) {
  ll n = w->n, m = w->m;
  int unscale = (w->s->scaling && !w->s->scaled_termination);
  double nd = 0.0, cost_scaling = 1.0;
  for (ll j = 0; j < n; j++) {
    double v = unscale ? w->D[j] * w->delta_x[j] : w->delta_x[j];
    v = fabs(v); if (v > nd) nd = v;
  }
  if (unscale) cost_scaling = w->c;
  if (nd > eps) {
    double qdx = 0.0;
    for (ll j = 0; j < n; j++) qdx += w->q[j] * w->delta_x[j];
    if (qdx < -cost_scaling * eps * nd) {
      sym_mat_vec(&w->P, w->delta_x, w->tmp_n);
      double np = 0.0;
      for (ll j = 0; j < n; j++) {
        double v = unscale ? w->Dinv[j] * w->tmp_n[j] : w->tmp_n[j];
        v = fabs(v); if (v > np) np = v;
      }
      if (np < cost_scaling * eps * nd) {
        mat_vec(&w->A, w->delta_x, w->tmp_m, 0);
        for (ll i = 0; i < m; i++) {
          double v = unscale ? w->Einv[i] * w->tmp_m[i] : w->tmp_m[i];
          if ((w->u[i] < OSQP_INFTY * MIN_SCALING && v > eps * nd) ||
              (w->l[i] > -OSQP_INFTY * MIN_SCALING && v < -eps * nd))
            return 0;
        }
        return 1;
      }
    }
  }
  return 0;
}

/* returns status or 0 (keep going) */
static int check_termination(Work *w, int approximate, double *pri_out, double *dua_out) {
  double eps_abs = w->s->eps_abs, eps_rel = w->s->eps_rel;
  double eps_pinf = w->s->eps_prim_inf, eps_dinf = w->s->eps_dual_inf;
  if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; eps_dinf *= 10; }
  double pri, dua, sp, sd;
  compute_res(w, &pri, &dua, &sp, &sd);
  if (pri_out) *pri_out = pri;
  if (dua_out) *dua_out = dua;
  int prim_ok = 0, dual_ok = 0, pinf = 0, dinf = 0;
  if (w->m == 0) prim_ok = 1;
  else {
    double eps_prim = eps_abs + eps_rel * sp;
    if (pri < eps_prim) prim_ok = 1; else pinf = is_primal_infeasible(w, eps_pinf);
  }
  double eps_dual = eps_abs + eps_rel * sd;
  if (dua < eps_dual) dual_ok = 1; else dinf = is_dual_infeasible(w, eps_dinf);
  if (prim_ok && dual_ok) return approximate ? OSQP_RESTATE_SOLVED_INACCURATE : OSQP_RESTATE_SOLVED;
  if (pinf) return approximate ? OSQP_RESTATE_PRIMAL_INFEASIBLE_INACCURATE : OSQP_RESTATE_PRIMAL_INFEASIBLE;
  if (dinf) return approximate ? OSQP_RESTATE_DUAL_INFEASIBLE_INACCURATE : OSQP_RESTATE_DUAL_INFEASIBLE;
  return 0;
}

static double compute_rho_estimate(Work *w) {
  /* always in the scaled space (OSQP compute_rho_estimate) */
  ll n = w->n, m = w->m;
  mat_vec(&w->A, w->x, w->Ax, 0);
  double pr = 0.0, nz = 0.0, nax = 0.0;
  for (ll i = 0; i < m; i++) {
    double r = fabs(w->Ax[i] - w->z[i]); if (r > pr) pr = r;
    double a = fabs(w->z[i]); if (a > nz) nz = a;
    a = fabs(w->Ax[i]); if (a > nax) nax = a;
  }
  sym_mat_vec(&w->P, w->x, w->Px);
  mat_tpose_vec(&w->A, w->y, w->Aty, 0);
  double dr = 0.0, nq = 0.0, npx = 0.0, naty = 0.0;
  for (ll j = 0; j < n; j++) {
    double r = fabs(w->Px[j] + w->q[j] + w->Aty[j]); if (r > dr) dr = r;
    double a = fabs(w->q[j]); if (a > nq) nq = a;
    a = fabs(w->Px[j]); if (a > npx) npx = a;
    a = fabs(w->Aty[j]); if (a > naty) naty = a;
  }
  pr /= ((nz > nax ? nz : nax) + 1e-10);
  double t = nq > npx ? nq : npx; t = t > naty ? t : naty;
  dr /= (t + 1e-10);
  double est = w->rho * sqrt(pr / (dr + 1e-10));
  est = est < RHO_MIN ? RHO_MIN : (est > RHO_MAX ? RHO_MAX : est);
  return est;
}

/* -------------------------------------------------- polish (OSQP polish.c, restated dense-free)
 * One round is OSQP's polish: guess the active set from (z, y), solve the equality-constrained
 * KKT system regularised by delta with iterative refinement.  With s->polish_rounds > 1 the
 * active set is then CORRECTED (rows violated by the polished point are added, active rows whose
 * multiplier has the wrong sign are dropped) and the solve repeated; when a round ends with no
 * violated row and no wrong-signed multiplier the point satisfies the KKT conditions of the
 * convex QP, i.e. it is the optimum: *verified = 1.  (This extension is only used for the
 * CONVERGED oracle; the reference's own settings have polish = 0.) */
static int polish(Work *w, double *pol_x, double *pol_y_full, int *n_active_out, int *verified) {
  ll n = w->n, m = w->m;
  int *act = (int *)calloc(m + 1, sizeof(int)); /* -1 lower, +1 upper, 0 inactive */
  for (ll i = 0; i < m; i++) {
    if (w->z[i] - w->l[i] < -w->y[i]) act[i] = -1;
    else if (w->u[i] - w->z[i] < w->y[i]) act[i] = 1;
    if (w->constr_type[i] == 1 && act[i] == 0) act[i] = -1; /* equality rows are always active */
  }
  ll *ind = (ll *)malloc((m + 1) * sizeof(ll));
  ll *rowmap = (ll *)malloc((m + 1) * sizeof(ll));
  double *b = (double *)malloc((m + 1) * sizeof(double));
  double *rhs = (double *)malloc((n + m + 1) * sizeof(double));
  double *px = (double *)malloc((n + 1) * sizeof(double));
  double *xx = (double *)calloc(n + 1, sizeof(double));
  double *yy = (double *)calloc(m + 1, sizeof(double));
  double *ax = (double *)calloc(m + 1, sizeof(double));
  double *dneg = (double *)malloc((m + 1) * sizeof(double));
  Csc Ar; Ar.m = 0; Ar.n = n;
  Ar.p = (ll *)malloc((n + 1) * sizeof(ll));
  Ar.i = (ll *)malloc((w->A.nnz + 1) * sizeof(ll));
  Ar.x = (double *)malloc((w->A.nnz + 1) * sizeof(double));
  int ok = 1, rounds = w->s->polish_rounds > 0 ? (int)w->s->polish_rounds : 1;
  *verified = 0;
  ll k = 0;
  for (int round = 0; round < rounds && ok; round++) {
    k = 0;
    for (ll i = 0; i < m; i++) {
      rowmap[i] = -1;
      if (act[i]) { ind[k] = i; b[k] = act[i] < 0 ? w->l[i] : w->u[i]; rowmap[i] = k; k++; }
    }
    ll nz = 0;
    for (ll j = 0; j < n; j++) {
      Ar.p[j] = nz;
      for (ll p = w->A.p[j]; p < w->A.p[j + 1]; p++)
        if (rowmap[w->A.i[p]] >= 0) { Ar.i[nz] = rowmap[w->A.i[p]]; Ar.x[nz] = w->A.x[p]; nz++; }
    }
    Ar.p[n] = nz; Ar.nnz = nz; Ar.m = k;
    double delta = w->s->delta;
    for (ll a = 0; a < k; a++) dneg[a] = -delta;
    Kkt K; Ldl F; memset(&F, 0, sizeof(F));
    /* converged-oracle mode (polish_rounds > 1): the primal regularisation is decoupled from delta and kept
     * tiny -- with delta on both blocks the refinement is a proximal iteration that stalls along the weakly
     * curved directions of P (scaled eigenvalues ~1e-9 on these problems) */
    kkt_build(&K, &w->P, &Ar, rounds > 1 ? 1e-12 : delta, dneg, n, k);
    ldl_symbolic(&F, K.dim, K.p, K.i);
    ok = ldl_numeric(&F, K.p, K.i, K.x) == 0;
    if (ok) {
      for (ll j = 0; j < n; j++) xx[j] = 0.0;
      for (ll a = 0; a < k; a++) yy[a] = 0.0;
      for (int it = 0; it <= w->s->polish_refine_iter; it++) {
        /* residual of the UNREGULARISED system: [-q - P x - Ar' y ; b - Ar x] */
        sym_mat_vec(&w->P, xx, px);
        mat_tpose_vec(&Ar, yy, w->tmp_n, 0);
        for (ll j = 0; j < n; j++) rhs[k + j] = -w->q[j] - px[j] - w->tmp_n[j];
        for (ll a = 0; a < k; a++) rhs[a] = b[a];
        for (ll j = 0; j < n; j++)
          for (ll p = Ar.p[j]; p < Ar.p[j + 1]; p++) rhs[Ar.i[p]] -= Ar.x[p] * xx[j];
        ldl_solve(&F, rhs);
        for (ll a = 0; a < k; a++) yy[a] += rhs[a];
        for (ll j = 0; j < n; j++) xx[j] += rhs[k + j];
      }
    }
    kkt_free(&K); ldl_free(&F);
    if (!ok) break;
    /* KKT check of the polished point and active-set correction */
    mat_vec(&w->A, xx, ax, 0);
    double nax = vmax_abs(ax, m), ny = vmax_abs(yy, k);
    double tol_p = 1e-9 * (1.0 + nax), tol_d = 1e-9 * (1.0 + ny);
    int changed = 0;
    for (ll i = 0; i < m; i++) {
      if (!act[i]) {
        if (ax[i] < w->l[i] - tol_p) { act[i] = -1; changed++; }
        else if (ax[i] > w->u[i] + tol_p) { act[i] = 1; changed++; }
      } else if (w->constr_type[i] != 1) {
        double yi = yy[rowmap[i]];
        if ((act[i] < 0 && yi > tol_d) || (act[i] > 0 && yi < -tol_d)) { act[i] = 0; changed++; }
      }
    }
    /* stationarity and active-row feasibility of the point must hold too before "no change" counts as a
     * KKT proof: on ill-conditioned instances the delta-regularised solve + refinement can stall short of
     * the solution of the UNREGULARISED active-set system */
    sym_mat_vec(&w->P, xx, px);
    mat_tpose_vec(&Ar, yy, w->tmp_n, 0);
    double r_d = 0.0, dscale = 0.0, r_p = 0.0;
    for (ll j = 0; j < n; j++) {
      double r = fabs(px[j] + w->q[j] + w->tmp_n[j]);
      double sc = fmax(fabs(px[j]), fmax(fabs(w->q[j]), fabs(w->tmp_n[j])));
      if (r > r_d) r_d = r;
      if (sc > dscale) dscale = sc;
    }
    for (ll i = 0; i < m; i++) {
      double v = fmax(fmax(w->l[i] - ax[i], ax[i] - w->u[i]), 0.0);
      if (v > r_p) r_p = v;
    }
    int exact = (r_d <= 1e-8 * (1.0 + dscale)) && (r_p <= 10.0 * tol_p) && r_d == r_d && r_p == r_p;
    if (!changed) { if (exact) *verified = 1; break; }
  }
  if (ok) {
    for (ll j = 0; j < n; j++) pol_x[j] = xx[j];
    for (ll i = 0; i < m; i++) pol_y_full[i] = 0.0;
    for (ll a = 0; a < k; a++) pol_y_full[ind[a]] = yy[a];
  }
  *n_active_out = (int)k;
  free(rhs); free(px); free(xx); free(yy); free(ax); free(dneg);
  csc_free(&Ar); free(rowmap); free(ind); free(b); free(act);
  return ok ? 0 : -1;
}

int osqp_restate_solve(ll n, ll m, const ll *Pp, const ll *Pi, const double *Px_, const double *q,
                       const ll *Ap, const ll *Ai, const double *Ax_, const double *l,
                       const double *u, const OsqpRestateSettings *s, double *x_out,
                       double *y_out, OsqpRestateInfo *info) {
  Work W; memset(&W, 0, sizeof(W));
  Work *w = &W;
  w->n = n; w->m = m; w->s = s;
  w->P = csc_copy(n, n, Pp, Pi, Px_);
  w->A = csc_copy(m, n, Ap, Ai, Ax_);
#define DV(name, cnt) w->name = (double *)calloc((cnt) + 1, sizeof(double))
  DV(q, n); DV(l, m); DV(u, m); DV(D, n); DV(E, m); DV(Dinv, n); DV(Einv, m);
  DV(rho_vec, m); DV(rho_inv_neg, m); DV(x, n); DV(z, m); DV(y, m); DV(x_prev, n); DV(z_prev, m);
  DV(xz_tilde, n + m); DV(delta_y, m); DV(delta_x, n); DV(Ax, m); DV(Px, n); DV(Aty, n);
  DV(tmp_n, n); DV(tmp_m, m);
#undef DV
  w->constr_type = (int *)calloc(m + 1, sizeof(int));
  memcpy(w->q, q, n * sizeof(double)); memcpy(w->l, l, m * sizeof(double)); memcpy(w->u, u, m * sizeof(double));
  if (s->scaling) scale_data(w);
  else {
    for (ll j = 0; j < n; j++) w->D[j] = w->Dinv[j] = 1.0;
    for (ll i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.0;
    w->c = w->cinv = 1.0;
  }
  w->rho = s->rho;
  set_rho_vec(w);
  kkt_build(&w->K, &w->P, &w->A, s->sigma, w->rho_inv_neg, n, m);
  ldl_symbolic(&w->F, w->K.dim, w->K.p, w->K.i);
  int status = OSQP_RESTATE_UNSOLVED;
  int iter = 0, rho_updates = 0;
  double pri = 0, dua = 0;
  ll adapt_interval = s->adaptive_rho_interval;
  if (s->adaptive_rho && adapt_interval == 0)
    adapt_interval = s->check_termination ? 4 * s->check_termination : 100;
  if (ldl_numeric(&w->F, w->K.p, w->K.i, w->K.x) != 0) { status = OSQP_RESTATE_NON_CVX; goto done; }
  /* OSQP's validate_data (called by osqp_setup): "Lower bound at index i is greater than upper bound" -> setup fails and
   * no iteration runs.  (The reference then dereferences the NULL workspace, solve_3d.cc:1251 -- undefined; the restatement
   * reports the problem as unsolved after 0 iterations, which the wrapper maps to its failure sentinel.)  Checked on the
   * caller's unscaled bounds like OSQP does. */
  for (ll i = 0; i < m; i++)
    if (l[i] > u[i]) { status = OSQP_RESTATE_UNSOLVED; iter = 0; goto done; }

  for (iter = 1; iter <= s->max_iter; iter++) {
    double *t;
    t = w->x; w->x = w->x_prev; w->x_prev = t;
    t = w->z; w->z = w->z_prev; w->z_prev = t;
    /* rhs in KKT ordering: [constraints ; variables] */
    for (ll i = 0; i < m; i++) w->xz_tilde[i] = w->z_prev[i] - w->y[i] / w->rho_vec[i];
    for (ll j = 0; j < n; j++) w->xz_tilde[m + j] = s->sigma * w->x_prev[j] - w->q[j];
    ldl_solve(&w->F, w->xz_tilde);
    /* z_tilde = z_prev + (nu - y)/rho */
    for (ll i = 0; i < m; i++)
      w->xz_tilde[i] = w->z_prev[i] + (w->xz_tilde[i] - w->y[i]) / w->rho_vec[i];
    for (ll j = 0; j < n; j++) {
      w->x[j] = s->alpha * w->xz_tilde[m + j] + (1.0 - s->alpha) * w->x_prev[j];
      w->delta_x[j] = w->x[j] - w->x_prev[j];
    }
    for (ll i = 0; i < m; i++) {
      double zr = s->alpha * w->xz_tilde[i] + (1.0 - s->alpha) * w->z_prev[i];
      double zz = zr + w->y[i] / w->rho_vec[i];
      zz = zz < w->l[i] ? w->l[i] : (zz > w->u[i] ? w->u[i] : zz);
      w->z[i] = zz;
      w->delta_y[i] = w->rho_vec[i] * (zr - zz);
      w->y[i] += w->delta_y[i];
    }
    int checked = 0;
    if (s->check_termination && iter % s->check_termination == 0) {
      int st = check_termination(w, 0, &pri, &dua);
      checked = 1;
      if (st) { status = st; break; }
    }
    if (s->adaptive_rho && adapt_interval && iter % adapt_interval == 0) {
      double est = compute_rho_estimate(w);
      if (est > w->rho * s->adaptive_rho_tolerance || est < w->rho / s->adaptive_rho_tolerance) {
        w->rho = est;
        set_rho_vec(w);
        if (refactor(w) != 0) { status = OSQP_RESTATE_NON_CVX; break; }
        rho_updates++;
      }
    }
    (void)checked;
  }
  if (iter > s->max_iter) iter = s->max_iter;
  if (status == OSQP_RESTATE_UNSOLVED) {
    int st = check_termination(w, 0, &pri, &dua);
    if (!st) st = check_termination(w, 1, &pri, &dua);
    status = st ? st : OSQP_RESTATE_MAX_ITER_REACHED;
  }
done:
  info->polish_status = 0;
  info->n_active = 0;
  if ((status == OSQP_RESTATE_SOLVED || status == OSQP_RESTATE_SOLVED_INACCURATE) && s->polish) {
    double *px_ = (double *)malloc((n + 1) * sizeof(double));
    double *py_ = (double *)malloc((m + 1) * sizeof(double));
    int nact = 0, verified = 0;
    if (polish(w, px_, py_, &nact, &verified) == 0) {
      /* accept iff the polished point has smaller (scaled) residuals than the ADMM point */
      double *sx = w->x, *sy = w->y, *sz = w->z;
      double *pz = (double *)malloc((m + 1) * sizeof(double));
      mat_vec(&w->A, px_, pz, 0);
      for (ll i = 0; i < m; i++) pz[i] = pz[i] < w->l[i] ? w->l[i] : (pz[i] > w->u[i] ? w->u[i] : pz[i]);
      double p0, d0, p1, d1, a, b2;
      compute_res(w, &p0, &d0, &a, &b2);
      w->x = px_; w->y = py_; w->z = pz;
      compute_res(w, &p1, &d1, &a, &b2);
      if ((p1 < p0 && d1 < d0) || (p1 < p0 && d0 < 1e-10) || (d1 < d0 && p0 < 1e-10)) {
        memcpy(sx, px_, n * sizeof(double)); memcpy(sy, py_, m * sizeof(double)); memcpy(sz, pz, m * sizeof(double));
        info->polish_status = verified ? 2 : 1; pri = p1; dua = d1;
      } else info->polish_status = -1;
      w->x = sx; w->y = sy; w->z = sz;
      free(pz);
    } else info->polish_status = -1;
    info->n_active = nact;
    free(px_); free(py_);
  }
  info->status = status; info->iter = iter; info->rho_updates = rho_updates;
  info->pri_res = pri; info->dua_res = dua; info->rho_final = w->rho;
  info->obj_val = compute_obj(w, w->x);
  for (ll j = 0; j < n; j++) x_out[j] = w->D[j] * w->x[j];
  if (y_out) for (ll i = 0; i < m; i++) y_out[i] = w->cinv * w->E[i] * w->y[i];
  csc_free(&w->P); csc_free(&w->A); kkt_free(&w->K); ldl_free(&w->F);
  free(w->q); free(w->l); free(w->u); free(w->D); free(w->E); free(w->Dinv); free(w->Einv);
  free(w->rho_vec); free(w->rho_inv_neg); free(w->x); free(w->z); free(w->y); free(w->x_prev);
  free(w->z_prev); free(w->xz_tilde); free(w->delta_y); free(w->delta_x); free(w->Ax); free(w->Px);
  free(w->Aty); free(w->tmp_n); free(w->tmp_m); free(w->constr_type);
  return status;
}
