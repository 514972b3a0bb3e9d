/* oracle/osqp_restate.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * Interface of the OSQP 0.5.0 restatement in osqp_restate.c (see that file's header). */
#ifndef SPECTRAL_ORACLE_OSQP_RESTATE_H
#define SPECTRAL_ORACLE_OSQP_RESTATE_H
#ifdef __cplusplus
extern "C" {
#endif

#define OSQP_RESTATE_SOLVED 1
#define OSQP_RESTATE_SOLVED_INACCURATE 2
#define OSQP_RESTATE_PRIMAL_INFEASIBLE_INACCURATE 3
#define OSQP_RESTATE_DUAL_INFEASIBLE_INACCURATE 4
#define OSQP_RESTATE_MAX_ITER_REACHED (-2)
#define OSQP_RESTATE_PRIMAL_INFEASIBLE (-3)
#define OSQP_RESTATE_DUAL_INFEASIBLE (-4)
#define OSQP_RESTATE_NON_CVX (-7)
#define OSQP_RESTATE_UNSOLVED (-10)

typedef struct {
  double rho, sigma;
  long long scaling;
  long long adaptive_rho, adaptive_rho_interval;
  double adaptive_rho_tolerance;
  long long max_iter;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, alpha, delta;
  long long polish, polish_refine_iter, scaled_termination, check_termination;
  long long polish_rounds; /* >1: active-set correction rounds with KKT verification (extension) */
} OsqpRestateSettings;

typedef struct {
  int status, iter, rho_updates, polish_status /* 0 none, 1 accepted, 2 accepted + KKT-verified, -1 rejected */, n_active;
  double obj_val, pri_res, dua_res, rho_final;
} OsqpRestateInfo;

void osqp_restate_default_settings(OsqpRestateSettings *s);

/* P: n x n upper-triangular CSC; A: m x n CSC.  x_out[n], y_out[m] (may be NULL). */
int osqp_restate_solve(long long n, long long m, const long long *Pp, const long long *Pi,
                       const double *Px, const double *q, const long long *Ap,
                       const long long *Ai, const double *Ax, const double *l, const double *u,
                       const OsqpRestateSettings *s, double *x_out, double *y_out,
                       OsqpRestateInfo *info);
#ifdef __cplusplus
}
#endif
#endif
