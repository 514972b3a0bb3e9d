/*
 * oracle/spectral_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the reference's planning hot path:
 * corridor generation / split / selection, QP assembly, Bezier sampling and the wrapper
 * cost.  Every function in spectral_oracle.c cites the reference file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (spectral_b200/, libspectral.so) never does.
 */
#ifndef SPECTRAL_ORACLE_H
#define SPECTRAL_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_TRP 0
#define ORACLE_CUB 1

/* Same field order / size (112 B) as the reference's struct Cube, include/btrapz/cube_type.h:2-24 */
typedef struct {
  int beg_t, end_t;
  double t;
  double t_dif;
  double beg_l, end_l;
  double upp_skew, upp_bias, down_skew, down_bias;
  double l_upp_skew, l_upp_bias, l_down_skew, l_down_bias;
  unsigned char merge, split;
  int count;
} OracleCube;

#define ORACLE_OK 0
#define ORACLE_SOLVED_INACCURATE 1   /* OSQP status 2: accepted by the reference */
#define ORACLE_FAIL_NO_CORRIDOR 2    /* CollisionCheck selected nothing: reference is UB (solve_3d.cc:617) */
#define ORACLE_FAIL_SOLVER 3         /* Optimize returned false (solve_3d.cc:1253-1277) */
#define ORACLE_FAIL_POINTS_CHECK 4   /* CHECK_EQ(var_index, num_of_points_) would abort (solve_3d.cc:1407) */
#define ORACLE_FAIL_TOO_MANY 5       /* more segments than the caller's k_max */

/* a2 + a3: one region -> cubes.  xb, yb: [n][2] (lo, hi).  Returns cube count (<= cap) or -1. */
int oracle_corridor_generation(int variant, int n, double delta, const double *xb,
                               const double *yb, OracleCube *out, int cap);

/* a4: select / dedupe / order / de-overlap.  corridors: [R][cap].  Returns K (0 = nothing selected). */
int oracle_collision_check(int variant, int R, const OracleCube *corridors, const int *counts,
                           int cap, int n, double delta, const double *s_ref,
                           const double *l_ref, OracleCube *out, int out_cap);

/* libstdc++ std::sort restated (introsort + final insertion sort) on cubes keyed by beg_t */
void oracle_std_sort_by_beg_t(OracleCube *a, int n);

typedef struct {
  int n, m;          /* n = 12K, m = 42K */
  long long *P_p, *P_i; double *P_x;   /* upper triangular CSC */
  long long *A_p, *A_i; double *A_x;
  double *q, *l, *u;
} OracleQP;
void oracle_qp_free(OracleQP *qp);

typedef struct {
  int n_knots; double delta;
  double init_s[3], init_l[3];
  double ds_ref, dl_ref;
  double dds_lo, dds_hi, ddds_lo, ddds_hi, ddl_lo, ddl_hi, dddl_lo, dddl_hi;
  const double *ds_bounds;  /* [n][2] */
  const double *dl_bounds;  /* [n][2] */
  const double *s_ref, *l_ref; /* [n]; ref[j >= n] is read as 0.0 */
  double w[10]; /* Params order: s_acc, s_jerk, l_acc, l_jerk, s_ref, ds_ref, l_ref, dl_ref, end_s, end_l */
} OracleProblem;

/* a5-a8: assemble the QP exactly as FormulateProblem does */
int oracle_formulate(int variant, const OracleProblem *p, const OracleCube *segs, int K,
                     OracleQP *qp);

/* a9 (sampling part): control points -> samples.  out: [cap][6] = s,ds,dds,l,dl,ddl.
 * Returns the number of points, or -1 when CHECK_EQ(var_index, num_of_points_) would abort. */
int oracle_sample(const OracleProblem *p, const OracleCube *segs, int K, const double *ctrl,
                  double *out, int cap);

/* a10: the wrapper's cost over the samples (trp_wrapper.cpp:217-286 / cub_wrapper.cpp:210-258) */
double oracle_cost(int variant, const OracleProblem *p, const double *samples, int npts);

/* objective 0.5 x'Px + q'x of the assembled QP */
double oracle_qp_objective(const OracleQP *qp, const double *x);

typedef struct {
  int status, K, iters, npts, polish_status;
  double obj, a_cost;
} OracleResult;

/* whole path for one scenario, in memory.  mode 0: the reference's OSQP settings;
 * mode 1: converged optimum (tight eps + polish).  segs[k_max], ctrl[12*k_max], samples[cap][6]. */
int oracle_find_traj_mem(int variant, const OracleProblem *p, int R, const double *s_bounds,
                         const double *l_bounds, int mode, int k_max, OracleCube *segs,
                         double *ctrl, double *samples, int samples_cap, OracleResult *res);

/* batch version used by tests and by bench.py's CPU baseline ("port"); nthreads via OpenMP */
int oracle_solve_batch(int variant, int B, int N, int R, double delta, const double *s_bounds,
                       const double *l_bounds, const double *ds_bounds, const double *dl_bounds,
                       const double *s_ref, const double *l_ref, const double *init,
                       const double *scalars, const double *weights, int weights_stride,
                       int mode, int k_max, int nthreads, int *K, OracleCube *segs, double *ctrl,
                       double *obj, double *a_cost, int *status, int *iters, int *npts,
                       double *samples, int samples_cap, int *polish /* 2 = KKT-verified optimum */);

#ifdef __cplusplus
}
#endif
#endif
