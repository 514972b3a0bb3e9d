/*
 * oracle/shim_include/osqp/osqp.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal restatement of the OSQP 0.5.0 public ABI (DLONG: c_int = long long,
 * DFLOAT off: c_float = double) -- exactly the five entry points and the struct
 * fields the reference touches:
 *   solve_3d.cc:1211-1228 (OSQPData, csc_matrix), :1235-1243 (settings fields),
 *   :1246-1277 (osqp_setup / osqp_solve / info->status_val / obj_val / solution),
 *   :1446-1462 (osqp_set_default_settings + overrides),
 *   piecewise_jerk_problem.cc:173-185 (FreeData: arrays of data->P / data->A).
 * OSQP itself is a third-party dependency that is NOT vendored in the reference
 * (makefile:2 links -losqp; version pinned only by the comment at solve_3d.cc:1246
 * "osqp-0.4.1, 0.5.0").  Field order / offsets follow SURVEY.md Appendix C.1, which
 * were recovered from the shipped libtrp.so/libcub.so, so the same header serves
 * (a) the shim libosqp.so that the SHIPPED binaries load and (b) the recompile of
 * the reference sources into oracle/_ref/.
 */
#ifndef SPECTRAL_ORACLE_OSQP_H
#define SPECTRAL_ORACLE_OSQP_H

#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef long long c_int;
typedef double c_float;

#define c_malloc malloc
#define c_calloc calloc
#define c_free free

typedef struct {
  c_int nzmax; /* 0x00 */
  c_int m;     /* 0x08 rows */
  c_int n;     /* 0x10 cols */
  c_int *p;    /* 0x18 column pointers (n+1) */
  c_int *i;    /* 0x20 row indices */
  c_float *x;  /* 0x28 values */
  c_int nz;    /* 0x30 -1 for CSC */
} csc;

typedef struct {
  c_int n;
  c_int m;
  csc *P; /* upper triangular */
  csc *A;
  c_float *q;
  c_float *l;
  c_float *u;
} OSQPData;

enum linsys_solver_type { QDLDL_SOLVER, MKL_PARDISO_SOLVER };

typedef struct {
  c_float rho;                    /* 0x00 */
  c_float sigma;                  /* 0x08 */
  c_int scaling;                  /* 0x10 */
  c_int adaptive_rho;             /* 0x18 */
  c_int adaptive_rho_interval;    /* 0x20 */
  c_float adaptive_rho_tolerance; /* 0x28 */
  c_float adaptive_rho_fraction;  /* 0x30 */
  c_int max_iter;                 /* 0x38 */
  c_float eps_abs;                /* 0x40 */
  c_float eps_rel;                /* 0x48 */
  c_float eps_prim_inf;           /* 0x50 */
  c_float eps_dual_inf;           /* 0x58 */
  c_float alpha;                  /* 0x60 */
  enum linsys_solver_type linsys_solver; /* 0x68 (+4 pad) */
  c_float delta;                  /* 0x70 */
  c_int polish;                   /* 0x78 */
  c_int polish_refine_iter;       /* 0x80 */
  c_int verbose;                  /* 0x88 */
  c_int scaled_termination;       /* 0x90 */
  c_int check_termination;        /* 0x98 */
  c_int warm_start;               /* 0xa0 */
  c_float time_limit;             /* 0xa8 */
} OSQPSettings;                   /* sizeof == 0xb0 */

typedef struct {
  c_int iter;          /* 0x00 */
  char status[32];     /* 0x08 */
  c_int status_val;    /* 0x28 */
  c_int status_polish; /* 0x30 */
  c_float obj_val;     /* 0x38 */
  c_float pri_res;
  c_float dua_res;
  c_float setup_time;
  c_float solve_time;
  c_float polish_time;
  c_float run_time;
  c_int rho_updates;
  c_float rho_estimate;
} OSQPInfo;

typedef struct {
  c_float *x;
  c_float *y;
} OSQPSolution;

typedef struct {
  void *opaque[23];       /* 0x00 .. 0xb0: data, linsys_solver, pol, rho_vec, ... */
  OSQPSettings *settings; /* 0xb8 */
  void *scaling;          /* 0xc0 */
  OSQPSolution *solution; /* 0xc8 */
  OSQPInfo *info;         /* 0xd0 */
  void *shim_state;       /* beyond what the reference reads */
} OSQPWorkspace;

#define OSQP_SOLVED 1
#define OSQP_SOLVED_INACCURATE 2
#define OSQP_MAX_ITER_REACHED (-2)
#define OSQP_PRIMAL_INFEASIBLE (-3)
#define OSQP_DUAL_INFEASIBLE (-4)
#define OSQP_PRIMAL_INFEASIBLE_INACCURATE 3
#define OSQP_DUAL_INFEASIBLE_INACCURATE 4
#define OSQP_NON_CVX (-7)
#define OSQP_UNSOLVED (-10)

csc *csc_matrix(c_int m, c_int n, c_int nzmax, c_float *x, c_int *i, c_int *p);
void osqp_set_default_settings(OSQPSettings *settings);
OSQPWorkspace *osqp_setup(const OSQPData *data, OSQPSettings *settings);
c_int osqp_solve(OSQPWorkspace *work);
c_int osqp_cleanup(OSQPWorkspace *work);

#ifdef __cplusplus
}
#endif
#endif
