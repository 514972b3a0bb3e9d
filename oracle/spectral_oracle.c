/*
 * oracle/spectral_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the reference's planning hot path.  Citations are to
 * /root/reference/src/solve_3d.cc (trapezoid-prism variant, "trp") and
 * /root/reference/src/cuboid_3d.cc (cuboid variant, "cub"); where only one is cited the other
 * is textually identical for that step.  Floating point: compile with -ffp-contract=off; every
 * expression keeps the reference's operand order so the corridor stage and the l/u/A data are
 * bit-exact.
 *
 * PINNING (see tests/test_oracle_*.py, tests/golden/): this restatement is checked against
 *   (1) QPs, corridors and trajectories captured from the reference's SHIPPED libtrp.so /
 *       libcub.so on every fixture (oracle/gen_golden.py, run in the build container),
 *   (2) the reference's own sources recompiled into oracle/_ref/ (oracle/Makefile),
 *   (3) the shipped output files s1_slt_3d_31.txt / s1_cub_3d_31.txt (3 decimals).
 * The QP *solution* beyond 3 decimals is "parity unpinned" by the reference itself (OSQP is
 * not vendored and stops at 1e-5 scaled residuals); it is anchored on the unique optimum.
 */
#include "spectral_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "osqp_restate.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ Cube() defaults
 * include/btrapz/cube_type.h:12-21 */
static void cube_init(OracleCube *c) {
  memset(c, 0, sizeof(*c));
  c->upp_bias = 1000.0;
  c->l_upp_bias = 1000.0;
}

/* ------------------------------------------------------------------ a3 CorridorSplit
 * solve_3d.cc:729-772 (trp copies the l_* faces, :750-759) ; cuboid_3d.cc:588-625 (no l_*) */
static int corridor_split(int variant, OracleCube *c, int count, int cap) {
  int temp_num = count;
  for (int k = 0; k < temp_num; k++) {
    while (c[k].t > 1) {
      if (count >= cap) return -1;
      c[k].t = c[k].t - 1;
      OracleCube m;
      cube_init(&m);
      m.beg_t = c[k].beg_t;
      c[k].beg_t = c[k].beg_t + 10;
      m.end_t = m.beg_t + 10;
      m.t = 1.0;
      m.down_skew = c[k].down_skew;
      m.down_bias = c[k].down_bias;
      if (variant == ORACLE_TRP) {
        m.l_down_skew = c[k].l_down_skew;
        m.l_down_bias = c[k].l_down_bias;
      }
      c[k].down_bias = m.down_bias + 1.0 * m.down_skew;
      m.upp_skew = c[k].upp_skew;
      m.upp_bias = c[k].upp_bias;
      if (variant == ORACLE_TRP) {
        m.l_upp_skew = c[k].l_upp_skew;
        m.l_upp_bias = c[k].l_upp_bias;
      }
      m.beg_l = c[k].beg_l;
      m.end_l = c[k].end_l;
      c[k].upp_bias = m.upp_bias + 1.0 * m.upp_skew;
      /* corridor.insert(corridor.begin() + k, mcube) */
      memmove(&c[k + 1], &c[k], (size_t)(count - k) * sizeof(OracleCube));
      c[k] = m;
      count++;
      temp_num++;
      k++;
    }
  }
  return count;
}

/* ------------------------------------------------------------------ a2 CorridorGeneration
 * solve_3d.cc:323-486 ; cuboid_3d.cc:301-407 */
int oracle_corridor_generation(int variant, int n, double delta, const double *xb,
                               const double *yb, OracleCube *out, int cap) {
#define XLO(i) xb[2 * (i)]
#define XHI(i) xb[2 * (i) + 1]
#define YLO(i) yb[2 * (i)]
#define YHI(i) yb[2 * (i) + 1]
  int j = 0;
  if (cap < 1) return -1;
  {
    OracleCube m;
    cube_init(&m);
    m.beg_t = 0;
    m.down_skew = (XLO(1) - XLO(0)) / delta; /* :332 */
    m.down_bias = XLO(0);
    m.upp_skew = (XHI(1) - XHI(0)) / delta;
    m.upp_bias = XHI(0);
    if (variant == ORACLE_TRP) { /* :338-341 */
      m.l_down_skew = (YLO(1) - YLO(0)) / delta;
      m.l_down_bias = YLO(0);
      m.l_upp_skew = (YHI(1) - YHI(0)) / delta;
      m.l_upp_bias = YHI(0);
    }
    m.beg_l = YLO(0);
    m.end_l = YHI(0);
    out[j++] = m;
  }
  for (int i = 2; i < n - 1; i++) { /* :349 */
    OracleCube m;
    cube_init(&m);
    double dskew = (XLO(i) - XLO(i - 1)) / delta; /* :355-356 */
    double uskew = (XHI(i) - XHI(i - 1)) / delta;
    if (variant == ORACLE_TRP) { /* :360-367: backward differences, recorded before the test */
      double l_dskew = (YLO(i) - YLO(i - 1)) / delta;
      double l_uskew = (YHI(i) - YHI(i - 1)) / delta;
      m.l_down_bias = YLO(i);
      m.l_upp_bias = YHI(i);
      m.l_down_skew = l_dskew;
      m.l_upp_skew = l_uskew;
    }
    double mthre = 0.2; /* :372 */
    if ((fabs(dskew - out[j - 1].down_skew) > mthre) || (fabs(uskew - out[j - 1].upp_skew) > mthre)) {
      if (j >= cap) return -1;
      out[j - 1].end_t = i;
      m.beg_t = i;
      m.down_skew = (XLO(i + 1) - XLO(i)) / delta; /* :380-386 */
      m.down_bias = XLO(i);
      m.upp_skew = (XHI(i + 1) - XHI(i)) / delta;
      m.upp_bias = XHI(i);
      m.beg_l = YLO(i);
      m.end_l = YHI(i);
      out[j++] = m;
    }
  }
  out[j - 1].end_t = n - 1;                                          /* :461 */
  for (int i = 0; i < j; i++) out[i].t = (out[i].end_t - out[i].beg_t) * delta; /* :465 */
  return corridor_split(variant, out, j, cap);                      /* :473 (Merge is a no-op :774) */
#undef XLO
#undef XHI
#undef YLO
#undef YHI
}

/* ------------------------------------------------------------------ libstdc++ std::sort
 * The reference calls std::sort(temp.begin(), temp.end(), beg_t <) at solve_3d.cc:630.  The
 * result for equal keys depends on the library's algorithm, so it is restated here: GNU
 * libstdc++ bits/stl_algo.h introsort (threshold 16, median-of-3 to first, depth 2*lg n with a
 * heap-sort fallback) followed by the final insertion sort. */
static int cube_less(const OracleCube *a, const OracleCube *b) { return a->beg_t < b->beg_t; }
static void cube_swap(OracleCube *a, OracleCube *b) { OracleCube t = *a; *a = *b; *b = t; }

static void ss_unguarded_linear_insert(OracleCube *last) {
  OracleCube val = *last;
  OracleCube *next = last - 1;
  while (cube_less(&val, next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void ss_insertion_sort(OracleCube *first, OracleCube *last) {
  if (first == last) return;
  for (OracleCube *i = first + 1; i != last; ++i) {
    if (cube_less(i, first)) {
      OracleCube val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(OracleCube));
      *first = val;
    } else ss_unguarded_linear_insert(i);
  }
}
static void ss_push_heap(OracleCube *first, long hole, long top, OracleCube value) {
  long parent = (hole - 1) / 2;
  while (hole > top && cube_less(first + parent, &value)) {
    first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
static void ss_adjust_heap(OracleCube *first, long hole, long len, OracleCube value) {
  const long top = hole;
  long second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (cube_less(first + second, first + (second - 1))) second--;
    first[hole] = first[second]; hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1]; hole = second - 1;
  }
  ss_push_heap(first, hole, top, value);
}
static void ss_heap_sort(OracleCube *first, OracleCube *last) {
  long len = last - first;
  if (len >= 2) {
    long parent = (len - 2) / 2;
    for (;;) {
      OracleCube v = first[parent];
      ss_adjust_heap(first, parent, len, v);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    OracleCube v = *last;
    *last = *first;
    ss_adjust_heap(first, 0, last - first, v);
  }
}
static void ss_move_median_to_first(OracleCube *result, OracleCube *a, OracleCube *b, OracleCube *c) {
  if (cube_less(a, b)) {
    if (cube_less(b, c)) cube_swap(result, b);
    else if (cube_less(a, c)) cube_swap(result, c);
    else cube_swap(result, a);
  } else if (cube_less(a, c)) cube_swap(result, a);
  else if (cube_less(b, c)) cube_swap(result, c);
  else cube_swap(result, b);
}
static OracleCube *ss_unguarded_partition(OracleCube *first, OracleCube *last, OracleCube *pivot) {
  for (;;) {
    while (cube_less(first, pivot)) ++first;
    --last;
    while (cube_less(pivot, last)) --last;
    if (!(first < last)) return first;
    cube_swap(first, last);
    ++first;
  }
}
static void ss_introsort_loop(OracleCube *first, OracleCube *last, long depth) {
  while (last - first > 16) {
    if (depth == 0) { ss_heap_sort(first, last); return; }
    --depth;
    OracleCube *mid = first + (last - first) / 2;
    ss_move_median_to_first(first, first + 1, mid, last - 1);
    OracleCube *cut = ss_unguarded_partition(first + 1, last, first);
    ss_introsort_loop(cut, last, depth);
    last = cut;
  }
}
void oracle_std_sort_by_beg_t(OracleCube *a, int n) {
  if (n <= 0) return;
  long lg = 0;
  for (long v = n; v > 1; v >>= 1) lg++;
  ss_introsort_loop(a, a + n, lg * 2);
  if (n > 16) {
    ss_insertion_sort(a, a + 16);
    for (OracleCube *i = a + 16; i != a + n; ++i) ss_unguarded_linear_insert(i);
  } else ss_insertion_sort(a, a + n);
}

/* ------------------------------------------------------------------ a4 CollisionCheck
 * solve_3d.cc:488-714 ; cuboid_3d.cc:409-573 */
static int point_inside(const OracleCube *c, double s, double l, double tt, double delta) {
  int pos = 0, neg = 0;
  double d;
  if (!(l <= c->end_l && l >= c->beg_l)) return 0; /* :534 */
  /* the degenerate (beg_t - beg_t) / (end_t - end_t) factors are the reference's own (:536,:559) */
  d = (s - c->down_bias) * (c->beg_t - c->beg_t) - (tt - c->beg_t) * (c->upp_bias - c->down_bias);
  if (d > 0) pos++;
  if (d < 0) neg++;
  if (pos > 0 && neg > 0) return 0;
  d = (s - c->upp_bias) * (c->end_t - c->beg_t) -
      (tt - c->beg_t) * (c->upp_skew * delta + c->upp_bias - c->upp_bias); /* :547 */
  if (d > 0) pos++;
  if (d < 0) neg++;
  if (pos > 0 && neg > 0) return 0;
  d = (s - c->upp_bias - c->upp_skew * delta) * (c->end_t - c->end_t) -
      (tt - c->end_t) * (c->down_skew * delta + c->down_bias - c->upp_skew * delta - c->upp_bias); /* :559 */
  if (d > 0) pos++;
  if (d < 0) neg++;
  if (pos > 0 && neg > 0) return 0;
  d = (s - c->down_bias - c->down_skew * delta) * (c->beg_t - c->end_t) -
      (tt - c->end_t) * (c->down_bias - c->down_skew * delta - c->down_bias); /* :571 */
  if (d > 0) pos++;
  if (d < 0) neg++;
  if (pos > 0 && neg > 0) return 0;
  return 1;
}

int oracle_collision_check(int variant, int R, const OracleCube *corridors, const int *counts,
                           int cap, int n, double delta, const double *s_ref,
                           const double *l_ref, OracleCube *out, int out_cap) {
  /* temp: every third inside-point pushes a copy of the cube it fell into; `count` is carried
   * across cubes and regions (:524-596; the `if (count > 2) k++` at :595 is dead code because
   * count was just reset, and the guards at :601-610 never fire). */
  int tcap = 64, tn = 0;
  OracleCube *temp = (OracleCube *)malloc((size_t)tcap * sizeof(OracleCube));
  int count = 0;
  for (int j = 0; j < R; j++) {
    for (int k = 0; k < counts[j]; k++) {
      const OracleCube *c = &corridors[(size_t)j * cap + k];
      for (int i = 0; i < n; i++) {
        if (point_inside(c, s_ref[i], l_ref[i], (double)i, delta)) {
          count++;
          if (count > 2) {
            if (tn == tcap) { tcap *= 2; temp = (OracleCube *)realloc(temp, (size_t)tcap * sizeof(OracleCube)); }
            temp[tn] = *c;
            temp[tn].count = count;
            tn++;
            count = 0;
          }
        }
      }
    }
  }
  if (tn == 0) { free(temp); return 0; } /* reference: temp.size()-1 underflows (:617) -> UB */

  /* exact-equality dedupe on 8 fields (:617-628) */
  for (int i = 0; i < tn - 1; i++) {
    for (int j = i + 1; j < tn; j++) {
      if (temp[i].beg_t == temp[j].beg_t && temp[i].end_t == temp[j].end_t &&
          temp[i].down_bias == temp[j].down_bias && temp[i].down_skew == temp[j].down_skew &&
          temp[i].upp_bias == temp[j].upp_bias && temp[i].upp_skew == temp[j].upp_skew &&
          temp[i].beg_l == temp[j].beg_l && temp[i].end_l == temp[j].end_l) {
        memmove(&temp[j], &temp[j + 1], (size_t)(tn - j - 1) * sizeof(OracleCube));
        tn--;
        j--;
      }
    }
  }

  if (variant == ORACLE_TRP) {
    oracle_std_sort_by_beg_t(temp, tn); /* :630 */
    /* lateral-continuity swap pass (:639-673) */
    for (int i = 0; i < tn - 1; i++) {
      for (int j = i + 1; j < tn; j++) {
        if (temp[i].beg_l == temp[j].beg_l && j - i == 1) break;
        for (int k = j + 1; k < tn; k++) {
          if (temp[i].beg_l == temp[k].beg_l && temp[i].end_t == temp[k].beg_t) {
            cube_swap(&temp[j], &temp[k]);
            break;
          }
        }
      }
    }
    /* de-overlap, only j = i+1 is examined because of the unconditional break (:678-703) */
    for (int i = 0; i < tn - 1; i++) {
      for (int j = i + 1; j < tn; j++) {
        if (temp[i].beg_t == temp[j].beg_t && temp[i].end_t == temp[j].end_t) {
          int diff = (temp[i].end_t - temp[i].beg_t) / 2;
          temp[i].end_t = temp[i].end_t - diff;
          temp[i].t = (temp[i].end_t - temp[i].beg_t) * delta;
          temp[j].beg_t = temp[j].beg_t + diff;
          temp[j].t = (temp[j].end_t - temp[j].beg_t) * delta;
        } else if (temp[i].beg_t > temp[j].beg_t && temp[i].end_t <= temp[j].end_t) {
          int diff = (temp[i].end_t - temp[i].beg_t) / 2;
          if (diff > 1) {
            temp[i].end_t = temp[i].end_t - diff;
            temp[i].t = (temp[i].end_t - temp[i].beg_t) * delta;
          }
          temp[j].beg_t = temp[i].end_t;
          temp[j].t = (temp[j].end_t - temp[j].beg_t) * delta;
        }
        break;
      }
    }
  } else {
    /* cub: no sort, no swap pass; every pair, diff = /3, no containment branch (cuboid_3d.cc:553-567) */
    for (int i = 0; i < tn - 1; i++) {
      for (int j = i + 1; j < tn; j++) {
        if (temp[i].beg_t == temp[j].beg_t && temp[i].end_t == temp[j].end_t) {
          int diff = (temp[i].end_t - temp[i].beg_t) / 3;
          temp[i].end_t = temp[i].end_t - diff;
          temp[i].t = (temp[i].end_t - temp[i].beg_t) * delta;
          temp[j].beg_t = temp[j].beg_t + diff;
          temp[j].t = (temp[j].end_t - temp[j].beg_t) * delta;
        }
      }
    }
  }
  int K = tn;
  if (K > out_cap) { free(temp); return -K; }
  memcpy(out, temp, (size_t)K * sizeof(OracleCube));
  free(temp);
  return K;
}

/* ------------------------------------------------------------------ 6x6 helpers (Eigen semantics)
 * Eigen evaluates these small dynamic products coefficient-wise, k = 0..5 left to right,
 * except Transpose * Matrix which reduces with 2-wide packets
 * (see oracle/shim_include/eigen3/Eigen/Dense; both verified against the shipped binaries' P). */
static const double kM[6][6] = { /* solve_3d.cc:122-127 */
    {1, 0, 0, 0, 0, 0},      {-5, 5, 0, 0, 0, 0},      {10, -20, 10, 0, 0, 0},
    {-10, 30, -30, 10, 0, 0}, {5, -20, 30, -20, 5, 0}, {-1, 5, -10, 10, -5, 1}};

/* Transpose * Matrix: SSE2 packet reduction, even / odd terms accumulate separately */
static void mat6_mul_packet(const double a[6][6], const double b[6][6], double out[6][6]) {
  for (int j = 0; j < 6; j++)
    for (int i = 0; i < 6; i++) {
      double e = a[i][0] * b[0][j], o = a[i][1] * b[1][j];
      e = e + a[i][2] * b[2][j]; o = o + a[i][3] * b[3][j];
      e = e + a[i][4] * b[4][j]; o = o + a[i][5] * b[5][j];
      out[i][j] = e + o;
    }
}

static void mat6_mul(const double a[6][6], const double b[6][6], double out[6][6]) {
  for (int j = 0; j < 6; j++)
    for (int i = 0; i < 6; i++) {
      double s = a[i][0] * b[0][j];
      for (int k = 1; k < 6; k++) s = s + a[i][k] * b[k][j];
      out[i][j] = s;
    }
}

/* inv_M(i,1) as Eigen's PartialPivLU::inverse() produces it for kM (solve_3d.cc:813); the bit
 * patterns are those of the shipped binaries and are re-derived from
 * the LU restatement by tests/test_oracle_ref.py. */
static const double kInvMCol1[6] = {0x1.999999999999ap-52, 0x1.99999999999a4p-3, 0x1.999999999999cp-2,
                                    0x1.3333333333334p-1,  0x1.999999999999ap-1, 0x1.0p+0};

/* CalculateKernel's MQM tables: solve_3d.cc:79-143 */
static void build_mqm(double w_ref, double w_dref, double w_dd, double w_ddd, double MQM[4][6][6]) {
  double pQp[4][6][6];
  memset(pQp, 0, sizeof(pQp));
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      pQp[0][i][j] = (double)(w_ref) / (i + j + 1);
      if (i >= 1 && j >= 1) pQp[1][i][j] = (double)(w_dref * i * j) / (i + j - 1);
      if (i >= 2 && j >= 2) pQp[2][i][j] = (double)(w_dd * i * j * (i - 1) * (j - 1)) / (i + j - 3);
      if (i >= 3 && j >= 3)
        pQp[3][i][j] = (double)(w_ddd * i * j * (i - 1) * (j - 1) * (i - 2) * (j - 2)) / (i + j - 5);
    }
  double Mt[6][6], tmp[6][6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) Mt[i][j] = kM[j][i];
  for (int k = 0; k < 4; k++) { /* (M' * pQp) * M */
    mat6_mul_packet(Mt, pQp[k], tmp);
    mat6_mul(tmp, kM, MQM[k]);
  }
}

static double ref_at(const double *ref, int n, int j) { return j < n ? ref[j] : 0.0; }

void oracle_qp_free(OracleQP *qp) {
  free(qp->P_p); free(qp->P_i); free(qp->P_x); free(qp->A_p); free(qp->A_i); free(qp->A_x);
  free(qp->q); free(qp->l); free(qp->u);
  memset(qp, 0, sizeof(*qp));
}

static double dmax(double a, double b) { return (a < b) ? b : a; } /* std::max(a, b) */
static double dmin(double a, double b) { return (b < a) ? b : a; } /* std::min(a, b) */

typedef struct { int row, col; double val; } Trip;

/* ------------------------------------------------------------------ a5-a8 FormulateProblem
 * solve_3d.cc:1143-1229 -> CalculateKernel :70-224, CalculateAffineConstraint :779-1129,
 * CalculateOffset :226-321 ; cuboid_3d.cc:1002-1088, :67-219, :632-988, :221-299 */
int oracle_formulate(int variant, const OracleProblem *p, const OracleCube *segs, int K,
                     OracleQP *qp) {
  const int np = 6, order = 5;
  const int nv = K * np;       /* per axis */
  const int n = 2 * nv;
  const int m = 2 * (K * (3 * np - 3 + np - 3) + 3 + 3 * (K - 1)); /* :785 */
  const int N = p->n_knots;
  const double delta = p->delta;
  memset(qp, 0, sizeof(*qp));
  qp->n = n; qp->m = m;
  const double w_s_acc = p->w[0], w_s_jerk = p->w[1], w_l_acc = p->w[2], w_l_jerk = p->w[3];
  const double w_s_ref = p->w[4], w_ds_ref = p->w[5], w_l_ref = p->w[6], w_dl_ref = p->w[7];
  const double w_end_s = p->w[8], w_end_l = p->w[9];

  /* x/y skew & bias from ref[10k], ref[10k+1] regardless of beg_t (:1159-1166) */
  double *x_skew = (double *)malloc(sizeof(double) * K * 4);
  double *x_bias = x_skew + K, *y_skew = x_skew + 2 * K, *y_bias = x_skew + 3 * K;
  for (int k = 0; k < K; k++) {
    x_skew[k] = (ref_at(p->s_ref, N, k * 10 + 1) - ref_at(p->s_ref, N, k * 10)) / delta;
    x_bias[k] = ref_at(p->s_ref, N, k * 10);
    y_skew[k] = (ref_at(p->l_ref, N, k * 10 + 1) - ref_at(p->l_ref, N, k * 10)) / delta;
    y_bias[k] = ref_at(p->l_ref, N, k * 10);
  }

  /* ---- P (:145-222): 2K upper-triangular 6x6 blocks, column by column */
  double MQMx[4][6][6], MQMy[4][6][6];
  build_mqm(w_s_ref, w_ds_ref, w_s_acc, w_s_jerk, MQMx);
  build_mqm(w_l_ref, w_dl_ref, w_l_acc, w_l_jerk, MQMy);
  qp->P_p = (long long *)malloc(sizeof(long long) * (n + 1));
  qp->P_i = (long long *)malloc(sizeof(long long) * 21 * 2 * K);
  qp->P_x = (double *)malloc(sizeof(double) * 21 * 2 * K);
  int idx = 0, sub_shift = 0, col = 0;
  for (int axis = 0; axis < 2; axis++) {
    double(*MQM)[6][6] = axis == 0 ? MQMx : MQMy;
    double w_end = axis == 0 ? w_end_s : w_end_l;
    for (int k = 0; k < K; k++) {
      double t = segs[k].t;
      for (int j = 0; j < np; j++) {
        qp->P_p[col++] = idx;
        for (int i = 0; i < np; i++) {
          if (j >= i) {
            double mval = pow(t, 3) * MQM[0][i][j] + t * MQM[1][i][j] + MQM[2][i][j] / t +
                          MQM[3][i][j] / pow(t, 3); /* :159-160 */
            if ((k == K - 1) && (i == np - 1) && (j == np - 1)) mval = mval + w_end * t * t; /* :167 */
            qp->P_i[idx] = sub_shift + i;
            qp->P_x[idx] = 2.0 * mval;
            idx++;
          }
        }
      }
      sub_shift += np;
    }
  }
  qp->P_p[col] = idx;

  /* ---- A, l, u (:779-1129) */
  double aval[3] = {1.0 * order * (order - 1), -2.0 * order * (order - 1), 1.0 * order * (order - 1)};
  double jval[4] = {-1.0 * order * (order - 1) * (order - 2), 3.0 * order * (order - 1) * (order - 2),
                    -3.0 * order * (order - 1) * (order - 2), 1.0 * order * (order - 1) * (order - 2)};
  qp->l = (double *)calloc(m + 1, sizeof(double));
  qp->u = (double *)calloc(m + 1, sizeof(double));
  Trip *tr = (Trip *)malloc(sizeof(Trip) * (size_t)(104 * K + 16));
  int nt = 0, ci = 0;
#define PUT(v, r, x) do { tr[nt].col = (v); tr[nt].row = (r); tr[nt].val = (x); nt++; } while (0)
  for (int axis = 0; axis < 2; axis++) {
    const int off = axis * nv;
    const double *init = axis == 0 ? p->init_s : p->init_l;
    int var_shift = 0;
    for (int k = 0; k < K; k++) {
      const OracleCube *c = &segs[k];
      /* containment rows */
      if (axis == 0 && variant == ORACLE_CUB) { /* cuboid_3d.cc:677-697 */
        double l_bound = 0, u_bound = 100;
        for (int i = 0; i < np; i++) {
          l_bound = dmax(l_bound, c->down_bias + c->down_skew * kInvMCol1[i] * c->t);
          u_bound = dmin(u_bound, c->upp_bias + c->upp_skew * kInvMCol1[i] * c->t);
        }
        for (int i = 0; i < np; i++) {
          PUT(off + var_shift + i, ci, 1.0 * c->t);
          qp->l[ci] = l_bound; qp->u[ci] = u_bound; ++ci;
        }
      } else {
        for (int i = 0; i < np; i++) {
          PUT(off + var_shift + i, ci, 1.0 * c->t);
          if (axis == 0) { /* :827-828 */
            qp->l[ci] = c->down_bias + c->down_skew * kInvMCol1[i] * c->t;
            qp->u[ci] = c->upp_bias + c->upp_skew * kInvMCol1[i] * c->t;
          } else if (variant == ORACLE_TRP) { /* :965-966 */
            qp->l[ci] = c->l_down_bias + c->l_down_skew * kInvMCol1[i] * c->t;
            qp->u[ci] = c->l_upp_bias + c->l_upp_skew * kInvMCol1[i] * c->t;
          } else { /* cuboid_3d.cc:826-827 */
            qp->l[ci] = c->beg_l; qp->u[ci] = c->end_l;
          }
          ++ci;
        }
      }
      /* physical bounds */
      double d_lo = 0.0, d_hi = 1000.0, dd_lo = -1000.0, dd_hi = 1000.0; /* :835-836 */
      if (axis == 0) {
        for (int i = c->beg_t; i <= c->end_t; i++) { /* :838-845 */
          d_lo = dmax(p->ds_bounds[2 * i], d_lo);
          d_hi = dmin(p->ds_bounds[2 * i + 1], d_hi);
          dd_lo = dmax(p->dds_lo, dd_lo); /* ddx_bounds_ is uniform (set_ddx_bounds(lo,hi), trp_wrapper.cpp:150) */
          dd_hi = dmin(p->dds_hi, dd_hi);
        }
      }
      for (int i = 0; i < np - 1; i++) { /* velocity rows :848-859 / l: :994-1007 */
        PUT(off + var_shift + i, ci, -1.0 * order);
        PUT(off + var_shift + i + 1, ci, 1.0 * order);
        if (axis == 0) { qp->l[ci] = d_lo; qp->u[ci] = d_hi; }
        else { qp->l[ci] = p->dl_bounds[2 * i]; qp->u[ci] = p->dl_bounds[2 * i + 1]; } /* dy_bounds_[i], control index */
        ++ci;
      }
      for (int i = 0; i < np - 2; i++) { /* acceleration rows :862-874 / l: :1010-1023 */
        PUT(off + var_shift + i, ci, aval[0]);
        PUT(off + var_shift + i + 1, ci, aval[1]);
        PUT(off + var_shift + i + 2, ci, aval[2]);
        if (axis == 0) { qp->l[ci] = dd_lo * c->t; qp->u[ci] = dd_hi * c->t; }
        else { qp->l[ci] = p->ddl_lo * c->t; qp->u[ci] = p->ddl_hi * c->t; } /* ddy_bounds_[i] uniform */
        ++ci;
      }
      for (int i = 0; i < np - 3; i++) { /* jerk rows :877-888 / l: :1026-1037 */
        PUT(off + var_shift + i, ci, jval[0]);
        PUT(off + var_shift + i + 1, ci, jval[1]);
        PUT(off + var_shift + i + 2, ci, jval[2]);
        PUT(off + var_shift + i + 3, ci, jval[3]);
        double jl = axis == 0 ? p->ddds_lo : p->dddl_lo, jh = axis == 0 ? p->ddds_hi : p->dddl_hi;
        qp->l[ci] = jl * c->t * c->t; qp->u[ci] = jh * c->t * c->t;
        ++ci;
      }
      var_shift += np;
    }
    /* initial state (:896-912) */
    PUT(off + 0, ci, 1.0 * segs[0].t);
    qp->l[ci] = init[0]; qp->u[ci] = init[0]; ++ci;
    PUT(off + 0, ci, -1.0 * order);
    PUT(off + 1, ci, 1.0 * order);
    qp->l[ci] = init[1]; qp->u[ci] = init[1]; ++ci;
    PUT(off + 0, ci, aval[0]);
    PUT(off + 1, ci, aval[1]);
    PUT(off + 2, ci, aval[2]);
    qp->l[ci] = init[2] * segs[0].t; qp->u[ci] = init[2] * segs[0].t; ++ci;
    /* joints (:918-949) */
    for (int k = 0; k < K - 1; k++) {
      int ss = (k + 1) * np;
      PUT(off + ss - 1, ci, -1.0 * segs[k].t);
      PUT(off + ss, ci, 1.0 * segs[k + 1].t);
      qp->l[ci] = 0.0; qp->u[ci] = 0.0; ++ci;
      PUT(off + ss - 2, ci, -1.0);
      PUT(off + ss - 1, ci, 1.0);
      PUT(off + ss, ci, 1.0);
      PUT(off + ss + 1, ci, -1.0);
      qp->l[ci] = 0.0; qp->u[ci] = 0.0; ++ci;
      PUT(off + ss - 3, ci, 1.0 * segs[k + 1].t);
      PUT(off + ss - 2, ci, -2.0 * segs[k + 1].t);
      PUT(off + ss - 1, ci, 1.0 * segs[k + 1].t);
      PUT(off + ss, ci, -1.0 * segs[k].t);
      PUT(off + ss + 1, ci, 2.0 * segs[k].t);
      PUT(off + ss + 2, ci, -1.0 * segs[k].t);
      qp->l[ci] = 0.0; qp->u[ci] = 0.0; ++ci;
    }
  }
#undef PUT
  if (ci != m) { free(tr); free(x_skew); return -1; } /* CHECK_EQ(constraint_index, num_of_constraints) :1107 */
  /* per-variable lists -> CSC (:1109-1128); insertion order == increasing row */
  qp->A_p = (long long *)calloc(n + 2, sizeof(long long));
  qp->A_i = (long long *)malloc(sizeof(long long) * (size_t)(nt + 1));
  qp->A_x = (double *)malloc(sizeof(double) * (size_t)(nt + 1));
  for (int e = 0; e < nt; e++) qp->A_p[tr[e].col + 1]++;
  for (int j = 0; j < n; j++) qp->A_p[j + 1] += qp->A_p[j];
  long long *fill = (long long *)malloc(sizeof(long long) * (size_t)(n + 1));
  memcpy(fill, qp->A_p, sizeof(long long) * (size_t)(n + 1));
  for (int e = 0; e < nt; e++) {
    long long pos = fill[tr[e].col]++;
    qp->A_i[pos] = tr[e].row;
    qp->A_x[pos] = tr[e].val;
  }
  free(fill);
  free(tr);

  /* ---- q (:226-321) */
  qp->q = (double *)calloc(n + 1, sizeof(double));
  for (int axis = 0; axis < 2; axis++) {
    const double w_ref = axis == 0 ? w_s_ref : w_l_ref;
    const double w_dref = axis == 0 ? w_ds_ref : w_dl_ref;
    const double dref = axis == 0 ? p->ds_ref : p->dl_ref;
    const double *skew = axis == 0 ? x_skew : y_skew, *bias = axis == 0 ? x_bias : y_bias;
    const double *ref = axis == 0 ? p->s_ref : p->l_ref;
    for (int k = 0; k < K; k++) {
      double t = segs[k].t, q_p[6];
      for (int i = 0; i < np; i++) {
        q_p[i] = 0.0;
        q_p[i] += -2.0 * pow(t, 3) * w_ref * skew[k] / (i + 2); /* :248 */
        q_p[i] += -2.0 * pow(t, 2) * w_ref * bias[k] / (i + 1); /* :249 */
        if (i > 0) q_p[i] += -2.0 * w_dref * dref * t;          /* :254 */
      }
      for (int j = 0; j < np; j++) { /* q_c = q_p * M (:258) */
        double s = q_p[0] * kM[0][j];
        for (int i = 1; i < np; i++) s = s + q_p[i] * kM[i][j];
        qp->q[axis * nv + k * np + j] = s;
      }
    }
    qp->q[axis * nv + nv - 1] -= dref * 2.0 * ref[N - 1] * segs[K - 1].t; /* :268 / :315 */
  }
  free(x_skew);
  return 0;
}

double oracle_qp_objective(const OracleQP *qp, const double *x) {
  double o = 0.0;
  for (int j = 0; j < qp->n; j++) {
    o += qp->q[j] * x[j];
    for (long long e = qp->P_p[j]; e < qp->P_p[j + 1]; e++) {
      long long i = qp->P_i[e];
      double v = qp->P_x[e] * x[i] * x[j];
      o += (i == j) ? 0.5 * v : v;
    }
  }
  return o;
}

/* ------------------------------------------------------------------ a9 sampling
 * solve_3d.cc:1279-1392,1407 */
int oracle_sample(const OracleProblem *p, const OracleCube *segs, int K, const double *ctrl,
                  double *out, int cap) {
  const int np = 6, order = 5, nv = 6 * K;
  const double delta = p->delta;
  int num_of_points = 1; /* solve_3d.h:114 */
  for (int i = 0; i < K; i++) num_of_points += segs[i].t / delta; /* int += double, :1281 */
  double factorial[6];
  factorial[0] = 1.0;
  for (int i = 1; i < np; i++) factorial[i] = factorial[i - 1] * i;
  double b_coe[6][3];
  memset(b_coe, 0, sizeof(b_coe));
  for (int i = 0; i < np; i++) b_coe[i][0] = factorial[order] / (factorial[i] * factorial[order - i]);
  for (int i = 0; i < np - 1; i++) b_coe[i][1] = factorial[order - 1] / (factorial[i] * factorial[order - 1 - i]);
  for (int i = 0; i < np - 2; i++) b_coe[i][2] = factorial[order - 2] / (factorial[i] * factorial[order - 2 - i]);
  int var_index = 0;
  if (cap < 1) return -2;
  out[0] = p->init_s[0]; out[1] = p->init_s[1]; out[2] = p->init_s[2];
  out[3] = p->init_l[0]; out[4] = p->init_l[1]; out[5] = p->init_l[2];
  var_index++;
  for (int k = 0; k < K; ++k) {
    double c[12];
    for (int i = 0; i < np; i++) { c[i] = ctrl[k * np + i]; c[i + np] = ctrl[k * np + i + nv]; }
    double t = segs[k].t;
    int linter = t / delta; /* :1351 */
    for (int l = 1; l <= linter; l++) {
      if (var_index >= cap) return -2;
      /* x_.at(var_index) throws std::out_of_range when var_index >= num_of_points_: the
       * reference would terminate; report it as the same failure class as the CHECK. */
      if (var_index >= num_of_points) return -1;
      double x = 0, dx = 0, ddx = 0, y = 0, dy = 0, ddy = 0;
      double u = (double)l / linter;
      for (int i = 0; i < np; i++) {
        x += c[i] * b_coe[i][0] * pow(u, i) * pow(1 - u, order - i);
        y += c[i + np] * b_coe[i][0] * pow(u, i) * pow(1 - u, order - i);
      }
      x = x * t; y = y * t;
      for (int i = 0; i < np - 1; i++) {
        dx += order * (c[i + 1] - c[i]) * b_coe[i][1] * pow(u, i) * pow(1 - u, order - 1 - i);
        dy += order * (c[i + 1 + np] - c[i + np]) * b_coe[i][1] * pow(u, i) * pow(1 - u, order - 1 - i);
      }
      for (int i = 0; i < np - 2; i++) {
        ddx += order * (order - 1) * (c[i + 2] - 2.0 * c[i + 1] + c[i]) * b_coe[i][2] * pow(u, i) * pow(1 - u, order - 2 - i);
        ddy += order * (order - 1) * (c[i + 2 + np] - 2.0 * c[i + 1 + np] + c[i + np]) * b_coe[i][2] * pow(u, i) * pow(1 - u, order - 2 - i);
      }
      ddx = ddx / t; ddy = ddy / t;
      double *o = out + 6 * var_index;
      o[0] = x; o[1] = dx; o[2] = ddx; o[3] = y; o[4] = dy; o[5] = ddy;
      var_index++;
    }
  }
  if (var_index != num_of_points) return -1; /* CHECK_EQ :1407 -> abort() */
  return num_of_points;
}

/* ------------------------------------------------------------------ a10 wrapper cost
 * trp: trp_wrapper.cpp:217-286.  cub: cub_wrapper.cpp:210-258, whose `double l_cost;` is
 * uninitialised (UB) -- defined here as 0.0.  trp's end term indexes l[num_of_knots-1]
 * (trp_wrapper.cpp:269), out of bounds when fewer than N samples exist: defined here on the
 * last available sample.  Both conventions are documented in DESIGN.md. */
double oracle_cost(int variant, const OracleProblem *p, const double *sm, int npts) {
  const double delta_t = p->delta;
  const int N = p->n_knots;
#define S_(i) sm[6 * (i) + 0]
#define DS_(i) sm[6 * (i) + 1]
#define DDS_(i) sm[6 * (i) + 2]
#define L_(i) sm[6 * (i) + 3]
#define DL_(i) sm[6 * (i) + 4]
#define DDL_(i) sm[6 * (i) + 5]
  double s_cost = 0.0, l_cost = 0.0, mmax_a = 0.0;
  if (npts < 2) return 0.0;
  for (int i = 0; i < npts; ++i) {
    double ddds = (i == 0) ? (DDS_(1) - DDS_(0)) / delta_t : (DDS_(i) - DDS_(i - 1)) / delta_t;
    double xr = ref_at(p->s_ref, N, i);
    if (variant == ORACLE_TRP) {
      s_cost += p->w[4] * (S_(i) - xr) * (S_(i) - xr) * delta_t;
      s_cost += p->w[5] * DS_(i) * DS_(i) * delta_t;
      s_cost += p->w[0] * DDS_(i) * DDS_(i) * delta_t;
      s_cost += p->w[1] * ddds * ddds * delta_t;
    } else {
      s_cost += (S_(i) - xr) * (S_(i) - xr) * delta_t;
      s_cost += DS_(i) * DS_(i) * delta_t;
      s_cost += DDS_(i) * DDS_(i) * DDS_(i) * DDS_(i) * delta_t;
      s_cost += ddds * ddds * ddds * ddds * delta_t;
    }
    mmax_a = dmax(mmax_a, fabs(DDS_(i)));
  }
  if (variant == ORACLE_CUB) s_cost += mmax_a * mmax_a * mmax_a * mmax_a;
  mmax_a = 0.0;
  for (int i = 0; i < npts; ++i) {
    double dddl = (i == 0) ? (DDL_(1) - DDL_(0)) / delta_t : (DDL_(i) - DDL_(i - 1)) / delta_t;
    double yr = ref_at(p->l_ref, N, i);
    if (variant == ORACLE_TRP) {
      l_cost += p->w[6] * (L_(i) - yr) * (L_(i) - yr) * delta_t;
      l_cost += p->w[7] * DL_(i) * DL_(i) * delta_t;
      l_cost += p->w[2] * DDL_(i) * DDL_(i) * delta_t;
      l_cost += p->w[3] * dddl * dddl * delta_t;
    } else {
      l_cost += (L_(i) - yr) * (L_(i) - yr) * delta_t;
      l_cost += DL_(i) * DL_(i) * delta_t;
      l_cost += DDL_(i) * DDL_(i) * delta_t;
      l_cost += dddl * dddl * delta_t;
    }
    mmax_a = dmax(mmax_a, fabs(DDL_(i)));
  }
  if (variant == ORACLE_TRP) {
    int last = (N - 1 < npts) ? N - 1 : npts - 1;
    l_cost += p->w[9] * (L_(last) - p->l_ref[N - 1]) * (L_(last) - p->l_ref[N - 1]) * delta_t;
  } else {
    l_cost += mmax_a * mmax_a;
  }
  return s_cost + l_cost;
#undef S_
#undef DS_
#undef DDS_
#undef L_
#undef DL_
#undef DDL_
}

/* ------------------------------------------------------------------ whole path, one scenario */
#define REGION_CAP 64

int oracle_find_traj_mem(int variant, const OracleProblem *p, int R, const double *s_bounds,
                         const double *l_bounds, int mode, int k_max, OracleCube *segs,
                         double *ctrl, double *samples, int samples_cap, OracleResult *res) {
  const int N = p->n_knots;
  memset(res, 0, sizeof(*res));
  res->a_cost = 100000000000.0; /* trp_wrapper.cpp:199 */
  OracleCube *corr = (OracleCube *)malloc(sizeof(OracleCube) * (size_t)R * REGION_CAP);
  int *counts = (int *)malloc(sizeof(int) * (size_t)R);
  for (int r = 0; r < R; r++) { /* trp_wrapper.cpp:176-184 */
    counts[r] = oracle_corridor_generation(variant, N, p->delta, s_bounds + (size_t)r * N * 2,
                                           l_bounds + (size_t)r * N * 2, corr + (size_t)r * REGION_CAP, REGION_CAP);
    if (counts[r] < 0) { free(corr); free(counts); res->status = ORACLE_FAIL_TOO_MANY; return res->status; }
  }
  OracleCube tmp[256];
  int K = oracle_collision_check(variant, R, corr, counts, REGION_CAP, N, p->delta, p->s_ref, p->l_ref, tmp, 256);
  free(corr); free(counts);
  if (K == 0) { res->status = ORACLE_FAIL_NO_CORRIDOR; return res->status; }
  if (K < 0 || K > k_max) { res->K = K < 0 ? -K : K; res->status = ORACLE_FAIL_TOO_MANY; return res->status; }
  res->K = K;
  memcpy(segs, tmp, sizeof(OracleCube) * (size_t)K);

  OracleQP qp;
  if (oracle_formulate(variant, p, segs, K, &qp) != 0) { res->status = ORACLE_FAIL_SOLVER; return res->status; }
  OsqpRestateSettings s;
  osqp_restate_default_settings(&s);
  /* SolverDefaultSettings (:1446-1462) then Optimize's overrides (:1236-1243), max_iter 5000 (trp_wrapper.cpp:191) */
  s.eps_prim_inf = 0.000025; s.eps_dual_inf = 0.000025; s.scaled_termination = 1;
  s.max_iter = 5000; s.eps_rel = 1e-5; s.eps_abs = 1e-5; s.scaling = 4; s.polish = 0;
  double *x = (double *)calloc((size_t)qp.n + 1, sizeof(double));
  OsqpRestateInfo info;
  memset(&info, 0, sizeof(info));
  if (mode == 0) {
    osqp_restate_solve(qp.n, qp.m, qp.P_p, qp.P_i, qp.P_x, qp.q, qp.A_p, qp.A_i, qp.A_x, qp.l, qp.u, &s, x, NULL, &info);
  } else {
    /* converged optimum: tighten until the polish is accepted */
    static const double eps_ladder[3] = {1e-6, 1e-8, 1e-10};
    s.polish = 1; s.delta = 1e-9; s.polish_refine_iter = 8; s.max_iter = 50000; s.polish_rounds = 12;
    double *xt = (double *)calloc((size_t)qp.n + 1, sizeof(double));
    int have = 0;
    for (int a = 0; a < 3; a++) {
      OsqpRestateInfo it;
      memset(&it, 0, sizeof(it));
      s.eps_abs = s.eps_rel = eps_ladder[a];
      osqp_restate_solve(qp.n, qp.m, qp.P_p, qp.P_i, qp.P_x, qp.q, qp.A_p, qp.A_i, qp.A_x, qp.l, qp.u, &s, xt, NULL, &it);
      const int solved = it.status == 1 || it.status == 2;
      /* a tighter rung that runs out of iterations does not un-solve the problem: keep the last solved rung */
      if (solved || !have) { info = it; memcpy(x, xt, sizeof(double) * (size_t)qp.n); have = solved; }
      if (it.polish_status == 2 || !solved) break;
    }
    free(xt);
  }
  res->iters = info.iter;
  res->polish_status = info.polish_status;
  int st = info.status;
  if (st < 0 || (st != 1 && st != 2) || info.obj_val != info.obj_val) { /* :1253-1277 */
    res->status = ORACLE_FAIL_SOLVER;
    oracle_qp_free(&qp); free(x);
    return res->status;
  }
  memcpy(ctrl, x, sizeof(double) * (size_t)qp.n);
  res->obj = oracle_qp_objective(&qp, x);
  oracle_qp_free(&qp); free(x);
  int npts = oracle_sample(p, segs, K, ctrl, samples, samples_cap);
  if (npts < 0) { res->status = ORACLE_FAIL_POINTS_CHECK; return res->status; }
  res->npts = npts;
  res->a_cost = oracle_cost(variant, p, samples, npts);
  res->status = (st == 2) ? ORACLE_SOLVED_INACCURATE : ORACLE_OK;
  return res->status;
}

int oracle_solve_batch(int variant, int B, int N, int R, double delta, const double *s_bounds,
                       const double *l_bounds, const double *ds_bounds, const double *dl_bounds,
                       const double *s_ref, const double *l_ref, const double *init,
                       const double *scalars, const double *weights, int weights_stride,
                       int mode, int k_max, int nthreads, int *K, OracleCube *segs, double *ctrl,
                       double *obj, double *a_cost, int *status, int *iters, int *npts,
                       double *samples, int samples_cap, int *polish) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 4)
  for (int b = 0; b < B; b++) {
    OracleProblem p;
    p.n_knots = N; p.delta = delta;
    for (int i = 0; i < 3; i++) { p.init_s[i] = init[6 * b + i]; p.init_l[i] = init[6 * b + 3 + i]; }
    const double *sc = scalars + 10 * (size_t)b;
    p.ds_ref = sc[0]; p.dl_ref = sc[1]; p.dds_lo = sc[2]; p.dds_hi = sc[3]; p.ddds_lo = sc[4];
    p.ddds_hi = sc[5]; p.ddl_lo = sc[6]; p.ddl_hi = sc[7]; p.dddl_lo = sc[8]; p.dddl_hi = sc[9];
    p.ds_bounds = ds_bounds + (size_t)b * N * 2;
    p.dl_bounds = dl_bounds + (size_t)b * N * 2;
    p.s_ref = s_ref + (size_t)b * N;
    p.l_ref = l_ref + (size_t)b * N;
    const double *w = weights + (size_t)(weights_stride ? b : 0) * 10;
    for (int i = 0; i < 10; i++) p.w[i] = w[i];
    OracleResult r;
    double *smp = samples ? samples + (size_t)b * samples_cap * 6 : (double *)malloc(sizeof(double) * 6 * 512);
    int cap = samples ? samples_cap : 512;
    OracleCube *sg = segs + (size_t)b * k_max;
    double *ct = ctrl + (size_t)b * 12 * k_max;
    memset(sg, 0, sizeof(OracleCube) * (size_t)k_max);
    memset(ct, 0, sizeof(double) * 12 * (size_t)k_max);
    oracle_find_traj_mem(variant, &p, R, s_bounds + (size_t)b * R * N * 2, l_bounds + (size_t)b * R * N * 2,
                         mode, k_max, sg, ct, smp, cap, &r);
    K[b] = r.K; obj[b] = r.obj; a_cost[b] = r.a_cost; status[b] = r.status; iters[b] = r.iters;
    if (npts) npts[b] = r.npts;
    if (polish) polish[b] = r.polish_status;
    if (!samples) free(smp);
  }
  return 0;
}
