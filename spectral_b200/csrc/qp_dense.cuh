// spectral_b200/csrc/qp_dense.cuh -- K4: the batched ADMM (OSQP-equivalent) hot loop, dense-operator form.
//
// Replaces the iteration that the reference delegates to OSQP (osqp_solve, called at solve_3d.cc:1249 |
// cuboid_3d.cc:1108 with the settings of solve_3d.cc:1235-1243,1446-1462) for scenarios with at most
// KC <= 16 Bezier segments (instantiated for KC = 8, 10, 12, 16).  One CTA = one scenario = both axis problems (s and l), solved as ONE OSQP
// instance like the reference does: one cost scaling c, one rho, joint termination and infeasibility norms.
//
// Why this form.  The lane-per-segment loop of qp.cuh (block-tridiagonal Cholesky, one lane per segment)
// spends its time in 2K dependent neighbour-to-neighbour hops per iteration and leaves 3/4 of the SM idle
// at the batch sizes of BASELINE.json (ncu, profiles/r1_qp_lanes.md: 13 k cycles per iteration, FP64 pipe
// 5 % busy).  Here the reduced KKT matrix S = P + sigma I + A' rho A of one axis (n = 6 KC <= 60) is inverted
// ONCE per rho (setup and the rare adaptive-rho updates) and the inverse G = S^-1 is kept in REGISTERS, one
// row per thread (n doubles), so the per-iteration solve x~ = G g is n independent FMAs per thread with the
// right-hand side broadcast from shared memory: no dependent chain, no shuffles.  A and A' are never stored:
// their rows are finite-difference stencils of the control points (solve_3d.cc:823-949), applied from
// zero-padded shared-memory arrays so that every thread runs the same instruction stream.
//
// Thread map (TA = 64 threads per axis for KC <= 10, 96 above; CTA = 2 TA):
//   variable thread v < n         : row v of G, relaxed iterate x_v, gathers (A' v)_v, computes x~_v
//   row slots e = ta + TA s < 21KC: w_e = z_e + y_e / rho_e, l_e, u_e, rho_e in registers; rows are ordered by
//                                   type [containment | velocity | acceleration | jerk | continuity/init]
// Per iteration (3 CTA barriers):  S2 gather g = A'(rho (2 clip(w) - w)) + sigma x - q   (13 LDS + 16 FMA)
//                                  S3 x~ = G g, x = alpha x~ + (1 - alpha) x              (n FMA)
//                                  S1 z~ = A x~ (stencil), w += alpha (z~ - clip(w)), next v   (per row slot)
// Every check_termination iterations the residual / infeasibility norms of OSQP are evaluated in the same
// layout (CTA-wide reductions); adaptive-rho updates re-run the factorisation of qp.cuh on warp 0 and
// rebuild G.  Setup (K3 assembly, Ruiz scaling), factorisation, polish and outputs are the lane-per-segment
// code of qp.cuh, executed by warp 0.
#pragma once
#include "common.cuh"
#include "qp.cuh"

#define QPD_VB 33         // doubles per segment block of the padded row-value array V
#define QPD_V0 0          // containment rows      V0[i], i = 0..5
#define QPD_V1 7          // velocity rows         V1[i] at 7 + i,  i = -1..5 (pads at 6, 12)
#define QPD_V2 15         // acceleration rows     V2[i] at 15 + i, i = -2..5 (pads at 13, 14, 19, 20)
#define QPD_V3 24         // jerk rows             V3[i] at 24 + i, i = -3..5 (pads at 21..23, 27..29)
#define QPD_VC 30         // continuity/init rows  Vc[r] at 30 + r
#define QPD_LS 24         // doubles of lane state per segment: t, tp, tn, q[6], sig[6], cD[6]
#define QPD_NRED 10

enum { QPD_T_CONT = 0, QPD_T_VEL = 1, QPD_T_ACC = 2, QPD_T_JERK = 3, QPD_T_JOIN = 4 };

template <int KC>
struct QpdLayout {
  static constexpr int N = 6 * KC;             // variables per axis
  static constexpr int ROWS = 21 * KC;         // constraint rows per axis
  static constexpr int TA = ((N + 31) / 32) * 32;          // threads per axis problem (64 for KC <= 10, 96 above)
  static constexpr int NWARPS = 2 * TA / 32;               // warps per CTA
  static constexpr int NS = (ROWS + TA - 1) / TA;          // row slots per thread
  static constexpr int LPA = KC <= 8 ? 8 : 16; // lanes per axis of the lane-per-segment (control) code
  static constexpr int STR = LPA;              // its shared-memory stride
  // per-axis shared memory (doubles)
  static constexpr int O_CTRL = 0;                                   // W, L, U, RHO, P slots of qp.cuh
  static constexpr int O_FS = O_CTRL + QP_SM_DOUBLES_PER_LANE * STR; // factor store [LPA][57]
  static constexpr int O_LS = O_FS + 57 * LPA;                       // lane state [LPA][QPD_LS]
  static constexpr int O_V = O_LS + QPD_LS * LPA;                    // V[(KC+1)][QPD_VB]
  static constexpr int O_GV = O_V + QPD_VB * (KC + 1) + 1;           // g[N]  (16-byte aligned below)
  static constexpr int O_C = O_GV + N + (N & 1);                     // x~: 3 pad + N (+ pad)
  static constexpr int O_XR = O_C + N + 4;                           // relaxed x for the checks: 3 pad + N
  static constexpr int O_CE = O_XR + N + 4;                          // continuity row coefficients [3KC][6]
  static constexpr int O_VCF = O_CE + 18 * KC;                       // continuity gather coefficients [N][3]
  static constexpr int O_TK = O_VCF + 3 * N;                         // segment durations [KC]
  static constexpr int AXIS = ((O_TK + KC + 1) / 2) * 2;             // doubles per axis (even)
  // per-CTA tail: reduction scratch [NWARPS][QPD_NRED], eqmask ints [2][LPA]
  static constexpr int O_RED = 2 * AXIS;
  static constexpr int O_EQ = O_RED + NWARPS * QPD_NRED;
  // with LPA = 8 the upper half of warp 0 holds no problem: its lanes run the lane-per-segment code on a
  // scratch copy of the control slots so that they never touch the real ones
  static constexpr int O_DUMMY = O_EQ + LPA;  // (2 * LPA ints before it)
  static constexpr int TOTAL = O_DUMMY + (2 * LPA < 32 ? QP_SM_DOUBLES_PER_LANE * STR : 0);
  static constexpr int BYTES = TOTAL * 8;
};

SP_DEV double qpd_clip(double w, double l, double u) { return fmin(fmax(w, l), u); }

// (A' V)_v for variable (k, j): vb = V block of segment k, vk = V block holding the continuity rows that
// touch this variable (own segment for j < 3, next segment for j >= 3), vc = their three coefficients
SP_DEV double qpd_gather(const double *vb, const double *vk, int j, double tk, double vc0, double vc1, double vc2) {
  double g = tk * vb[QPD_V0 + j];
  g += 5.0 * (vb[QPD_V1 + j - 1] - vb[QPD_V1 + j]);
  g += 20.0 * ((vb[QPD_V2 + j - 2] - vb[QPD_V2 + j - 1]) - (vb[QPD_V2 + j - 1] - vb[QPD_V2 + j]));
  const double a0 = vb[QPD_V3 + j - 3], a1 = vb[QPD_V3 + j - 2], a2 = vb[QPD_V3 + j - 1], a3 = vb[QPD_V3 + j];
  g += 60.0 * ((a0 - a3) + 3.0 * (a2 - a1));
  g += vc0 * vk[QPD_VC] + vc1 * vk[QPD_VC + 1] + vc2 * vk[QPD_VC + 2];
  return g;
}

// (A c)_e for a row of the given type; cp points at c[k][i] (continuity rows: at c[k-1][3]), ce = its 6 coefficients
SP_DEV double qpd_row(int type, const double *cp, double tk, const double *ce) {
  if (type == QPD_T_CONT) return tk * cp[0];
  if (type == QPD_T_VEL) return 5.0 * (cp[1] - cp[0]);
  if (type == QPD_T_ACC) return 20.0 * ((cp[2] - cp[1]) - (cp[1] - cp[0]));
  if (type == QPD_T_JERK) return 60.0 * (((cp[3] - cp[2]) - (cp[2] - cp[1])) - ((cp[2] - cp[1]) - (cp[1] - cp[0])));
  return ce[0] * cp[0] + ce[1] * cp[1] + ce[2] * cp[2] + ce[3] * cp[3] + ce[4] * cp[4] + ce[5] * cp[5];
}

template <int KC>
struct QpdRowSlot {
  double w, l, u, rho;
  int coff;   // offset of the first control point of the stencil in C / XR
  int voff;   // offset of this row's value in V
  int ooff;   // offset of this row in the lane-per-segment slots (r * STR + k), add QP_SM_* * STR
  int meta;   // bits 0-2 type, bit 3 valid, bit 4 equality row, bits 8.. segment, bits 16.. row index in type
};

// in-place solve S x = e_v with the block factor of qp.cuh read from shared memory (broadcast loads):
// on return g[] holds row v of G = S^-1
template <int KC>
SP_DEV void qpd_inverse_row(const double *fs, int v, double *g) {
  const int kv = v / 6, iv = v - 6 * kv;
  // forward: y_k = Linv_k e_k - C_k y_{k-1}[3..5]
#pragma unroll
  for (int k = 0; k < KC; k++) {
    const double *F = fs + 57 * k;
#pragma unroll
    for (int a = 0; a < 6; a++) {
      double y = 0.0;
      if (k == kv && a >= iv) y = F[LT(a, 0) + iv];
      if (k > 0) y -= F[21 + a * 3 + 0] * g[6 * (k - 1) + 3] + F[21 + a * 3 + 1] * g[6 * (k - 1) + 4] + F[21 + a * 3 + 2] * g[6 * (k - 1) + 5];
      g[6 * k + a] = y;
    }
  }
  // backward: x_k = Linv_k' y_k - E_k x_{k+1}[0..2]
#pragma unroll
  for (int k = KC - 1; k >= 0; k--) {
    const double *F = fs + 57 * k;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      double x = 0.0;
#pragma unroll
      for (int a = i; a < 6; a++) x += F[LT(a, i)] * g[6 * k + a];
      if (k < KC - 1) x -= F[39 + i * 3 + 0] * g[6 * (k + 1) + 0] + F[39 + i * 3 + 1] * g[6 * (k + 1) + 1] + F[39 + i * 3 + 2] * g[6 * (k + 1) + 2];
      g[6 * k + i] = x;
    }
  }
}

// CTA-wide reduction of QPD_NRED values: slots 0..8 max, slot 9 sum.  All threads return the result in r[].
template <typename SyncFn>
SP_DEV void qpd_reduce(double r[QPD_NRED], double *red, int warp, int lane, int nwarps, SyncFn sync_cta) {
#pragma unroll
  for (int i = 0; i < QPD_NRED - 1; i++) r[i] = sp_group_max(r[i], 32);
  r[QPD_NRED - 1] = sp_group_sum(r[QPD_NRED - 1], 32);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < QPD_NRED; i++) red[warp * QPD_NRED + i] = r[i];
  }
  sync_cta();
#pragma unroll
  for (int i = 0; i < QPD_NRED; i++) r[i] = red[i];
  for (int w = 1; w < nwarps; w++) {
#pragma unroll
    for (int i = 0; i < QPD_NRED - 1; i++) r[i] = fmax(r[i], red[w * QPD_NRED + i]);
    r[QPD_NRED - 1] += red[w * QPD_NRED + QPD_NRED - 1];
  }
  sync_cta();  // red[] may be reused
}

// warp 0: (re)factorise S for the current RHO slots and publish the factor; returns bad pivot flag (joint)
template <int KC>
SP_DEV_NOINLINE int qpd_refactor(double *smc, double *fs, const QpLane &Q, bool publish) {
  constexpr int LPA = QpdLayout<KC>::LPA, STR = QpdLayout<KC>::STR;
  QpFactor F;
  int bad = qp_factorize<LPA, STR>(smc, Q.seg, QP_SM_RHO, Q.sig, Q.t, Q.tp, Q.tn, Q.first, Q.last, Q.active, Q.seg, Q.kmaxw, 0u,
                                   Q.eqmask, 0.0, F);
  bad = sp_group_or(bad, 2 * LPA);
  if (publish) {
    double *dst = fs + 57 * Q.seg;
#pragma unroll
    for (int e = 0; e < 21; e++) dst[e] = F.Linv[e];
#pragma unroll
    for (int e = 0; e < 18; e++) { dst[21 + e] = F.C[e]; dst[39 + e] = F.E[e]; }
  }
  return bad;
}

SP_DEV void qpd_store_lane(const QpLane &Q, double *ls, int *eq) {
  double *d = ls + QPD_LS * Q.seg;
  d[0] = Q.t; d[1] = Q.tp; d[2] = Q.tn;
#pragma unroll
  for (int j = 0; j < 6; j++) { d[3 + j] = Q.q[j]; d[9 + j] = Q.sig[j]; d[15 + j] = Q.cD[j]; }
  eq[Q.seg] = (int)Q.eqmask;
}

// rebuilds the lane state of a control lane from shared memory (everything but c / rhobar, passed in)
SP_DEV void qpd_load_lane(QpLane &Q, const QpArgs &a, int ap, int seg, const double *ls, const int *eq, double c, double rhobar,
                          bool real) {
  Q.b = a.list[ap >> 1]; Q.axis = ap & 1; Q.K = a.K[Q.b]; Q.seg = seg; Q.kmaxw = Q.K;
  Q.have = real; Q.active = real && seg < Q.K; Q.first = seg == 0; Q.last = seg == Q.K - 1;
  const double *d = ls + QPD_LS * seg;
  Q.t = d[0]; Q.tp = d[1]; Q.tn = d[2];
#pragma unroll
  for (int j = 0; j < 6; j++) { Q.q[j] = d[3 + j]; Q.sig[j] = d[9 + j]; Q.cD[j] = d[15 + j]; }
  Q.c = c; Q.rhobar = rhobar; Q.eqmask = (unsigned)eq[seg];
}

// warp 0: K3 assembly, Ruiz scaling, per-row rho, stencil tables, first factorisation.
// Publishes (c, state) in red[0..1].
template <int KC>
SP_DEV_NOINLINE void qpd_control_setup(const QpArgs &a, int slot, int lane, double *smem) {
  using L = QpdLayout<KC>;
  constexpr int LPA = L::LPA, STR = L::STR, JW = 2 * L::LPA;
  double *red = smem + L::O_RED;
  int *eqm = (int *)(smem + L::O_EQ);
  const int cgrp = lane / LPA, cseg = lane % LPA, caxis = cgrp & 1;
  const bool creal = lane < JW;  // lanes beyond the two axis groups idle on scratch slots
  double *smc = creal ? smem + caxis * L::AXIS + L::O_CTRL : smem + L::O_DUMMY;
  double *cfs = smem + caxis * L::AXIS + L::O_FS;
  double *cls = smem + caxis * L::AXIS + L::O_LS;
  int *ceq = eqm + caxis * LPA;
  const int cap = 2 * slot + caxis;
  (void)red; (void)cfs; (void)cls; (void)ceq; (void)cap; (void)smc; (void)cseg; (void)STR;
  int state = QP_RUNNING;
  {
    QpLane Q;
    qp_setup<LPA, STR, JW>(a, cap, creal, cseg, cseg, smc, Q);
    if (creal) qpd_store_lane(Q, cls, ceq);
    if (creal && cseg < KC) {
      // stencil tables of the dense loop: continuity row coefficients, gather coefficients, durations
      double *ce = smem + caxis * L::AXIS + L::O_CE + 18 * cseg;
      double *vcf = smem + caxis * L::AXIS + L::O_VCF + 18 * cseg;
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          ce[6 * r + j] = Q.active ? prev_coef(18 + r, j, Q.t, Q.tp, Q.first) : 0.0;
          ce[6 * r + 3 + j] = Q.active ? row_coef(18 + r, j, Q.t, Q.tp, Q.first) : 0.0;
          vcf[3 * j + r] = Q.active ? row_coef(18 + r, j, Q.t, Q.tp, Q.first) : 0.0;                          // own rows, j < 3
          vcf[3 * (3 + j) + r] = (Q.active && !Q.last) ? prev_coef(18 + r, j, Q.tn, Q.t, false) : 0.0;       // next segment's rows
        }
      smem[caxis * L::AXIS + L::O_TK + cseg] = Q.active ? Q.t : 0.0;
    }
    const int bad = qpd_refactor<KC>(smc, cfs, Q, creal);
    if (bad) state = QP_ST_INFEASIBLE;
    if (lane == 0) { red[0] = Q.c; red[1] = (double)state; }
  }
}

// warp 0: adaptive-rho refactorisation; publishes the bad-pivot flag in red[0]
template <int KC>
SP_DEV_NOINLINE void qpd_control_refactor(const QpArgs &a, int slot, int lane, double *smem, double c_scale, double rhobar) {
  using L = QpdLayout<KC>;
  constexpr int LPA = L::LPA, STR = L::STR, JW = 2 * L::LPA;
  double *red = smem + L::O_RED;
  int *eqm = (int *)(smem + L::O_EQ);
  const int cgrp = lane / LPA, cseg = lane % LPA, caxis = cgrp & 1;
  const bool creal = lane < JW;  // lanes beyond the two axis groups idle on scratch slots
  double *smc = creal ? smem + caxis * L::AXIS + L::O_CTRL : smem + L::O_DUMMY;
  double *cfs = smem + caxis * L::AXIS + L::O_FS;
  double *cls = smem + caxis * L::AXIS + L::O_LS;
  int *ceq = eqm + caxis * LPA;
  const int cap = 2 * slot + caxis;
  (void)red; (void)cfs; (void)cls; (void)ceq; (void)cap; (void)smc; (void)cseg; (void)STR;
  QpLane Q;
  qpd_load_lane(Q, a, cap, cseg, cls, ceq, c_scale, rhobar, creal);
  const int bad = qpd_refactor<KC>(smc, cfs, Q, creal);
  if (lane == 0) red[0] = (double)bad;
}

// warp 0: final status at max_iter, polish, outputs (qp_finish of qp.cuh)
template <int KC>
SP_DEV_NOINLINE void qpd_control_finish(const QpArgs &a, int slot, int lane, double *smem, double c_scale, double rhobar, int state,
                                        int iters) {
  using L = QpdLayout<KC>;
  constexpr int LPA = L::LPA, STR = L::STR, JW = 2 * L::LPA;
  double *red = smem + L::O_RED;
  int *eqm = (int *)(smem + L::O_EQ);
  const int cgrp = lane / LPA, cseg = lane % LPA, caxis = cgrp & 1;
  const bool creal = lane < JW;  // lanes beyond the two axis groups idle on scratch slots
  double *smc = creal ? smem + caxis * L::AXIS + L::O_CTRL : smem + L::O_DUMMY;
  double *cfs = smem + caxis * L::AXIS + L::O_FS;
  double *cls = smem + caxis * L::AXIS + L::O_LS;
  int *ceq = eqm + caxis * LPA;
  const int cap = 2 * slot + caxis;
  (void)red; (void)cfs; (void)cls; (void)ceq; (void)cap; (void)smc; (void)cseg; (void)STR;
  QpLane Q;
  qpd_load_lane(Q, a, cap, cseg, cls, ceq, c_scale, rhobar, creal);
  double x[6];
  const double *xs = smem + caxis * L::AXIS + L::O_XR + 3 + 6 * (cseg < KC ? cseg : 0);
#pragma unroll
  for (int j = 0; j < 6; j++) x[j] = (creal && cseg < KC) ? xs[j] : 0.0;
  qp_finish<LPA, STR, JW>(a, smc, cseg, Q, x, creal ? state : QP_ST_MAXITER, iters);
}

// ------------------------------------------------------------------ the CTA body
// slot: index of the scenario in this class' list.  tid in [0, 2 TA).  smem: QpdLayout<KC>::BYTES, 16-byte aligned.
template <int KC, typename SyncFn>
SP_DEV void qpd_cta_body(const QpArgs &a, int slot, int tid, double *smem, SyncFn sync_cta) {
  using L = QpdLayout<KC>;
  constexpr int N = L::N, ROWS = L::ROWS, NS = L::NS, LPA = L::LPA, STR = L::STR, JW = 2 * L::LPA, QPD_TA = L::TA;
  (void)LPA; (void)JW;
  if (slot >= *a.count) return;
  const SpOptionsDev &o = a.opt;
  const int warp = tid >> 5, lane = tid & 31;
  const int axis = tid / QPD_TA, ta = tid - axis * QPD_TA;
  double *smx = smem + axis * L::AXIS;  // this thread's axis region (fast layout)
  double *red = smem + L::O_RED;
  int *eqm = (int *)(smem + L::O_EQ);
  const int b = a.list[slot];
  const int K = a.K[b];

  // (warp 0 doubles as the control warp: adjacent lane groups hold the s-axis and the l-axis problem in the
  // lane-per-segment layout of qp.cuh, see qpd_control_*)
  // ---------------- setup on warp 0: K3 assembly, Ruiz scaling, rho, first factorisation ----------------
  double c_scale = 1.0, rhobar = o.rho0;
  int state = QP_RUNNING;
  if (warp == 0) qpd_control_setup<KC>(a, slot, lane, smem);
  // zero the padded arrays of the fast layout (V incl. pads and the extra block, C / XR pads)
  for (int i = ta; i < QPD_VB * (KC + 1) + 1; i += QPD_TA) smx[L::O_V + i] = 0.0;
  for (int i = ta; i < N + 4; i += QPD_TA) { smx[L::O_C + i] = 0.0; smx[L::O_XR + i] = 0.0; }
  sync_cta();
  c_scale = red[0];
  state = (int)red[1];
  sync_cta();

  // ---------------- fast-layout state ----------------
  QpdRowSlot<KC> rs[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    const int e = ta + QPD_TA * s;
    int type, k, i;
    if (e < 6 * KC) { type = QPD_T_CONT; k = e / 6; i = e - 6 * k; }
    else if (e < 11 * KC) { type = QPD_T_VEL; k = (e - 6 * KC) / 5; i = (e - 6 * KC) - 5 * k; }
    else if (e < 15 * KC) { type = QPD_T_ACC; k = (e - 11 * KC) / 4; i = (e - 11 * KC) - 4 * k; }
    else if (e < 18 * KC) { type = QPD_T_JERK; k = (e - 15 * KC) / 3; i = (e - 15 * KC) - 3 * k; }
    else { type = QPD_T_JOIN; k = (e - 18 * KC) / 3; i = (e - 18 * KC) - 3 * k; }
    const bool valid = e < ROWS;
    if (!valid) { type = QPD_T_CONT; k = 0; i = 0; }
    const int r_old = (type == QPD_T_CONT ? 0 : type == QPD_T_VEL ? 6 : type == QPD_T_ACC ? 11 : type == QPD_T_JERK ? 15 : 18) + i;
    const int vbase = type == QPD_T_CONT ? QPD_V0 : type == QPD_T_VEL ? QPD_V1 : type == QPD_T_ACC ? QPD_V2 : type == QPD_T_JERK ? QPD_V3 : QPD_VC;
    rs[s].coff = type == QPD_T_JOIN ? 6 * k : 3 + 6 * k + i;
    rs[s].voff = QPD_VB * k + vbase + i;
    rs[s].ooff = r_old * STR + k;
    const bool live = valid && k < K;  // rows of unused segments stay inert: rho = 0, v = 0
    const int eq = live ? ((eqm[axis * LPA + k] >> r_old) & 1) : 0;
    rs[s].meta = type | (valid ? 8 : 0) | (eq ? 16 : 0) | (k << 8) | (i << 16);
    const double *ctl = smx + L::O_CTRL;
    rs[s].w = 0.0;
    rs[s].l = live ? ctl[QP_SM_L * STR + rs[s].ooff] : -1.0;
    rs[s].u = live ? ctl[QP_SM_U * STR + rs[s].ooff] : 1.0;
    rs[s].rho = live ? ctl[QP_SM_RHO * STR + rs[s].ooff] : 0.0;
  }
  // variable thread
  const bool isvar = ta < N;
  const int vk = isvar ? ta / 6 : 0, vj = isvar ? ta - 6 * vk : 0;
  const double *vb = smx + L::O_V + QPD_VB * vk;
  const double *vkk = smx + L::O_V + QPD_VB * (vj < 3 ? vk : vk + 1);
  double xv = 0.0, sigv = 0.0, qv = 0.0, cDv = 0.0, tkv = 0.0, vc0 = 0.0, vc1 = 0.0, vc2 = 0.0;
  if (isvar) {
    const double *d = smx + L::O_LS + QPD_LS * vk;
    qv = d[3 + vj]; sigv = d[9 + vj]; cDv = d[15 + vj];
    tkv = smx[L::O_TK + vk];
    const double *vcf = smx + L::O_VCF + 3 * ta;
    vc0 = vcf[0]; vc1 = vcf[1]; vc2 = vcf[2];
    if (vk >= K) { qv = 0.0; sigv = 0.0; cDv = 0.0; }
  }
  double G[N];
  if (isvar) qpd_inverse_row<KC>(smx + L::O_FS, ta, G);
  else {
#pragma unroll
    for (int e = 0; e < N; e++) G[e] = 0.0;
  }
  sync_cta();

  // ---------------- ADMM ----------------
  const double alpha = o.alpha;
  int iters = 0;
  double *gv = smx + L::O_GV;
  double *cx = smx + L::O_C;
  double *xr = smx + L::O_XR;
  double *vv = smx + L::O_V;
  const double *cetab = smx + L::O_CE;
  const double *tktab = smx + L::O_TK;
  for (int it = 1; it <= o.max_iter && state == QP_RUNNING; it++) {
    const bool first_it = it == 1;
    const bool check = (o.check_every > 0) && (it % o.check_every == 0);
    // S2: right-hand side g = A' v + sigma x - q  (V holds v = rho (2 clip(w) - w); zeros on the cold start)
    if (isvar) gv[ta] = qpd_gather(vb, vkk, vj, tkv, vc0, vc1, vc2) + sigv * xv - qv;
    sync_cta();
    // S3: x~ = G g, relaxation
    if (isvar) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int e = 0; e < N; e += 4) {
        a0 += G[e] * gv[e]; a1 += G[e + 1] * gv[e + 1]; a2 += G[e + 2] * gv[e + 2]; a3 += G[e + 3] * gv[e + 3];
      }
      const double xt = (a0 + a1) + (a2 + a3);
      xv = alpha * xt + (1.0 - alpha) * xv;
      cx[3 + ta] = xt;
    }
    sync_cta();
    // S1: z~ = A x~, w update, next v (check iterations: delta y instead of v)
    double red_v[QPD_NRED];
#pragma unroll
    for (int i = 0; i < QPD_NRED; i++) red_v[i] = 0.0;
    const double c_over_rhobar = c_scale / rhobar;
#pragma unroll
    for (int s = 0; s < NS; s++) {
      const int meta = rs[s].meta;
      if (!(meta & 8)) continue;
      const int type = meta & 7, k = (meta >> 8) & 0xff, i = meta >> 16;
      const double zt = qpd_row(type, cx + rs[s].coff, tktab[k], cetab + 18 * k + 6 * i);
      const double w = rs[s].w, l = rs[s].l, u = rs[s].u, rho = rs[s].rho;
      const double p = first_it ? 0.0 : qpd_clip(w, l, u);
      const double wn = first_it ? alpha * zt : w + alpha * (zt - p);
      const double pn = qpd_clip(wn, l, u);
      rs[s].w = wn;
      if (!check) {
        vv[rs[s].voff] = rho * (2.0 * pn - wn);
      } else {
        const double yo = first_it ? 0.0 : rho * (w - p);
        const double dy = rho * (wn - pn) - yo;
        vv[rs[s].voff] = dy;
        if (rho > 0.0) {
          const double Er = sqrt(rho * c_over_rhobar * ((meta & 16) ? 1e-3 : 1.0));
          red_v[7] = fmax(red_v[7], fabs(c_scale * dy / Er));
          red_v[9] += c_scale * (u * fmax(dy, 0.0) + l * fmin(dy, 0.0));
        }
      }
    }
    iters = it;
    sync_cta();
    if (!check) continue;

    // ---------------- termination / infeasibility check (OSQP, scaled space, joint over both axes) ----------------
    if (isvar) {
      const double atd = qpd_gather(vb, vkk, vj, tkv, vc0, vc1, vc2);
      red_v[8] = fabs(cDv * atd);
      xr[3 + ta] = xv;
    }
    sync_cta();
#pragma unroll
    for (int s = 0; s < NS; s++) {
      if (!(rs[s].meta & 8)) continue;
      const double w = rs[s].w;
      vv[rs[s].voff] = rs[s].rho * (w - qpd_clip(w, rs[s].l, rs[s].u));  // y
    }
    sync_cta();
    if (isvar) {
      const double aty = qpd_gather(vb, vkk, vj, tkv, vc0, vc1, vc2);
      double px = 0.0;
      const double *pk = smx + L::O_CTRL + QP_SM_P * STR + vk;
#pragma unroll
      for (int i = 0; i < 6; i++) {
        const int e = (i >= vj) ? LT(i, vj) : LT(vj, i);
        px += pk[e * STR] * xr[3 + 6 * vk + i];
      }
      if (vk >= K) px = 0.0;
      red_v[1] = cDv * fabs(px + qv + aty);
      red_v[4] = cDv * fabs(qv);
      red_v[5] = cDv * fabs(px);
      red_v[6] = cDv * fabs(aty);
    }
#pragma unroll
    for (int s = 0; s < NS; s++) {
      const int meta = rs[s].meta;
      if (!(meta & 8) || !(rs[s].rho > 0.0)) continue;
      const int type = meta & 7, k = (meta >> 8) & 0xff, i = meta >> 16;
      const double ax = qpd_row(type, xr + rs[s].coff, tktab[k], cetab + 18 * k + 6 * i);
      const double z = qpd_clip(rs[s].w, rs[s].l, rs[s].u);
      const double Er = sqrt(rs[s].rho * c_over_rhobar * ((meta & 16) ? 1e-3 : 1.0));
      red_v[0] = fmax(red_v[0], Er * fabs(ax - z));
      red_v[2] = fmax(red_v[2], Er * fabs(z));
      red_v[3] = fmax(red_v[3], Er * fabs(ax));
    }
    qpd_reduce(red_v, red, warp, lane, L::NWARPS, sync_cta);
    const double pri = red_v[0], dua = red_v[1], nz = red_v[2], nax = red_v[3], nq = red_v[4], npx = red_v[5], naty = red_v[6];
    const double nd = red_v[7], na = red_v[8], lhs = red_v[9];
    const double eps_p = o.eps_abs + o.eps_rel * fmax(nz, nax);
    const double eps_d = o.eps_abs + o.eps_rel * fmax(nq, fmax(npx, naty));
    if (pri < eps_p && dua < eps_d) state = QP_ST_SOLVED;
    else if (!(pri < eps_p) && nd > o.eps_pinf && lhs < -o.eps_pinf * nd && na < o.eps_pinf * nd) state = QP_ST_INFEASIBLE;
    // adaptive rho (OSQP's rule, every adaptive_rho_interval iterations)
    if (state == QP_RUNNING && o.adapt_every > 0 && (it % o.adapt_every == 0)) {
      const double pr = pri / (fmax(nz, nax) + 1e-10);
      const double dr = dua / (fmax(nq, fmax(npx, naty)) + 1e-10);
      double est = rhobar * sqrt(pr / (dr + 1e-10));
      est = fmin(fmax(est, 1e-6), 1e6);
      if (est > rhobar * o.adapt_tol || est < rhobar / o.adapt_tol) {
        const double ratio = est / rhobar;
        double *ctl = smx + L::O_CTRL;
#pragma unroll
        for (int s = 0; s < NS; s++) {
          if (!(rs[s].meta & 8)) continue;
          const double w = rs[s].w, z = qpd_clip(w, rs[s].l, rs[s].u);
          rs[s].w = z + (w - z) / ratio;  // keep (z, y): w' = z + y / rho'
          rs[s].rho *= ratio;
          if (((rs[s].meta >> 8) & 0xff) < K) ctl[QP_SM_RHO * STR + rs[s].ooff] = rs[s].rho;
        }
        rhobar = est;
        sync_cta();
        if (warp == 0) qpd_control_refactor<KC>(a, slot, lane, smem, c_scale, rhobar);
        sync_cta();
        if (red[0] != 0.0) state = QP_ST_INFEASIBLE;
        if (isvar) qpd_inverse_row<KC>(smx + L::O_FS, ta, G);
        sync_cta();
      }
    }
    // V must hold v = rho (2 clip(w) - w) again for the next iteration
#pragma unroll
    for (int s = 0; s < NS; s++) {
      if (!(rs[s].meta & 8)) continue;
      const double w = rs[s].w;
      vv[rs[s].voff] = rs[s].rho * (2.0 * qpd_clip(w, rs[s].l, rs[s].u) - w);
    }
    sync_cta();
  }

  // ---------------- hand the iterate back to the lane-per-segment layout: W slots, x ----------------
  {
    double *ctl = smx + L::O_CTRL;
#pragma unroll
    for (int s = 0; s < NS; s++) {
      if (!(rs[s].meta & 8) || ((rs[s].meta >> 8) & 0xff) >= K) continue;
      ctl[QP_SM_W * STR + rs[s].ooff] = rs[s].w;
      ctl[QP_SM_RHO * STR + rs[s].ooff] = rs[s].rho;
    }
    if (isvar) xr[3 + ta] = xv;
  }
  sync_cta();
  if (warp != 0) return;
  qpd_control_finish<KC>(a, slot, lane, smem, c_scale, rhobar, state, iters);
}
