// spectral_b200/csrc/tables.cuh -- K3a: the weight-dependent tables MQM_d = M' pQp_d M (d = 0..3) of
// CalculateKernel (solve_3d.cc:79-143 | cuboid_3d.cc:76-139).  They depend only on the 10 Params
// weights, so they are built once per weight vector (one thread per (weight set, axis)) and the QP
// kernel scales them by t^3, t, 1/t, 1/t^3 per segment (:159-160).
#pragma once
#include "common.cuh"

// out: 4 x 21 doubles, packed lower triangle (the matrices are symmetric)
SP_DEV void mqm_tables(double w_ref, double w_dref, double w_dd, double w_ddd, double *out) {
  const double M[6][6] = {{1, 0, 0, 0, 0, 0}, {-5, 5, 0, 0, 0, 0}, {10, -20, 10, 0, 0, 0},
                          {-10, 30, -30, 10, 0, 0}, {5, -20, 30, -20, 5, 0}, {-1, 5, -10, 10, -5, 1}};
  for (int d = 0; d < 4; d++) {
    double Q[6][6], T[6][6];
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) {
        double v = 0.0;
        if (d == 0) v = w_ref / (double)(i + j + 1);
        else if (d == 1) { if (i >= 1 && j >= 1) v = (w_dref * i * j) / (double)(i + j - 1); }
        else if (d == 2) { if (i >= 2 && j >= 2) v = (w_dd * i * j * (i - 1) * (j - 1)) / (double)(i + j - 3); }
        else { if (i >= 3 && j >= 3) v = (w_ddd * i * j * (i - 1) * (j - 1) * (i - 2) * (j - 2)) / (double)(i + j - 5); }
        Q[i][j] = v;
      }
    for (int i = 0; i < 6; i++)  // T = M' Q
      for (int j = 0; j < 6; j++) {
        double s = 0.0;
        for (int k = 0; k < 6; k++) s += M[k][i] * Q[k][j];
        T[i][j] = s;
      }
    for (int i = 0; i < 6; i++)  // (M' Q) M, lower triangle
      for (int j = 0; j <= i; j++) {
        double s = 0.0;
        for (int k = 0; k < 6; k++) s += T[i][k] * M[k][j];
        out[d * 21 + i * (i + 1) / 2 + j] = s;
      }
  }
}

// weights in Params order: s_acc, s_jerk, l_acc, l_jerk, s_ref, ds_ref, l_ref, dl_ref, end_s, end_l
SP_DEV void mqm_body(const double *weights, double *mqm, int wset, int axis) {
  const double *w = weights + 10 * (size_t)wset;
  double *out = mqm + ((size_t)wset * 2 + axis) * 84;
  if (axis == 0) mqm_tables(w[4], w[5], w[0], w[1], out);
  else mqm_tables(w[6], w[7], w[2], w[3], out);
}
