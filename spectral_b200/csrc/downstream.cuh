// spectral_b200/csrc/downstream.cuh -- the step right AFTER the planning hot path, batched on the device:
//   k_ego_states           /root/reference/src/cart_frenet.py:1126-1221  run_ego(): sampled trajectory -> ego state list
//                          (position along the road, lateral position, speed, heading rounded to 0.01 rad)
//   k_frenet_to_cartesian  /root/reference/src/cart_frenet.py:347-381    frenet_to_cartesian3D() + NormalizeAngle (:193-203)
// so that a sweep can score / hand over Cartesian trajectories without a round trip through the host.  Element-wise, HBM-bound:
// 48 B in + 32 B out per sample (ego states), 96 B in + 48 B out per point (Frenet -> Cartesian).
#pragma once
#include "common.cuh"

// Python's round(x, 2) (cart_frenet.py:1152): the decimal value nearest to the EXACT binary x, ties to even; returned as the
// nearest double.  y = rint(100 x) is corrected with the exact remainder fma(x, 100, -y).
SP_DEV double sp_py_round2(double x) {
  if (!(fabs(x) < 1e13)) return x;
  double y = rint(x * 100.0);
  const double d = fma(x, 100.0, -y);  // exact 100 x - y, rounded once
  if (d > 0.5) y += 1.0;
  else if (d < -0.5) y -= 1.0;
  else if (d == 0.5 && fmod(y, 2.0) != 0.0) y += 1.0;   // exact tie resolved the wrong way by the rounded product
  else if (d == -0.5 && fmod(y, 2.0) != 0.0) y -= 1.0;
  return y / 100.0;
}

// one thread per (scenario, sample).  samples: [B][cap][6] = (s, ds, dds, l, dl, ddl); states: [B][cap][4]
SP_DEV void ego_state_body(const double *samples, const int *npts, int cap, const double *s_offset, int off_stride, double *states, int b,
                           int i) {
  const int n = npts[b] < cap ? npts[b] : cap;
  if (i >= n) return;
  const double *smp = samples + ((size_t)b * cap) * 6;
  const double s = smp[6 * i], ds = smp[6 * i + 1], l = smp[6 * i + 3], dl = smp[6 * i + 4];
  const double dy = ds > 5.0 ? ds : 5.0;                         // :1142-1145
  double a = 0.0;
  if (i + 1 < n) a = atan2(smp[6 * (i + 1) + 3] - l, smp[6 * (i + 1)] - s);       // :1151-1154
  else if (n > 1) a = atan2(l - smp[6 * (i - 1) + 3], s - smp[6 * (i - 1)]);      // :1156-1159 (the except branch: last sample)
  a = sp_py_round2(a);
  if (a != a) a = 0.0;                                           // :1160-1161
  double *o = states + ((size_t)b * cap + i) * 4;
  o[0] = s + (i > 0 ? s_offset[(size_t)b * off_stride] : 0.0);   // :1180 (state 0), :1190-1191 (states i >= 1)
  o[1] = l;
  o[2] = sqrt(dl * dl + dy * dy);                                // (ego_dx^2 + ego_dy^2) ** 0.5
  o[3] = a;
}

// fmod(angle + pi, 2 pi) folded back to [-pi, pi)  (:193-203)
SP_DEV double sp_normalize_angle(double angle) {
  const double pi = 3.141592653589793;
  double a = fmod(angle + pi, 2.0 * pi);
  if (a < 0.0) a += 2.0 * pi;
  return a - pi;
}

// one thread per point.  ref: [n][6] = (rs, rx, ry, rtheta, rkappa, rdkappa), s_cond / d_cond: [n][3]; out: [n][6] = (x, y, v, a, theta, kappa)
SP_DEV void frenet_to_cartesian_body(const double *ref, const double *s_cond, const double *d_cond, double *out, size_t i) {
  const double *r = ref + 6 * i, *sc = s_cond + 3 * i, *dc = d_cond + 3 * i;
  const double rx = r[1], ry = r[2], rtheta = r[3], rkappa = r[4], rdkappa = r[5];
  const double cos_r = cos(rtheta), sin_r = sin(rtheta);
  const double x = rx - sin_r * dc[0];
  const double y = ry + cos_r * dc[0];
  const double om = 1.0 - rkappa * dc[0];
  const double tan_dt = dc[1] / om;
  const double dt = atan2(dc[1], om);
  const double cos_dt = cos(dt);
  const double theta = sp_normalize_angle(dt + rtheta);
  const double kp = rdkappa * dc[0] + rkappa * dc[1];
  const double kappa = ((((dc[2] + kp * tan_dt) * cos_dt * cos_dt) / om + rkappa) * cos_dt / om);
  const double d_dot = dc[1] * sc[1];
  const double v = sqrt(om * om * sc[1] * sc[1] + d_dot * d_dot);
  const double dtp = om / cos_dt * kappa - rkappa;
  const double a = (sc[2] * om / cos_dt + sc[1] * sc[1] / cos_dt * (dc[1] * dtp - kp));
  double *o = out + 6 * i;
  o[0] = x; o[1] = y; o[2] = v; o[3] = a; o[4] = theta; o[5] = kappa;
}
