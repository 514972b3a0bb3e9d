"""oracle/downstream_oracle.py -- TEST INFRASTRUCTURE: CPU restatement of the step right AFTER the planning hot path
(SURVEY.md 8f row 3): the sampled Frenet trajectory -> ego state list, and the Frenet -> Cartesian map.

    ego_states()              /root/reference/src/cart_frenet.py:1126-1221  run_ego(): rows `t s l ds dl dds ddl` of the
                              trajectory file -> per sample (position along the road, lateral position, speed, heading)
    frenet_to_cartesian3d()   /root/reference/src/cart_frenet.py:347-381    frenet_to_cartesian3D() + NormalizeAngle :193-203
Pinned against the reference's own functions (executed from its source by oracle/gen_downstream_golden.py):
tests/golden/downstream.npz, tests/test_oracle_golden.py.  Only tests/ use this module.
"""
import math

import numpy as np


def py_round2(x):
    """Python's round(x, 2) (cart_frenet.py:1152: correctly rounded decimal, ties to even on the exact binary value)."""
    return round(float(x), 2)


def ego_states(samples, s_offset=0.0):
    """samples: [n, 6] rows (s, ds, dds, l, dl, ddl) -- the C-ABI sample layout; returns [n, 4] (pos_s, pos_l, speed, heading).
    run_ego(): ego_y = s (:1140), ego_x = l (:1141), ego_dy = ds floored at 5.0 (:1142-1145), ego_dx = dl (:1146);
    heading[i] = round(arctan2(l[i+1] - l[i], s[i+1] - s[i]), 2), backward difference on the last sample, NaN -> 0
    (:1150-1161); state 0 at (s[0], l[0]) (:1180-1183), states i >= 1 at (s[i] + s_offset, l[i]) (:1190-1194), speed
    (dx^2 + dy^2)^0.5."""
    smp = np.asarray(samples, dtype=np.float64)
    n = smp.shape[0]
    s, ds, l, dl = smp[:, 0], smp[:, 1], smp[:, 3], smp[:, 4]
    dy = np.where(ds > 5.0, ds, 5.0)
    out = np.zeros((n, 4))
    for i in range(n):
        if i + 1 < n:
            a = math.atan2(l[i + 1] - l[i], s[i + 1] - s[i])
        elif n > 1:
            a = math.atan2(l[i] - l[i - 1], s[i] - s[i - 1])
        else:
            a = 0.0
        a = py_round2(a)
        if math.isnan(a):
            a = 0.0
        out[i] = (s[i] + (s_offset if i > 0 else 0.0), l[i], (dl[i] ** 2 + dy[i] ** 2) ** 0.5, a)
    return out


def normalize_angle(angle):
    """cart_frenet.py:193-203."""
    a = math.fmod(angle + math.pi, 2.0 * math.pi)
    if a < 0.0:
        a += 2.0 * math.pi
    return a - math.pi


def frenet_to_cartesian3d(ref, s_cond, d_cond):
    """ref = (rs, rx, ry, rtheta, rkappa, rdkappa); returns (x, y, v, a, theta, kappa) -- cart_frenet.py:347-381, term by term."""
    rs, rx, ry, rtheta, rkappa, rdkappa = [float(v) for v in ref]
    s0, s1, s2 = [float(v) for v in s_cond]
    d0, d1, d2 = [float(v) for v in d_cond]
    cos_r, sin_r = math.cos(rtheta), math.sin(rtheta)
    x = rx - sin_r * d0
    y = ry + cos_r * d0
    om = 1 - rkappa * d0
    tan_dt = d1 / om
    dt = math.atan2(d1, om)
    cos_dt = math.cos(dt)
    theta = normalize_angle(dt + rtheta)
    kp = rdkappa * d0 + rkappa * d1
    kappa = ((((d2 + kp * tan_dt) * cos_dt * cos_dt) / om + rkappa) * cos_dt / om)
    d_dot = d1 * s1
    v = math.sqrt(om * om * s1 * s1 + d_dot * d_dot)
    dtp = om / cos_dt * kappa - rkappa
    a = (s2 * om / cos_dt + s1 * s1 / cos_dt * (d1 * dtp - kp))
    return x, y, v, a, theta, kappa
