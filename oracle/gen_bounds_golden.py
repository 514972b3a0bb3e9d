"""oracle/gen_bounds_golden.py -- TEST INFRASTRUCTURE: golden vectors of the obstacle -> region-bounds step, produced by the
reference's OWN Car / get_bounds / lineFromPoints (executed from /root/reference/src/cart_frenet.py by oracle/ref_python.py)
with the caller's call sequence (cart_frenet.py:1539-1557).  Run once in the build container:
    python oracle/gen_bounds_golden.py        -> tests/golden/bounds.npz (committed)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_python  # noqa: E402


def reference_bounds(ns, obstacles):
    Car = ns["Car"]
    Car._lateral = []
    for r, (centre, vel_s, vel_l, horizon) in enumerate(obstacles):
        c = Car(tuple(centre), vel_s=vel_s, vel_l=vel_l, time=horizon, ref=10 + r)
        c.getCar()
    return ns["get_bounds"](Car._lateral)


def cases():
    """Obstacle sets: the reference's own two-car configuration (cart_frenet.py:1536-1546) at several relative positions,
    plus seeded random sets of 1-3 cars (lateral positions on a 0.25 m grid so that edge coincidences occur)."""
    out = []
    for ds2, l2, ds1, l1 in ((12.0, 3.5, 25.0, 0.0), (5.0, 3.5, 30.0, 3.5), (8.0, 0.5, 20.0, 3.0), (15.0, 2.0, 18.0, 2.0),
                             (-6.0, 3.5, 14.0, 0.0), (22.0, 4.0, -9.0, 1.0)):
        c2 = ((abs(ds2), l2, 0 if ds2 > 0 else abs(ds2) / 5.0), 4.5, 0.0, 4.0)
        c1 = ((ds1, l1, 0 if ds1 > 0 else abs(ds1) / 5.0), 6.0, 0.0, 4.0)
        out.append([c2, c1])
    rng = np.random.default_rng(20230606)
    for _ in range(58):
        n = int(rng.integers(1, 4))
        obs = []
        for _k in range(n):
            s = float(np.round(rng.uniform(2.0, 35.0), 1))
            l = float(rng.integers(-4, 25)) * 0.25
            t0 = 0 if rng.random() < 0.6 else float(np.round(rng.uniform(0.5, 3.0), 1))
            vel_s = float(np.round(rng.uniform(1.0, 8.0), 1))
            vel_l = float(rng.choice([0.0, 0.0, 0.25, -0.25, 0.5]))
            horizon = float(rng.choice([3.0, 4.0, 5.0]))
            obs.append(((s, l, t0), vel_s, vel_l, horizon))
        out.append(obs)
    return out


def main():
    ns = ref_python.load(["Car", "delete_multiple_element", "lineFromPoints", "get_bounds"], extra={"print": lambda *a, **k: None})
    sys.path.insert(0, HERE)
    import bounds_oracle as bo
    out, kept = {}, 0
    for obs in cases():
        _p, edges = bo.lateral_edges(obs)
        firsts = [e[0] for e in edges]
        if len(set(firsts)) != len(firsts):
            continue   # tie in the smallest lateral edge: the reference's car order is hash-dependent there (not pinned)
        ref = reference_bounds(ns, obs)
        arr = np.array([[list(c), vs, vl, T] for c, vs, vl, T in obs], dtype=object)
        flat = np.array([[c[0], c[1], c[2], vs, vl, T] for c, vs, vl, T in obs], dtype=np.float64)
        out["case%02d/obstacles" % kept] = flat
        out["case%02d/s_bounds" % kept] = np.array([b[0] for b in ref], dtype=np.float64)          # [R][N][2]
        out["case%02d/l_bounds" % kept] = np.array([b[1] for b in ref], dtype=np.float64)          # [R][N][2]
        kept += 1
    out["n_cases"] = np.array(kept)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bounds.npz"), **out)
    print("wrote", kept, "cases")


if __name__ == "__main__":
    main()
