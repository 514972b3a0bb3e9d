"""oracle/gen_downstream_golden.py -- TEST INFRASTRUCTURE: golden vectors of the step after the hot path, produced by the
reference's OWN run_ego() and frenet_to_cartesian3D() (executed from /root/reference/src/cart_frenet.py by
oracle/ref_python.py).  Run once in the build container:  python oracle/gen_downstream_golden.py
Writes tests/golden/downstream.npz (committed)."""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_python  # noqa: E402

LITERAL_DIR = "/home/srujan_d/RISS/code/btrapz/src"   # run_ego() opens <dir>/s1_cub_3d_3.txt (cart_frenet.py:1127)


def main():
    ns = ref_python.load(["NormalizeAngle", "frenet_to_cartesian3D", "run_ego"], extra={"print": lambda *a, **k: None})
    out = {}
    os.makedirs(LITERAL_DIR, exist_ok=True)
    # ---- run_ego on the two reproducible shipped trajectories (and on a perturbed copy with a lane change)
    cases = {"cub31": os.path.join(ROOT, "tests", "golden", "s1_cub_3d_31.txt"), "slt31": os.path.join(ROOT, "tests", "golden", "s1_slt_3d_31.txt")}
    for name, path in cases.items():
        shutil.copy(path, os.path.join(LITERAL_DIR, "s1_cub_3d_3.txt"))
        rows = np.loadtxt(path)
        for succ, offset in ((True, 12.5), (True, 0.0)):
            ego = ref_python._Obj(prediction=None)
            curr_state = ref_python._Obj(position=np.array([offset, 0.0]))
            ns["run_ego"](ego, 0, 7, succ, curr_state)
            st = ego.prediction.trajectory.state_list
            arr = np.array([[s.position[0], s.position[1], s.velocity, s.orientation] for s in st])
            key = "%s/off%g" % (name, offset)
            out[key + "/rows"] = rows
            out[key + "/states"] = arr
            out[key + "/time_steps"] = np.array([s.time_step for s in st])
    # ---- frenet_to_cartesian3D on random conditions
    rng = np.random.default_rng(20230605)
    n = 512
    ref = np.stack([rng.uniform(0, 100, n), rng.uniform(-50, 50, n), rng.uniform(-50, 50, n), rng.uniform(-4, 4, n),
                    rng.uniform(-0.05, 0.05, n), rng.uniform(-0.01, 0.01, n)], axis=1)
    s_cond = np.stack([ref[:, 0], rng.uniform(0, 15, n), rng.uniform(-3, 3, n)], axis=1)
    d_cond = np.stack([rng.uniform(-3, 3, n), rng.uniform(-0.5, 0.5, n), rng.uniform(-0.2, 0.2, n)], axis=1)
    res = np.array([ns["frenet_to_cartesian3D"](*ref[i], s_cond[i], d_cond[i]) for i in range(n)])
    out["f2c/ref"], out["f2c/s_cond"], out["f2c/d_cond"], out["f2c/out"] = ref, s_cond, d_cond, res
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "downstream.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
