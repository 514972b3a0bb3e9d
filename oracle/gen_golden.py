#!/usr/bin/env python
"""oracle/gen_golden.py -- TEST INFRASTRUCTURE: generates tests/golden/ (run in the build container).

Needs /root/reference (fixtures + shipped libtrp.so/libcub.so) and `make -C oracle`.  Produces
  spectral_b200/data/fixtures.npz  the 14 input fixtures (src/c*.txt, bounds.txt) parsed into arrays
  tests/golden/shipped_qp.npz      QP data (P,q,A,l,u) captured from the SHIPPED binaries for every
                                   fixture x variant with weights.txt (full precision) -- pins a2-a8
  tests/golden/ref_segments.npz    new_corridor (all Cube fields) from the reference's own sources
                                   recompiled (oracle/_ref) -- pins a2-a4
  tests/golden/converged.npz       converged QP optimum per feasible fixture x variant (oracle mode 1),
                                   with the HiGHS cross-check distance recorded
  tests/golden/shipped_sampling.npz  for c1/c2/c4_2 x variant: the trajectory file and return value the
                                   SHIPPED binary produces when handed the converged control points
                                   (pins a9 sampling, a10 cost and the %.3f writer)
  tests/golden/s1_slt_3d_31.txt, s1_cub_3d_31.txt   the reference's shipped known-answer outputs
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import pyoracle as po  # noqa: E402
from run_shipped import GOLDEN_W_CUB, GOLDEN_W_TRP, parse_qp, run_shipped  # noqa: E402
from spectral_b200.wire import ScenarioBatch, read_scenario_text  # noqa: E402

REF_SRC = "/root/reference/src"
OUT = os.path.join(ROOT, "tests", "golden")
FIXTURES = ["c1", "c2", "c3", "c4", "c4_2", "c5", "c6", "c7", "c7_7", "c7_10", "c_road_s1", "c_road_s1_2",
            "c_road_s1_3", "bounds"]
W_FILE = (35.73, 41.61, 25.57, 41.59, 0.12, 10.04, 0.71, 14.3, 7.27, 32.13)  # src/weights.txt


def highs_solve(qp):
    import scipy.sparse as sp
    from scipy.optimize._highspy import _core as hp
    n, m = qp["n"], qp["m"]
    Pu = sp.csc_matrix((qp["P_x"], qp["P_i"], qp["P_p"]), shape=(n, n))
    Pl = sp.csc_matrix(Pu.T)
    h = hp._Highs()
    h.setOptionValue("output_flag", False)
    model = hp.HighsModel()
    lp = model.lp_
    lp.num_col_, lp.num_row_ = n, m
    lp.col_cost_ = qp["q"]
    lp.col_lower_ = np.full(n, -hp.kHighsInf)
    lp.col_upper_ = np.full(n, hp.kHighsInf)
    lp.row_lower_, lp.row_upper_ = qp["l"], qp["u"]
    lp.a_matrix_.format_ = hp.MatrixFormat.kColwise
    lp.a_matrix_.start_, lp.a_matrix_.index_, lp.a_matrix_.value_ = qp["A_p"], qp["A_i"], qp["A_x"]
    model.hessian_.dim_ = n
    model.hessian_.format_ = hp.HessianFormat.kTriangular
    model.hessian_.start_, model.hessian_.index_, model.hessian_.value_ = Pl.indptr, Pl.indices, Pl.data
    h.passModel(model)
    h.run()
    ok = "kOptimal" in str(h.getModelStatus())
    return np.array(h.getSolution().col_value), ok


def main():
    os.makedirs(OUT, exist_ok=True)
    fx, qps, segs, conv, samp = {}, {}, {}, {}, {}
    for name in FIXTURES:
        path = os.path.join(REF_SRC, name + ".txt")
        sc = read_scenario_text(path)
        b = ScenarioBatch.from_scenarios([sc])
        fx[name + "/n_knots"] = np.int32(sc.n_knots)
        fx[name + "/delta_t"] = np.float64(sc.delta_t)
        for key, arr in zip(("s_bounds", "l_bounds", "ds_bounds", "dl_bounds", "s_ref", "l_ref", "init", "scalars"),
                            b.arrays()):
            fx[name + "/" + key] = arr[0]
        for variant in ("trp", "cub"):
            tag = "%s/%s" % (name, variant)
            r = run_shipped(variant, path, W_FILE)
            qp = parse_qp(r["qp_text"])
            for k in ("q", "l", "u", "P_p", "P_i", "P_x", "A_p", "A_i", "A_x"):
                qps["%s/%s" % (tag, k)] = qp[k]
            ref = po.solve_batch(variant, b, W_FILE, mode=0, k_max=32, kind="reference")
            K = int(ref["K"][0])
            segs[tag + "/segs"] = ref["segs"][0][:K]
            segs[tag + "/status"] = ref["status"][0]
            segs[tag + "/iters"] = ref["iters"][0]
            # converged optimum (mode 1) + HiGHS cross-check
            c = po.solve_batch(variant, b, W_FILE, mode=1, k_max=32, kind="port")
            xh, ok = highs_solve(qp)
            if c["status"][0] <= 1 and ok:
                x = c["ctrl"][0][:12 * K]
                conv[tag + "/ctrl"] = x
                conv[tag + "/obj"] = c["obj"][0]
                conv[tag + "/highs_maxdiff"] = np.abs(x - xh).max()
                conv[tag + "/polish_ok"] = np.int32(1)
                print("%-18s K=%2d converged obj=%.9f  |x-x_highs|max=%.2e  iters=%d" %
                      (tag, K, c["obj"][0], np.abs(x - xh).max(), c["iters"][0]))
            else:
                print("%-18s K=%2d oracle status=%d highs_optimal=%s" % (tag, K, c["status"][0], ok))
    # sampling / cost / writer goldens from the shipped binaries, golden-recipe weights
    for name in ("c1", "c2", "c4_2"):
        path = os.path.join(REF_SRC, name + ".txt")
        sc = read_scenario_text(path)
        b = ScenarioBatch.from_scenarios([sc])
        for variant, w in (("trp", GOLDEN_W_TRP), ("cub", GOLDEN_W_CUB)):
            w0 = list(w)
            w0[9] = 0.0  # weight_end_l = 0: trp's end term reads l[N-1] out of bounds otherwise (a10)
            c = po.solve_batch(variant, b, w0, mode=1, k_max=32, kind="port")
            K = int(c["K"][0])
            x = c["ctrl"][0][:12 * K]
            r = run_shipped(variant, path, w0, iteration=31, solution=x)
            tag = "%s/%s" % (name, variant)
            samp[tag + "/weights"] = np.array(w0)
            samp[tag + "/ctrl"] = x
            samp[tag + "/retval"] = np.float64(r["retval"])
            rows = np.array([[float(v) for v in ln.split()] for ln in r["out_text"].strip().split("\n")])
            samp[tag + "/traj"] = rows
            print("%-18s shipped retval %.6f rows %d oracle cost %.6f" % (tag, r["retval"], len(rows), c["a_cost"][0]))
    np.savez_compressed(os.path.join(ROOT, "spectral_b200", "data", "fixtures.npz"), **fx)
    np.savez_compressed(os.path.join(OUT, "shipped_qp.npz"), **qps)
    np.savez_compressed(os.path.join(OUT, "ref_segments.npz"), **segs)
    np.savez_compressed(os.path.join(OUT, "converged.npz"), **conv)
    np.savez_compressed(os.path.join(OUT, "shipped_sampling.npz"), **samp)
    for f in ("s1_slt_3d_31.txt", "s1_cub_3d_31.txt"):
        shutil.copy(os.path.join(REF_SRC, f), os.path.join(OUT, f))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
