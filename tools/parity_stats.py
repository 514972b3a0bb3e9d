#!/usr/bin/env python
"""tools/parity_stats.py -- diagnostic (GPU box): solved/failed/verified statistics of the CUDA path vs the CPU
oracle on the config-2 batch.  Prints one JSON line; used to write the parity section of DESIGN.md."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import helpers as H  # noqa: E402
import pyoracle as po  # noqa: E402
from spectral_b200 import api  # noqa: E402
from spectral_b200.scenarios import GOLDEN_W_CUB, GOLDEN_W_TRP, config2, load_fixture, perturbed_obstacles  # noqa: E402


def stats(variant, batch, w, planner):
    got = planner.solve(variant, batch, w)
    ref, ref0 = H.oracle_pair(variant, batch, w)
    ok0, ok1 = ref0["status"] <= 1, ref["status"] <= 1
    dec = H.decided_classes(ref, ref0)
    mism = got.ok() != ok0
    both = got.verified() & ok1 & (ref["polish"] == 2)
    d = np.abs(got.ctrl[both] - ref["ctrl"][both])
    rel = d / (1e-6 + 1e-5 * np.abs(ref["ctrl"][both]))
    worst = []
    if rel.size:
        per = rel.max(axis=1)
        idx = np.nonzero(both)[0]
        for j in np.argsort(-per)[:5]:
            b = int(idx[j])
            worst.append(dict(b=b, ratio=float(per[j]), absdiff=float(d[j].max()), K=int(got.K[b]), gpu_iters=int(got.iters[b]),
                              ref0_iters=int(ref0["iters"][b]), ref1_iters=int(ref["iters"][b]), gpu_status=int(got.status[b]),
                              obj_gpu=float(got.obj[b]), obj_ref=float(ref["obj"][b])))
    it_same = (got.iters == ref0["iters"])
    return dict(B=int(batch.batch), gpu_ok=int(got.ok().sum()), ref0_ok=int(ok0.sum()), ref1_ok=int(ok1.sum()),
                decided=int(dec.sum()), mismatch_decided=int((mism & dec).sum()), mismatch_undecided=int((mism & ~dec).sum()),
                gpu_verified=int(got.verified().sum()), ref_verified=int((ok1 & (ref["polish"] == 2)).sum()),
                verified_both=int(both.sum()), max_tol_ratio=float(rel.max()) if rel.size else 0.0, worst=worst,
                solved0=int((got.status == 0).sum()), solved0_verified=int(((got.status == 0) & got.verified()).sum()),
                iters_equal=int(it_same.sum()), iters_close=int((np.abs(got.iters - ref0["iters"]) <= 100).sum()),
                gpu_mean_iters=float(got.iters.mean()), ref0_mean_iters=float(ref0["iters"].mean()),
                gpu_status_hist=np.bincount(got.status, minlength=6).tolist(),
                ref0_status_hist=np.bincount(ref0["status"], minlength=6).tolist())


if __name__ == "__main__":
    planner = api.SpectralPlanner(device=0, max_batch=4096, n_max=128, r_max=8, k_max=32)
    out = {"config2_cub_1024": stats("cub", config2(1024), GOLDEN_W_CUB, planner),
           "c1_trp_512": stats("trp", perturbed_obstacles(load_fixture("c1"), 512, seed=77), GOLDEN_W_TRP, planner)}
    print(json.dumps(out))
