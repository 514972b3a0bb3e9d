#!/usr/bin/env python
"""Experiment: how loose can the ADMM termination be before the polish (active-set correction rounds) stops finding the
KKT-verified optimum?  Prints, per eps, solved / verified counts, mean iterations and the distance to the baseline optimum."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spectral_b200 import api
from spectral_b200.scenarios import GOLDEN_W_CUB, config2

B = 1024
batch = config2(B)
p = api.SpectralPlanner(device=0, max_batch=B, n_max=128, r_max=8, k_max=16)
base = p.solve("cub", batch, GOLDEN_W_CUB)
print("baseline: ok %d verified %d mean iters %.0f" % (base.ok().sum(), base.verified().sum(), base.iters.mean()))
for eps, rounds, maxit in ((1e-4, 8, 5000), (1e-3, 8, 5000), (1e-3, 16, 5000), (1e-2, 16, 5000), (1e-3, 16, 1000), (1e-2, 16, 500), (1e-1, 16, 500)):
    o = api.default_options(eps_abs=eps, eps_rel=eps, polish_rounds=rounds, max_iter=maxit)
    t = time.time()
    r = p.solve("cub", batch, GOLDEN_W_CUB, options=o)
    dt = time.time() - t
    both = r.verified() & base.verified()
    d = np.abs(r.ctrl[both] - base.ctrl[both])
    tol = 1e-6 + 1e-5 * np.abs(base.ctrl[both])
    bad = (d > tol).any(axis=1).sum() if both.any() else 0
    newv = (r.verified() & ~base.verified()).sum()
    lost = (~r.verified() & base.verified()).sum()
    print("eps %.0e rounds %2d max_iter %4d: ok %4d verified %4d (new %d lost %d) mean iters %6.0f  both %d off-tolerance %d maxdiff %.2e  %.1f ms" % (
        eps, rounds, maxit, r.ok().sum(), r.verified().sum(), newv, lost, r.iters.mean(), both.sum(), bad, d.max() if both.any() else 0, dt * 1e3))
