#!/usr/bin/env python
"""Experiment: ADMM iteration counts / time of the config-2 batch under different adaptive-rho schedules and alpha."""
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
for kw in (dict(adaptive_rho_interval=50), dict(adaptive_rho_interval=25), dict(adaptive_rho_interval=25, adaptive_rho_tolerance=2.0),
           dict(adaptive_rho_interval=50, adaptive_rho_tolerance=2.0), dict(adaptive_rho_interval=200), dict(alpha=1.8), dict(rho=1.0), dict(rho=0.01),
           dict(check_termination=50, adaptive_rho_interval=100)):
    o = api.default_options(**kw)
    r = p.solve("cub", batch, GOLDEN_W_CUB, options=o)
    t = time.time()
    r = p.solve("cub", batch, GOLDEN_W_CUB, options=o)
    dt = time.time() - t
    both = r.verified() & base.verified()
    d = np.abs(r.ctrl[both] - base.ctrl[both])
    tol = 1e-6 + 1e-5 * np.abs(base.ctrl[both])
    bad = (d > tol).any(axis=1).sum() if both.any() else 0
    print("%-70s ok %4d (same class %4d) verified %4d mean iters %6.0f  both %d off-tol %d  %.1f ms" % (
        kw, r.ok().sum(), (r.ok() == base.ok()).sum(), r.verified().sum(), r.iters.mean(), both.sum(), bad, dt * 1e3))
