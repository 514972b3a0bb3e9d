#!/usr/bin/env python
"""tools/sanitize_cases.py -- one config-2 scenario per segment count K = 3..12 (all dense solver classes) plus a c7 fixture
(N = 101) through the host API, for compute-sanitizer runs:
    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spectral_b200 import api
from spectral_b200.scenarios import GOLDEN_W_CUB, WEIGHTS_FILE, config2, load_fixture
from spectral_b200.wire import ScenarioBatch

b = config2(1024)
idx = [8, 31, 38, 16, 1, 0, 7, 6, 37, 579]  # K = 3 .. 12
sub = ScenarioBatch(b.n_knots, b.n_regions, b.delta_t, *[x[idx] for x in b.arrays()])
p = api.SpectralPlanner(device=0, max_batch=16, n_max=128, r_max=8, k_max=32)
opt = api.default_options(max_iter=int(os.environ.get("SANITIZE_MAX_ITER", "400")))
r = p.solve("cub", sub, GOLDEN_W_CUB, options=opt)
print("cub K", r.K.tolist(), "status", r.status.tolist(), "iters", r.iters.tolist())
r = p.solve("cub", sub, GOLDEN_W_CUB, options=api.default_options(max_iter=100, infeasibility_precheck=1))
print("precheck status", r.status.tolist(), "iters", r.iters.tolist())
r = p.solve("trp", ScenarioBatch.from_scenarios([load_fixture("c7"), load_fixture("c1")]) if False else ScenarioBatch.from_scenarios([load_fixture("c1")]), WEIGHTS_FILE, options=opt)
print("trp c1 K", r.K.tolist(), "status", r.status.tolist())
r = p.solve("trp", ScenarioBatch.from_scenarios([load_fixture("c7")]), WEIGHTS_FILE, options=opt)
print("trp c7 K", r.K.tolist(), "status", r.status.tolist())
p.close()
