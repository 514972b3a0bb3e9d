#!/usr/bin/env python
"""tools/bench_configs.py -- throughput of the other BASELINE.json workloads through the host API (spectral_solve_batch:
H2D + kernels + D2H, synchronous), for DESIGN.md; the headline bench (config 2) is bench.py.
    config 3: scenario_2 (c2.txt), trapezoid-prism, 65 536 variants in 8 shared-KKT groups
    config 4: mixed trp + cub, heterogeneous segment counts (5 bases x 2 variants), 40 960 scenarios"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spectral_b200 import api
from spectral_b200.scenarios import WEIGHTS_FILE, config3, mixed_batches

p = api.SpectralPlanner(device=0, max_batch=65536, n_max=128, r_max=8, k_max=32)
b3 = config3(65536, groups=8)
p.solve("trp", b3.slice(0, 4096), WEIGHTS_FILE)
t = time.time(); r = p.solve("trp", b3, WEIGHTS_FILE); dt = time.time() - t
print("config 3: B=65536 trp, 8 groups: %.0f solves/s (%.2f s), solved %.3f, verified %.3f of solved, mean iters %.0f, K in %s" % (
    65536 / dt, dt, r.ok().mean(), (r.verified() & r.ok()).sum() / max(r.ok().sum(), 1), r.iters.mean(), sorted(set(r.K.tolist()))))
mb = mixed_batches(40960, seed=20230602)
tot = 0; dt = 0.0; ok = 0; its = 0; Ks = set()
for variant, b in mb:
    t = time.time(); r = p.solve(variant, b, WEIGHTS_FILE); dt += time.time() - t
    tot += b.batch; ok += int(r.ok().sum()); its += int(r.iters.sum()); Ks |= set(r.K.tolist())
print("config 4: %d mixed scenarios in %d calls: %.0f solves/s (%.2f s), solved %.3f, mean iters %.0f, K in [%d, %d]" % (
    tot, len(mb), tot / dt, dt, ok / tot, its / tot, min(k for k in Ks if k > 0), max(Ks)))
