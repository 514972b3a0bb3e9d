#!/usr/bin/env python
"""tools/sanitize_cases.py -- one config-2 scenario per segment count K = 3..12 (all dense solver classes), a c7 fixture
(N = 101), and the round-2 kernels (shared-KKT tiles, K > 16, weight sweep, k_bounds, downstream) through the host API, for compute-sanitizer runs:
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
# ---- round-2 kernels: shared-KKT tiles (k_qps_*), K > 16 (k_qp<32,2>), weight sweep, k_bounds, downstream kernels
import torch
from spectral_b200.scenarios import GOLDEN_W_TRP, config3, random_obstacles, zigzag_breaks
dev = torch.device("cuda", 0)
r = p.solve("trp", config3(16, groups=2), WEIGHTS_FILE, options=api.default_options(max_iter=200, shared_kkt=1))
print("shared-KKT K", r.K.tolist(), "status", r.status.tolist(), "iters", r.iters.tolist())
zz = zigzag_breaks(load_fixture("c7"), 4)
r = p.solve("trp", zz, WEIGHTS_FILE, options=api.default_options(max_iter=100))
print("K>16 K", r.K.tolist(), "status", r.status.tolist())
one = ScenarioBatch.from_scenarios([load_fixture("c1")])
w = np.random.default_rng(1).uniform(0.0, 50.0, (8, 10))
r = p.solve_weights("trp", one, w, options=opt)
print("weight sweep status", r.status.tolist())
obs, n_obs = random_obstacles(16, max_obs=4, seed=3)
sb, lb, nl = p.bounds_device(torch.from_numpy(obs).to(dev), 71, 8, n_obs=torch.from_numpy(n_obs).to(dev))
print("k_bounds lanes", nl.cpu().tolist())
got = p.solve("cub", sub, GOLDEN_W_CUB, samples_cap=96, options=opt)
st = p.ego_states_device(torch.from_numpy(got.samples).to(dev), torch.from_numpy(got.npts).to(dev), torch.zeros(1, dtype=torch.float64, device=dev))
f2c = p.frenet_to_cartesian_device(torch.rand(100, 6, dtype=torch.float64, device=dev), torch.rand(100, 3, dtype=torch.float64, device=dev),
                                   torch.rand(100, 3, dtype=torch.float64, device=dev) * 0.1)
torch.cuda.synchronize()
print("downstream", tuple(st.shape), tuple(f2c.shape))
p.close()
