#!/usr/bin/env python
"""tools/profile_side.py -- every kernel that is NOT the ADMM solver, once, at B = 65 536 (for `ncu --set full -k regex:...`):
k_bounds, k_tables, k_corridor, k_classify, k_finalize, k_argmin, k_ego_states, k_frenet_to_cartesian."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spectral_b200 import api
from spectral_b200.scenarios import GOLDEN_W_CUB, config2, random_obstacles

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
p = api.SpectralPlanner(device=0, max_batch=B, n_max=128, r_max=8, k_max=16)
obs, n_obs = random_obstacles(4096, max_obs=3, seed=20230607)
d_obs = torch.from_numpy(np.tile(obs, (B // 4096, 1, 1))).to(dev)
d_n = torch.from_numpy(np.tile(n_obs, B // 4096)).to(dev)
p.bounds_device(d_obs, 71, 8, n_obs=d_n)
base = config2(4096)
names = ("s_bounds", "l_bounds", "ds_bounds", "dl_bounds", "s_ref", "l_ref", "init", "scalars")
inp = {k: torch.from_numpy(np.ascontiguousarray(np.tile(a, (B // 4096,) + (1,) * (a.ndim - 1)))).to(dev) for k, a in zip(names, base.arrays())}
inp["weights"] = torch.tensor(GOLDEN_W_CUB, dtype=torch.float64, device=dev)
outs = p.alloc_device_outputs(B, samples_cap=72)
opt = api.default_options(max_iter=50, polish=0)   # the solver kernels are not what this run is for
p.solve_device("cub", 71, 2, 0.1, inp, outs, options=opt)
torch.cuda.synchronize()
p.set_timing(True)
for _ in range(3):
    p.solve_device("cub", 71, 2, 0.1, inp, outs, options=opt)
torch.cuda.synchronize()
kt = p.get_timing()
calls = max(kt.get("calls", 1), 1)
cor_ms = kt["corridor"] / calls
cor_bytes = B * (8 * (4 * 2 * 71 + 2 * 71) + 4) + 112 * float(outs["K"].double().sum())
print("k_corridor %.4f ms per %d scenarios, %.1f GB/s algorithmic (%.3f of 6420.7)" % (cor_ms, B, cor_bytes / cor_ms / 1e6, cor_bytes / cor_ms / 1e6 / 6420.7))
p.set_timing(False)
c = torch.zeros(1, dtype=torch.float64, device=dev); i = torch.zeros(1, dtype=torch.int64, device=dev)
p.argmin_device(outs["a_cost"], 0, c, i)
p.ego_states_device(outs["samples"], outs["npts"], torch.zeros(1, dtype=torch.float64, device=dev))
n = B * 72
p.frenet_to_cartesian_device(torch.rand(n, 6, dtype=torch.float64, device=dev), torch.rand(n, 3, dtype=torch.float64, device=dev),
                             torch.rand(n, 3, dtype=torch.float64, device=dev) * 0.1)
torch.cuda.synchronize()
print("K mean", float(outs["K"].double().mean()), "argmin", float(c), int(i))
p.close()
