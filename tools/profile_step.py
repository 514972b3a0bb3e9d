#!/usr/bin/env python
"""tools/profile_step.py [B] [steps] [k_max] -- the bench workload (config 2: cub, obstacle-perturbed copies of c1)
through the device-resident C-ABI entry on one stream, nothing else: the target of the ncu captures under profiles/.
    ncu --set full --clock-control none --import-source on -k regex:k_qpd -c 3 -o gpurun_out/qpd python tools/profile_step.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spectral_b200 import api  # noqa: E402
from spectral_b200.scenarios import GOLDEN_W_CUB, config2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
k_max = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = torch.device("cuda", 0)
batch = config2(B)
names = ("s_bounds", "l_bounds", "ds_bounds", "dl_bounds", "s_ref", "l_ref", "init", "scalars")
inputs = {k: torch.from_numpy(np.ascontiguousarray(a)).to(dev) for k, a in zip(names, batch.arrays())}
inputs["weights"] = torch.tensor(GOLDEN_W_CUB, dtype=torch.float64, device=dev)
planner = api.SpectralPlanner(device=0, max_batch=B, n_max=128, r_max=8, k_max=k_max)
outs = planner.alloc_device_outputs(B)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(steps):
    e0.record()
    planner.solve_device("cub", batch.n_knots, batch.n_regions, batch.delta_t, inputs, outs)
    e1.record()
    torch.cuda.synchronize()
    print("step %d: %.3f ms, solved %d of %d, mean iters %.0f" % (
        i, e0.elapsed_time(e1), int((outs["status"] <= 1).sum()), B, float(outs["iters"].double().mean())))
planner.close()
