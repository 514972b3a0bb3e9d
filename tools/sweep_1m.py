#!/usr/bin/env python
"""tools/sweep_1m.py [TOTAL] -- BASELINE.json configs[4] at full size on whatever GPUs this process group has (1 rank when run
plainly): TOTAL (default 1 048 576) synthetic scenarios in contiguous shards (sweep.shard_range), solved in chunks through the
pipelined host API, local arg-min per rank (k_argmin), arg-min gather across ranks.  Prints solves/s and the winner; with
one rank it also re-derives the winner from 8 simulated shards to show that the result does not depend on the rank count."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spectral_b200 import api, sweep
from spectral_b200.scenarios import GOLDEN_W_CUB, config2

TOTAL = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
CHUNK = 16384
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
lo, hi = sweep.shard_range(TOTAL, rank, world)
planners = [api.SpectralPlanner(device=local, max_batch=CHUNK, n_max=128, r_max=8, k_max=16) for _ in range(3)]
best = (np.inf, -1)
solved = 0
costs = np.empty(hi - lo)
t0 = time.time()
pending = {}
starts = list(range(lo, hi, CHUNK))


def collect(j):
    global best, solved
    s, n = pending.pop(j)
    r = planners[j].wait()
    costs[s - lo:s - lo + n] = r.a_cost[:n]
    solved += int(r.ok()[:n].sum())
    i = int(np.argmin(r.a_cost[:n]))
    if (r.a_cost[i], s + i) < best:
        best = (float(r.a_cost[i]), s + i)


for c, s in enumerate(starts):
    j = c % len(planners)
    if j in pending:
        collect(j)
    n = min(CHUNK, hi - s)
    planners[j].solve_async("cub", config2(n, first=s), GOLDEN_W_CUB)
    pending[j] = (s, n)
for j in list(pending):
    collect(j)
dt = time.time() - t0
dev = torch.device("cuda", local)
cost_t = torch.tensor([best[0]], dtype=torch.float64, device=dev)
idx_t = torch.tensor([best[1]], dtype=torch.int64, device=dev)
bc, bi, owner = sweep.gather_best(cost_t, idx_t)
if rank == 0:
    print("sweep: %d scenarios on %d rank(s): %.1f s, %.0f solves/s per rank (scenario generation on the host included), solved %.3f" % (
        TOTAL, world, dt, (hi - lo) / dt, solved / (hi - lo)))
    print("winner: scenario %d (rank %d), a_cost %.9f" % (bi, owner, bc))
    if world == 1:
        shards = [sweep.shard_range(TOTAL, r, 8) for r in range(8)]
        cand = [min((float(costs[a:b].min()), a + int(np.argmin(costs[a:b])))) if False else (float(costs[a:b].min()), a + int(np.argmin(costs[a:b]))) for a, b in shards]
        w = min(cand)
        assert w == (bc, bi), (w, bc, bi)
        print("8 simulated shards give the same winner: %s" % (w,))
if world > 1:
    torch.distributed.destroy_process_group()
