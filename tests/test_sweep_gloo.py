"""CPU tests (`-m "not gpu"`): the N > 1 host logic of a sharded sweep on world_size-2 gloo.

The path shards by scenario with no data-path collective; the only exchange is the final
(cost, global index) arg-min gather + winner broadcast (spectral_b200/sweep.py).  Here the per-rank
"solver" is a stand-in cost table (the CUDA path needs a GPU); what is tested is sharding, the exchange,
the deterministic tie-break and the independence of the result from the number of ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectral_b200.sweep import broadcast_winner, gather_best, pack_record, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _costs(total):
    rng = np.random.default_rng(7)
    c = np.round(rng.uniform(10.0, 20.0, total), 1)      # many exact ties
    c[rng.random(total) < 0.3] = 100000000000.0           # failed scenarios carry the sentinel
    c[[5, total - 3]] = 1.25                               # the minimum, twice: lowest index must win
    return c


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(total, rank, world)
        costs = _costs(total)[lo:hi]
        j = int(np.argmin(costs)) if hi > lo else 0      # first minimum = lowest local index
        cost = torch.tensor([costs[j] if hi > lo else 1e300], dtype=torch.float64)
        idx = torch.tensor([lo + j], dtype=torch.int64)
        best_cost, best_idx, owner = gather_best(cost, idx)
        rec = torch.full((8,), float(rank), dtype=torch.float64)
        if rank == owner:
            rec[0] = best_cost
        broadcast_winner(rec, owner)
        q.put((rank, best_cost, best_idx, owner, float(rec[0]), float(rec[1])))
    finally:
        dist.destroy_process_group()


def _run(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 1024, 1048576):
        for world in (1, 2, 3, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_pack_record_is_bit_exact():
    c = torch.tensor([1.0000000000000002], dtype=torch.float64)
    i = torch.tensor([(1 << 53) + 1], dtype=torch.int64)    # not representable as a double
    rec = pack_record(c, i)
    assert rec.dtype == torch.int64 and int(rec[1]) == (1 << 53) + 1
    assert rec[0:1].view(torch.float64).item() == 1.0000000000000002


def test_argmin_gather_world2_gloo():
    total = 1001
    res = _run(2, total)
    costs = _costs(total)
    want_idx = int(np.argmin(costs))
    assert want_idx == 5
    for rank, best_cost, best_idx, owner, rec0, rec1 in res:
        assert best_cost == costs[want_idx] and best_idx == want_idx and owner == 0
        assert rec0 == best_cost and rec1 == float(owner)      # every rank holds the owner's record


def test_argmin_result_independent_of_world_size():
    total = 257
    r1 = gather_best(torch.tensor([_costs(total).min()], dtype=torch.float64),
                     torch.tensor([int(np.argmin(_costs(total)))], dtype=torch.int64))
    r3 = _run(3, total)
    assert all((x[1], x[2]) == (r1[0], r1[1]) for x in r3)
