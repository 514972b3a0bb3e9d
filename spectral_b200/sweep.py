"""Multi-GPU scenario sweeps: one process per GPU, contiguous scenario shards, no data-path
collective; the only exchange is the final best-trajectory arg-min:

    local (min a_cost, lowest global index)            spectral_argmin_device()   [CUDA kernel]
    all-gather of one 16-byte (cost, index) record     torch.distributed (NCCL over NVLink; gloo in CPU tests)
    broadcast of the winner's K / segments / control points from the owning rank

Ties are broken on the lowest global scenario index, so the result does not depend on the number of
ranks.  Failed scenarios carry the reference's sentinel cost 1e11 (trp_wrapper.cpp:199) and lose.
The functions below are backend-agnostic (tensors live wherever the process group's backend wants them).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` scenarios for `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_record(cost: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """(float64 cost, int64 index) -> 2 x int64 record; the index travels bit-exactly (no float rounding)."""
    return torch.stack([cost.reshape(()).view(torch.int64), index.reshape(()).to(torch.int64)])


def gather_best(cost: torch.Tensor, index: torch.Tensor, group: Optional[dist.ProcessGroup] = None
                ) -> Tuple[float, int, int]:
    """All-gather every rank's local best and reduce: returns (best cost, best global index, owner rank).
    `cost`: float64[1], `index`: int64[1] (global scenario index), both on the backend's device."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(cost.item()), int(index.item()), 0
    world = dist.get_world_size(group)
    rec = pack_record(cost, index)
    allrec = torch.empty(world * 2, dtype=torch.int64, device=rec.device)
    dist.all_gather_into_tensor(allrec, rec, group=group)
    allrec = allrec.view(world, 2).cpu()
    costs = allrec[:, 0].contiguous().view(torch.float64)
    idxs = allrec[:, 1]
    best, owner = None, 0
    for r in range(world):
        c, i = float(costs[r]), int(idxs[r])
        if best is None or c < best[0] or (c == best[0] and i < best[1]):
            best, owner = (c, i), r
    return best[0], best[1], owner


def broadcast_winner(record: torch.Tensor, owner: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Broadcast the winner's packed trajectory record (K, segments, control points; <= 2 KB) from `owner`."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(record, src=owner, group=group)
    return record
