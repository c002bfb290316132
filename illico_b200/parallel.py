"""Gene sharding across ranks (one process per GPU).

Every statistic is per gene, so the path shards naturally with no data-path collective
(SURVEY.md section 8e): rank r ranks genes ``gene_shard(N, r, world)`` and the per-rank result
slabs ``[G, n_r, 3]`` are gathered once at the end (``torch.distributed`` all_gather: NCCL between
GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def gene_shard(n_genes: int, rank: int, world: int, weights=None) -> tuple[int, int]:
    """Contiguous gene range of ``rank``.  ``weights`` (e.g. non-zeros per gene) balances sparse input."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    if weights is None:
        base, rem = divmod(n_genes, world)
        lb = rank * base + min(rank, rem)
        return lb, lb + base + (1 if rank < rem else 0)
    w = np.asarray(weights, dtype=np.float64)
    if w.size != n_genes:
        raise ValueError("one weight per gene expected")
    cum = np.concatenate([[0.0], np.cumsum(w + 1e-9)])
    targets = cum[-1] * np.arange(world + 1) / world
    cuts = np.searchsorted(cum, targets, side="left")
    cuts[0], cuts[-1] = 0, n_genes
    cuts = np.maximum.accumulate(cuts)
    return int(cuts[rank]), int(cuts[rank + 1])


def gather_results(local, n_genes: int, weights=None):
    """All-gathers per-rank ``[G, n_r, 3]`` float64 tensors into ``[G, n_genes, 3]`` on every rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    shards = [gene_shard(n_genes, r, world, weights) for r in range(world)]
    G = local.shape[0]
    assert local.shape[1] == shards[rank][1] - shards[rank][0]
    widest = max(ub - lb for lb, ub in shards)
    pad = torch.zeros((G, widest, 3), dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.empty((G, n_genes, 3), dtype=local.dtype, device=local.device)
    for (lb, ub), buf in zip(shards, bufs):
        out[:, lb:ub] = buf[:, : ub - lb]
    return out
