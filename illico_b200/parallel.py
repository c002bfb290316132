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


class _ColumnView:
    """The four attributes ``asymptotic_wilcoxon`` reads, restricted to genes ``[lb, ub)`` of an AnnData-like object."""

    def __init__(self, adata, lb: int, ub: int, layer):
        X = adata.layers[layer] if layer is not None else adata.X
        self.X = X[:, lb:ub]
        self.layers = {}
        self.obs = adata.obs
        self.var_names = adata.var_names[lb:ub]


def asymptotic_wilcoxon_sharded(adata, is_log1p: bool, group_keys: str, reference=None, *, layer=None, balance_by_nnz=True,
                                return_array: bool = False, **kwargs):
    """``asymptotic_wilcoxon`` with the genes sharded across the ranks of an initialised ``torch.distributed`` job
    (one process per GPU, every rank holding the same ``adata``): rank r ranks its contiguous gene range on its own
    GPU, the ``[G, n_r, 3]`` slabs are all-gathered once (NCCL over NVLink) and every rank returns the full result.
    There is no collective in the data path.  Without an initialised process group this is the
    single-GPU call."""
    import torch
    import torch.distributed as dist
    from scipy import sparse

    from .asymptotic_wilcoxon import _result_frame, asymptotic_wilcoxon

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return asymptotic_wilcoxon(adata, is_log1p, group_keys, reference, layer=layer, return_array=return_array, **kwargs)
    world, rank = dist.get_world_size(), dist.get_rank()
    X = adata.layers[layer] if layer is not None else adata.X
    n_genes = X.shape[1]
    weights = None
    if balance_by_nnz and sparse.issparse(X):   # same cut points on every rank: they all hold the same matrix
        weights = np.diff(X.tocsc().indptr) if not sparse.isspmatrix_csc(X) else np.diff(X.indptr)
    lb, ub = gene_shard(n_genes, rank, world, weights)
    view = _ColumnView(adata, lb, ub, layer)
    groups, _names, local = asymptotic_wilcoxon(view, is_log1p, group_keys, reference, return_array=True, **kwargs)
    dev = torch.device("cuda", torch.cuda.current_device())
    full = gather_results(torch.from_numpy(np.ascontiguousarray(local)).to(dev), n_genes, weights).cpu().numpy()
    if return_array:
        return groups, np.asarray(adata.var_names), full
    return _result_frame(groups, adata.var_names, full)
