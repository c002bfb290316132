"""Public entry point: ``asymptotic_wilcoxon`` with the reference's signature and output
(``illico/asymptotic_wilcoxon.py:71-258``), computed on a B200 through ``libillico_b200.so``.

What changed relative to the reference's driver (same contract, different machinery):
  * gene batches go to CUDA streams on one GPU instead of joblib threads (``n_threads`` is accepted and
    used only for host-side readers of backed data; ``precompile`` is a no-op: the kernels are built
    ahead of time);
  * ``batch_size`` is the number of genes per device batch; ``"auto"`` lets the engine size batches to
    its memory budget.  Every gene is computed (the reference's ``"auto"`` split skips one boundary gene
    per split);
  * the one-versus-reference row of the reference group is always ``(p=1, U=-1, fc)`` (the reference's
    dense kernel leaves it uninitialised, ``ovo/dense_ovo.py:116-120``).
"""
from __future__ import annotations

import math
import threading
from queue import Queue
from typing import Literal

import numpy as np
import pandas as pd
import torch
from scipy import sparse

from .engine import CSR, Engine, make_flags, require_cuda
from .groups import GroupContainer, encode_and_count_groups
from .registry import DataHandler, Test, data_handler_registry, dispatcher_registry  # noqa: F401

__all__ = ["asymptotic_wilcoxon"]


def _batches(n_genes: int, batch_size, engine: Engine, in_ram: bool):
    if isinstance(batch_size, (int, np.integer)) and not isinstance(batch_size, bool):
        if batch_size <= 0:
            raise ValueError(f"Invalid batch_size value: {batch_size}. Must be 'auto' or an integer.")
        step = min(int(batch_size), engine.max_batch_genes(n_genes))
    elif batch_size == "auto":
        step = engine.max_batch_genes(n_genes)
        if not in_ram:
            step = min(step, 256)  # backed data: stream ~256 genes at a time like the reference aims for
    else:
        raise ValueError(f"Invalid batch_size value: {batch_size}. Must be 'auto' or an integer.")
    step = max(1, step)
    bounds = list(range(0, n_genes, step)) + [n_genes]
    return list(zip(bounds[:-1], bounds[1:]))


def operator(data_handler: DataHandler, lb: int, ub: int, engine: Engine, flags, results: torch.Tensor,
             fetched=None) -> None:
    """One gene batch: fetch -> device -> stage + rank, written into ``results[:, lb:ub, :]`` on the device
    (the reference's ``operator``, ``asymptotic_wilcoxon.py:29-68``)."""
    n_cols = data_handler.data.shape[1]
    if lb < 0 or ub > n_cols or lb > ub:
        raise ValueError(f"Invalid chunk bounds: {(lb, ub)} for data with {n_cols} columns.")
    if fetched is None:
        fetched = data_handler.fetch(lb, ub)
    data, bounds = fetched
    M = data_handler.to_device(data, engine)
    engine.run_batch(M, bounds[0], bounds[1], flags, results, lb)


def asymptotic_wilcoxon(
    adata,
    is_log1p: bool,
    group_keys: str,
    reference: str | None = None,
    n_threads: int = 1,
    batch_size: int | Literal["auto"] = "auto",
    alternative: str = "two-sided",
    use_continuity: bool = True,
    tie_correct: bool = True,
    layer: str | None = None,
    precompile: bool = True,
    device=None,
    return_array: bool = False,
):
    """Asymptotic Mann-Whitney / Wilcoxon rank-sum tests for every (group, gene) on a B200.

    Same parameters and result as ``illico.asymptotic_wilcoxon``: a ``pd.DataFrame`` indexed by
    ``MultiIndex.from_product([groups, var_names], names=["pert", "feature"])`` with float64 columns
    ``p_value``, ``statistic`` (U of the reference / rest sample) and ``fold_change``.

    Extra keyword arguments (not in the reference): ``device`` (CUDA device, default current) and
    ``return_array`` (return ``(groups, var_names, results[G, N, 3])`` and skip the DataFrame).
    """
    del precompile  # kernels are compiled ahead of time
    X = adata.layers[layer] if layer is not None else adata.X
    if isinstance(X, (sparse.csr_array, sparse.csc_array)):
        X = sparse.csr_matrix(X) if isinstance(X, sparse.csr_array) else sparse.csc_matrix(X)
    data_handler = data_handler_registry.get(X)  # KeyError for unsupported containers, like the reference

    # In-RAM input: start the host->device copy first; it runs (asynchronously for pinned memory) while the
    # host encodes the groups and builds the plan.
    dev = require_cuda(device)
    if data_handler.in_ram:
        with torch.cuda.device(dev):
            data_handler.to_device(X, dev)

    # the reference goes through `.tolist()` + a dict loop (utils/groups.py:42-45); same encoding, vectorised
    unique_raw_groups, grpc = encode_and_count_groups(groups=adata.obs[group_keys], ref_group=reference)
    n_cells, n_genes = X.shape
    if grpc.encoded_groups.size != n_cells:
        raise ValueError(f"{grpc.encoded_groups.size} group labels for {n_cells} cells")

    engine = Engine(grpc, dev)
    fmt = data_handler.kernel_data_format().value
    flags = make_flags(is_log1p, use_continuity, tie_correct, alternative, fmt)
    iterator = _batches(n_genes, batch_size, engine, data_handler.in_ram)

    with torch.cuda.device(engine.device):
        results = torch.empty((engine.n_groups, n_genes, 3), dtype=torch.float64, device=engine.device)
        if data_handler.in_ram:
            M = data_handler.to_device(X, engine)
            if fmt == CSR and not engine.check_csr_sorted(M):
                raise ValueError(
                    "Input data matrix indices are not sorted. This is very unusual and may lead to incorrect results. "
                    "This can be the result of operations like `adata[:, np.random.choice(…)]` that do not preserve sorting."
                    "Please make sure that indices used to chunk the adata or the expression matrix have been sorted "
                    "prior to computing DE genes."
                )
            for lb, ub in iterator:
                operator(data_handler, lb, ub, engine, flags, results)
        else:
            _run_backed(data_handler, iterator, engine, flags, results, max(1, int(n_threads)))
        host = torch.empty(results.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(results, non_blocking=True)
        torch.cuda.current_stream(engine.device).synchronize()
    out = host.numpy()
    if return_array:
        return unique_raw_groups, np.asarray(adata.var_names), out
    cols = pd.Series(adata.var_names, name="feature", dtype=str)
    rows = pd.Series(unique_raw_groups, name="pert", dtype=str)
    return pd.DataFrame(
        data=out.reshape(-1, 3),
        index=pd.MultiIndex.from_product([rows, cols], names=["pert", "feature"]),
        columns=["p_value", "statistic", "fold_change"],
    )


def _run_backed(data_handler: DataHandler, iterator, engine: Engine, flags, results, n_readers: int) -> None:
    """Out-of-core input: reader threads slice the next batches from disk while the GPU ranks the current one."""
    q: Queue = Queue(maxsize=max(2, n_readers))
    it = iter(iterator)
    lock = threading.Lock()
    errors: list = []

    def reader():
        while True:
            with lock:
                nxt = next(it, None)
            if nxt is None:
                q.put(None)
                return
            try:
                q.put((nxt, data_handler.fetch(*nxt)))
            except BaseException as e:  # surfaced on the main thread
                errors.append(e)
                q.put(None)
                return

    threads = [threading.Thread(target=reader, daemon=True) for _ in range(n_readers)]
    for t in threads:
        t.start()
    done = 0
    pending: dict = {}
    order = list(iterator)
    nxt_i = 0
    while done < n_readers:
        item = q.get()
        if item is None:
            done += 1
        else:
            pending[item[0]] = item[1]
        while nxt_i < len(order) and order[nxt_i] in pending:  # keep the gene order deterministic
            lb, ub = order[nxt_i]
            operator(data_handler, lb, ub, engine, flags, results, fetched=pending.pop(order[nxt_i]))
            nxt_i += 1
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    assert nxt_i == len(order)


_ = (math, GroupContainer, Test, dispatcher_registry)
