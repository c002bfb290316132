"""Import harness for the UNMODIFIED reference (remydubois/illico) in the build container.

Only ``tests/golden/make_golden.py`` uses this module, and only where
``/root/reference`` exists (the build container).  Nothing in the GPU tests,
``smoke()`` or ``bench.py`` imports it: the reference does not travel to the
GPU box, the committed fixtures under ``tests/golden/*.npz`` do.

The reference imports ``anndata`` and ``h5py`` at module import time
(``illico/asymptotic_wilcoxon.py:5``, ``illico/utils/registry.py:5-6``) but only
uses them as type keys.  Neither is installed here, so tiny stand-in modules are
put into ``sys.modules`` first.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import pandas as pd

REFERENCE_ROOT = "/root/reference"


class _AnnData:
    """Holder with the four attributes the reference reads (`asymptotic_wilcoxon.py:178-208`)."""

    def __init__(self, X, obs, var_names, layers=None):
        self.X = X
        self.obs = obs
        self.var_names = var_names
        self.layers = layers or {}


def _install_stubs() -> None:
    if "h5py" not in sys.modules:
        h5 = types.ModuleType("h5py")
        h5.Dataset = type("Dataset", (), {})
        sys.modules["h5py"] = h5
    if "anndata" not in sys.modules:
        ad = types.ModuleType("anndata")
        core = types.ModuleType("anndata._core")
        sd = types.ModuleType("anndata._core.sparse_dataset")
        sd._CSCDataset = type("_CSCDataset", (), {})
        sd._CSRDataset = type("_CSRDataset", (), {})
        core.sparse_dataset = sd
        ad._core = core
        ad.AnnData = _AnnData
        sys.modules["anndata"] = ad
        sys.modules["anndata._core"] = core
        sys.modules["anndata._core.sparse_dataset"] = sd


def import_reference():
    """Returns the reference's ``illico`` package, imported from /root/reference."""
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        from loguru import logger

        logger.remove()
    except Exception:  # pragma: no cover
        pass
    import illico  # noqa: F401

    return illico


def make_adata(X, labels, var_names=None, key="pert"):
    n_genes = X.shape[1]
    if var_names is None:
        var_names = [f"gene_{i}" for i in range(n_genes)]
    return _AnnData(X, pd.DataFrame({key: list(labels)}), pd.Index(var_names))


def ref_run(X, labels, reference, *, is_log1p=False, use_continuity=True, tie_correct=True,
            alternative="two-sided", batch_size=None, n_threads=1):
    """Runs the reference's public entry point; returns ``(groups, p, U, fc)`` with [G, N] arrays.

    ``batch_size`` is always an integer: the reference's ``"auto"`` mode skips the
    boundary gene of every split.
    """
    illico = import_reference()
    n_genes = X.shape[1]
    if batch_size is None:
        batch_size = max(n_genes, 1)
    adata = make_adata(X, labels)
    df = illico.asymptotic_wilcoxon(
        adata,
        is_log1p=is_log1p,
        group_keys="pert",
        reference=reference,
        n_threads=n_threads,
        batch_size=int(batch_size),
        alternative=alternative,
        use_continuity=use_continuity,
        tie_correct=tie_correct,
    )
    groups = np.unique(np.asarray(list(labels)))
    G = len(groups)
    arr = df.to_numpy().reshape(G, n_genes, 3)
    return groups, arr[:, :, 0].copy(), arr[:, :, 1].copy(), arr[:, :, 2].copy()
