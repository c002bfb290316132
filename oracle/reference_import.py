"""Import harness for the UNMODIFIED reference (remydubois/illico).  TEST INFRASTRUCTURE, like the rest of ``oracle/``.

The reference is pure Python + numba.  It is importable in two places:
  * ``/root/reference`` -- the build container (golden-vector generation, ``tests/golden/make_golden.py``);
  * ``baseline/_ref``   -- a ``pip install --target`` of that tree (git-ignored, travels to the GPU box with the
    snapshot), which lets ``bench.py --impl reference`` time the real numba kernels on the GPU box's host cores and
    lets the GPU dispatchers be exercised under the reference's own driver (``tests/test_reference_driver.py``).
    Recipe: the tree's build backend (poetry-core) is not installed offline, so the install runs on
    a copy under /tmp whose packaging file is replaced by a three-line setuptools ``setup.py``; the package sources are
    untouched.

The reference imports ``anndata`` and ``h5py`` at module import time (``illico/asymptotic_wilcoxon.py:5``,
``illico/utils/registry.py:5-6``) but only uses them as type keys; neither is installed, so tiny stand-in modules are
put into ``sys.modules`` first.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ("/root/reference", os.path.join(ROOT, "baseline", "_ref"))


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


def reference_root():
    for c in CANDIDATES:
        if os.path.isdir(os.path.join(c, "illico")):
            return c
    return None


def import_reference():
    """Returns the reference's ``illico`` package; raises ImportError when it is not available here."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference is neither at /root/reference nor installed under baseline/_ref")
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
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
            alternative="two-sided", batch_size=None, n_threads=1, precompile=True):
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
        precompile=precompile,
    )
    groups = np.unique(np.asarray(list(labels)))
    G = len(groups)
    arr = df.to_numpy().reshape(G, n_genes, 3)
    return groups, arr[:, :, 0].copy(), arr[:, :, 1].copy(), arr[:, :, 2].copy()
