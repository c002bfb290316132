"""Test helpers: a minimal AnnData stand-in (anndata is not installed here) and result reshaping."""
import numpy as np
import pandas as pd


class FakeAnnData:
    """The four attributes ``asymptotic_wilcoxon`` reads (reference asymptotic_wilcoxon.py:178-208)."""

    def __init__(self, X, labels, var_names=None, key="pert", layers=None):
        self.X = X
        self.obs = pd.DataFrame({key: list(labels)})
        n_genes = X.shape[1]
        self.var_names = pd.Index(var_names if var_names is not None else [f"gene_{i}" for i in range(n_genes)])
        self.layers = layers or {}


def planes(df_or_arr, n_groups, n_genes):
    arr = df_or_arr.to_numpy() if hasattr(df_or_arr, "to_numpy") else np.asarray(df_or_arr)
    arr = arr.reshape(n_groups, n_genes, 3)
    return arr[:, :, 0], arr[:, :, 1], arr[:, :, 2]
