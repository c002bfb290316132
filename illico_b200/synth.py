"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py.

The shapes follow BASELINE.json's configs and SURVEY.md section 8(d); the small
fixture mirrors the reference's own ``rand_adata`` fixture
(reference ``tests/conftest.py:76-100``): Poisson counts with gene-specific
means, ~50 % masked to zero, ``RandomState(0)``, float32.
"""
from __future__ import annotations

import numpy as np

CONTROL = "non-targeting"


def conftest_fixture(seed: int = 0, n_cells: int = 10_000, n_genes: int = 15, n_groups: int = 5,
                     sparsity: float = 0.5, dtype=np.float32):
    """Same draw order as the reference fixture (``tests/conftest.py:82-100``)."""
    rng = np.random.RandomState(seed)
    gene_means = rng.uniform(0.1, 15, size=n_genes)
    dense = rng.poisson(gene_means, size=(n_cells, n_genes)).astype(dtype)
    mask = rng.rand(n_cells, n_genes) < sparsity
    dense[mask] = 0
    groups = rng.randint(0, n_groups, size=n_cells)
    labels = [f"pert_{g}" for g in groups]
    return dense, labels


def perturbation_labels(rng: np.random.RandomState, n_cells: int, n_perts: int, p_control: float = 0.036):
    """K562-like labels: P(control) = 0.036, the rest uniform over ``n_perts`` perturbations."""
    is_ctrl = rng.rand(n_cells) < p_control
    pert = rng.randint(0, n_perts, size=n_cells)
    width = max(4, len(str(n_perts - 1)))
    names = np.array([f"p{i:0{width}d}" for i in range(n_perts)] + [CONTROL])
    codes = np.where(is_ctrl, n_perts, pert)
    return names[codes].tolist(), codes


def k562_like(seed: int, n_cells: int, n_genes: int, n_perts: int, lam: float = 1.0, masked: float = 0.85,
              continuous: bool = False, dtype=np.float32):
    """Integer counts Poisson(lam) with ``masked`` of the entries zeroed (~9.5 % nnz at the defaults).

    ``continuous=True`` gives the stress variant: log1p of library-size-normalised
    counts, i.e. almost no ties among the non-zeros.
    """
    rng = np.random.RandomState(seed)
    labels, _ = perturbation_labels(rng, n_cells, n_perts)
    X = rng.poisson(lam, size=(n_cells, n_genes)).astype(dtype)
    X[rng.rand(n_cells, n_genes) < masked] = 0
    if continuous:
        lib = X.sum(axis=1, keepdims=True) + rng.uniform(0.5, 1.5, size=(n_cells, 1)).astype(dtype)
        X = np.log1p(X / lib * dtype(1.0e4)).astype(dtype)
    return X, labels


def cluster_labels(seed: int, n_cells: int, n_clusters: int):
    rng = np.random.RandomState(seed)
    width = max(2, len(str(n_clusters - 1)))
    return [f"c{g:0{width}d}" for g in rng.randint(0, n_clusters, size=n_cells)]


def k562_like_torch(seed: int, n_cells: int, n_genes: int, lam: float = 1.0, masked: float = 0.85,
                    device="cuda", out=None, chunk_rows: int = 16384):
    """Device-side generator of the K562-shape count matrix for bench.py (numpy is too slow at 2.4e9 draws).

    Not bit-identical to :func:`k562_like` (different RNG); the distribution is the same.
    """
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if out is None:
        out = torch.empty((n_cells, n_genes), dtype=torch.float32, device=device)
    for r0 in range(0, n_cells, chunk_rows):
        r1 = min(n_cells, r0 + chunk_rows)
        rate = torch.full((r1 - r0, n_genes), lam, dtype=torch.float32, device=device)
        x = torch.poisson(rate, generator=g)
        keep = torch.rand((r1 - r0, n_genes), device=device, generator=g) >= masked
        out[r0:r1] = (x * keep).to(out.device, non_blocking=True)
    return out
