"""Named, seeded parity cases.  ``make_golden.py`` runs the reference on them here;
the tests regenerate the same inputs from the seeds and compare with the stored outputs.
"""
from __future__ import annotations

import itertools

import numpy as np

from illico_b200 import synth

ALTERNATIVES = ("two-sided", "less", "greater")


def _conftest():
    X, labels = synth.conftest_fixture()
    return X, labels, labels[0]  # reference = first cell's label (reference tests/test_asymptotic_wilcoxon.py:116-117)


def _k562_mini():
    X, labels = synth.k562_like(seed=11, n_cells=20_000, n_genes=48, n_perts=100)
    return X, labels, synth.CONTROL


def _k562_mini_cont():
    X, labels = synth.k562_like(seed=12, n_cells=6_000, n_genes=24, n_perts=30, continuous=True)
    return X, labels, synth.CONTROL


def _bign():
    """n = 300k: the zero block alone exceeds 2**53 in t**3 - t, so the f64 tie sum is order dependent
    (hard part 1).  Group 'a' is 60 % so the OVO pair (a, b) exceeds 208 063 cells."""
    rng = np.random.RandomState(5)
    n = 300_000
    X = np.empty((n, 4), dtype=np.float32)
    X[:, 0] = rng.poisson(1.0, n) * (rng.rand(n) >= 0.85)
    X[:, 1] = rng.poisson(5.0, n) * (rng.rand(n) >= 0.5)
    X[:, 2] = np.round(rng.gamma(2.0, 2.0, n), 1) * (rng.rand(n) >= 0.3)  # many small tie runs after a zero block
    X[:, 3] = rng.poisson(40.0, n)  # no zeros at all
    u = rng.rand(n)
    labels = np.where(u < 0.6, "a", np.where(u < 0.9, "b", "c")).tolist()
    return X, labels, "a"


def _negatives():
    """Dense-only: centred/scaled data with negative values, exact zeros and -0.0."""
    rng = np.random.RandomState(7)
    n = 5_000
    X = np.round(rng.randn(n, 6) * 2.0, 1).astype(np.float32)
    X[rng.rand(n, 6) < 0.2] = 0.0
    X[rng.rand(n, 6) < 0.02] = -0.0
    X[:, 5] = rng.randn(n).astype(np.float32)  # no ties
    labels = [f"g{v}" for v in rng.randint(0, 7, size=n)]
    return X, labels, "g3"


def _edge():
    """All-zero gene, constant gene, single non-zero, size-1 and size-2 groups."""
    rng = np.random.RandomState(9)
    n = 2_000
    X = rng.poisson(2.0, size=(n, 8)).astype(np.float32)
    X[rng.rand(n, 8) < 0.6] = 0
    X[:, 0] = 0.0
    X[:, 1] = 3.0
    X[:, 2] = 0.0
    X[17, 2] = 5.0
    X[:, 3] = rng.rand(n).astype(np.float32) + 1.0
    codes = rng.randint(0, 6, size=n)
    codes[0] = 6  # size-1 group
    codes[1:3] = 7  # size-2 group
    labels = [f"g{v}" for v in codes]
    return X, labels, "g0"


def _log1p32():
    X, labels = synth.conftest_fixture(seed=3, n_cells=3_000, n_genes=10, n_groups=4)
    return np.log1p(X).astype(np.float32), labels, labels[0]


def _log1p64():
    X, labels = synth.conftest_fixture(seed=3, n_cells=3_000, n_genes=10, n_groups=4)
    return np.log1p(X.astype(np.float64)), labels, labels[0]


def _batched():
    """More than 256 genes, integer batch size smaller than N (reference asymptotic_wilcoxon.py:217-220)."""
    X, labels = synth.conftest_fixture(seed=4, n_cells=1_500, n_genes=300, n_groups=6)
    return X, labels, labels[0]


def _highcount():
    """Dense, highly expressed genes: dozens of distinct values per gene, every group above the thread tier of the
    general OVO kernel (search table, warp tier), far more than the fused path's 12-slot table (handed back), next to
    plain count genes that stay on the fused path."""
    rng = np.random.RandomState(13)
    n = 6_000
    X = (rng.poisson(1.0, size=(n, 16)) * (rng.rand(n, 16) >= 0.8)).astype(np.float32)
    for j, lam in ((1, 30.0), (6, 80.0), (7, 12.0), (12, 300.0)):
        X[:, j] = rng.poisson(lam, n)
    X[:, 9] = rng.poisson(25.0, n) * (rng.rand(n) >= 0.5)
    sizes = np.array([900, 40, 700, 33, 260, 150, 5, 1100])          # the reference group is the largest
    codes = np.repeat(np.arange(sizes.size), sizes)
    codes = np.concatenate([codes, rng.randint(0, sizes.size, size=n - codes.size)])
    rng.shuffle(codes)
    labels = [f"g{c}" for c in codes]
    return X, labels, "g7"


def _c4_clusters():
    """BASELINE config 4 structure: one-versus-rest over 50 clusters, every cluster (1 200-1 400 cells) cut into several
    plan segments (multi-segment groups on the fused paths, multi-segment records of the table kernels)."""
    rng = np.random.RandomState(17)
    n, N = 65_000, 20
    X = (rng.poisson(1.0, size=(n, N)) * (rng.rand(n, N) >= 0.85)).astype(np.float32)
    X[:, 4] = rng.poisson(20.0, n) * (rng.rand(n) >= 0.3)                  # a high-count gene: general path
    X[:, 9] = np.where(rng.rand(n) < 0.1, rng.gamma(2.0, 2.0, n), 0.0)      # continuous non-zeros
    labels = synth.cluster_labels(18, n, 50)
    return X, labels, labels[0]


def _c5_many_groups():
    """BASELINE config 5 structure: 10 000 perturbations + control (10 001 groups of ~12 cells, a control of 5 000 cut
    into ten segments), CSR."""
    rng = np.random.RandomState(19)
    n_perts, per, n_ctrl, N = 10_000, 12, 5_000, 10
    width = len(str(n_perts - 1))
    codes = np.concatenate([np.repeat(np.arange(n_perts), per), np.full(n_ctrl, n_perts)])
    rng.shuffle(codes)
    names = np.array([f"p{i:0{width}d}" for i in range(n_perts)] + [synth.CONTROL])
    labels = names[codes].tolist()
    n = codes.size
    X = (rng.poisson(1.0, size=(n, N)) * (rng.rand(n, N) >= 0.9)).astype(np.float32)
    X[:, 3] = np.where(rng.rand(n) < 0.1, np.round(rng.gamma(2.0, 2.0, n) * 4) / 4, 0.0)   # ~60 distinct values
    X[:, 7] = rng.poisson(15.0, n) * (rng.rand(n) >= 0.5)
    return X, labels, synth.CONTROL


def _grid(fmts, tests, ccs=(True,), tcs=(True,), alts=("two-sided",), log1p=(False,)):
    return list(itertools.product(fmts, tests, ccs, tcs, alts, log1p))


ALL_FMT = ("dense", "csr", "csc")
BOTH = ("ovo", "ovr")

# name -> (builder, list of (fmt, test, use_continuity, tie_correct, alternative, is_log1p), batch_size)
CASES = {
    "conftest": (_conftest, _grid(ALL_FMT, BOTH, (True, False), (True, False), ALTERNATIVES), 16),
    "k562_mini": (_k562_mini, _grid(ALL_FMT, BOTH), 16),
    "k562_mini_cont": (_k562_mini_cont, _grid(ALL_FMT, BOTH), 24),
    "bign": (_bign, _grid(ALL_FMT, BOTH), 4),
    "negatives": (_negatives, _grid(("dense",), BOTH, (True,), (True, False), ALTERNATIVES), 6),
    "edge": (_edge, _grid(ALL_FMT, BOTH, (True, False)), 8),
    "log1p32": (_log1p32, _grid(ALL_FMT, BOTH, log1p=(True,)), 10),
    "log1p64": (_log1p64, _grid(ALL_FMT, BOTH, log1p=(True,)), 10),
    "batched": (_batched, _grid(ALL_FMT, BOTH), 128),
    "highcount": (_highcount, _grid(ALL_FMT, BOTH, alts=("two-sided", "greater")), 16),
    "c4_clusters": (_c4_clusters, _grid(("csc", "dense", "csr"), ("ovr",)) + _grid(("csc",), ("ovo",)), 8),
    "c5_many_groups": (_c5_many_groups, _grid(("csr", "dense"), ("ovo",)) + _grid(("csr",), ("ovr",)), 5),
}


def combo_key(fmt, test, cc, tc, alt, log1p) -> str:
    return f"{fmt}|{test}|cc{int(cc)}|tc{int(tc)}|{alt}|log{int(log1p)}"


def to_format(X, fmt):
    from scipy import sparse

    if fmt == "dense":
        return X
    if fmt == "csr":
        return sparse.csr_matrix(X)
    if fmt == "csc":
        return sparse.csc_matrix(X)
    raise ValueError(fmt)
