"""CPU oracle: ctypes front end of ``oracle/wilcoxon_oracle.c``.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Importers allowed: ``tests/``,
``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs.
``illico_b200`` never imports this package.

Parity status: pinned against golden vectors generated from the unmodified reference
(``tests/golden/make_golden.py``) and against ``scipy.stats.mannwhitneyu``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

ALTERNATIVES = {"two-sided": 0, "less": 1, "greater": 2}
FMT_DENSE, FMT_CSC, FMT_CSR = 0, 1, 2


def build(force: bool = False) -> str:
    """Compiles the C restatement with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "wilcoxon_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i64, f64, vp, ci = C.c_int64, C.c_double, C.c_void_p, C.c_int
        L.oracle_accumulate_group_ranksums_from_argsort.restype = f64
        L.oracle_accumulate_group_ranksums_from_argsort.argtypes = [vp, vp, vp, i64, vp]
        L.oracle_rank_sum_and_ties_from_sorted.restype = None
        L.oracle_rank_sum_and_ties_from_sorted.argtypes = [vp, i64, vp, i64, vp, vp]
        L.oracle_compute_pval.restype = f64
        L.oracle_compute_pval.argtypes = [i64, i64, i64, f64, f64, f64, f64, ci]
        L.oracle_asymptotic_wilcoxon.restype = ci
        L.oracle_asymptotic_wilcoxon.argtypes = [ci, ci, vp, vp, vp, i64, i64, i64, i64, i64,
                                                 i64, vp, vp, vp, vp, i64,
                                                 ci, ci, ci, ci, i64, ci, vp, vp]
        L.oracle_check_indices_sorted_per_parcel.restype = ci
        L.oracle_check_indices_sorted_per_parcel.argtypes = [vp, vp, i64]
        L.oracle_max_threads.restype = ci
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---- primitives ---------------------------------------------------------------------------

def accumulate_group_ranksums_from_argsort(arr, idx, groups, n_groups):
    """reference illico/utils/ranking.py:7-49 -> (ranksums[G], tie_sum)"""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    groups = np.ascontiguousarray(groups, dtype=np.int64)
    ranks = np.zeros(n_groups, dtype=np.float64)
    ts = lib().oracle_accumulate_group_ranksums_from_argsort(_p(arr), _p(idx), _p(groups), arr.size, _p(ranks))
    return ranks, ts


def rank_sum_and_ties_from_sorted(A, B):
    """reference illico/utils/ranking.py:52-158 -> (rank sum of B, tie sum)"""
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    out = np.zeros(2, dtype=np.float64)
    lib().oracle_rank_sum_and_ties_from_sorted(_p(A), A.size, _p(B), B.size, _p(out[0:1]), _p(out[1:2]))
    return float(out[0]), float(out[1])


def compute_pval(n_ref, n_tgt, n, tie_sum, U, mu, contin_corr, alternative):
    """reference illico/utils/math.py:64-118"""
    return lib().oracle_compute_pval(int(n_ref), int(n_tgt), int(n), float(tie_sum), float(U), float(mu),
                                     float(contin_corr), ALTERNATIVES[alternative])


# ---- group encoding (reference illico/utils/groups.py:18-58) ---------------------------------

def encode_groups(labels, reference):
    labels = np.asarray(list(labels))
    if reference is not None and reference not in labels:
        raise ValueError(f"Reference group `{reference}` is not present in the group labels.")
    uniq, inv, counts = np.unique(labels, return_inverse=True, return_counts=True)
    inv = inv.astype(np.int64)
    indices = np.argsort(inv, kind="stable").astype(np.int64)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ref = -1 if reference is None else int(np.searchsorted(uniq, reference))
    return uniq, inv, counts.astype(np.int64), indices, indptr, ref


# ---- full path -----------------------------------------------------------------------------

class Prepared:
    """Inputs converted once to the layouts the C code reads (excluded from baseline timing)."""

    def __init__(self, X, labels, reference):
        from scipy import sparse

        self.uniq, self.enc, self.counts, self.gidx, self.gptr, self.ref = encode_groups(labels, reference)
        self.n_rows, self.n_cols = X.shape
        self.indices = self.indptr = None
        if isinstance(X, np.ndarray):
            self.fmt = FMT_DENSE
            if X.dtype not in (np.float32, np.float64):
                X = X.astype(np.float64)
            self.data = np.ascontiguousarray(X)
            self.ld = self.data.shape[1]
        else:
            if sparse.isspmatrix_csr(X) or isinstance(X, sparse.csr_array):
                self.fmt = FMT_CSR
            elif sparse.isspmatrix_csc(X) or isinstance(X, sparse.csc_array):
                self.fmt = FMT_CSC
            else:
                raise KeyError(f"Support for data type {type(X)} is not implemented.")
            d = X.data
            if d.dtype not in (np.float32, np.float64):
                d = d.astype(np.float64)
            self.data = np.ascontiguousarray(d)
            self.indices = np.ascontiguousarray(X.indices, dtype=np.int32)
            self.indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
            self.ld = 0
        self.dtype = 0 if self.data.dtype == np.float32 else 1


def run_prepared(P: Prepared, *, is_log1p=False, use_continuity=True, tie_correct=True, alternative="two-sided",
                 batch_size=256, n_threads=1, gene_lb=0, gene_ub=None, want_ties=False):
    """Runs the oracle on genes [gene_lb, gene_ub); returns ``results[G, N, 3]`` (+ tie sums)."""
    G = P.counts.size
    if gene_ub is None:
        gene_ub = P.n_cols
    results = np.full((G, P.n_cols, 3), np.nan, dtype=np.float64)
    ties = None
    if want_ties:
        ties = np.full((G, P.n_cols) if P.ref >= 0 else (P.n_cols,), np.nan, dtype=np.float64)
    rc = lib().oracle_asymptotic_wilcoxon(
        P.fmt, P.dtype, _p(P.data), _p(P.indices), _p(P.indptr), P.n_rows, P.n_cols, P.ld, gene_lb, gene_ub,
        G, _p(P.enc), _p(P.counts), _p(P.gidx), _p(P.gptr), P.ref,
        int(is_log1p), int(use_continuity), int(tie_correct), ALTERNATIVES[alternative],
        int(batch_size), int(n_threads), _p(results), _p(ties))
    if rc != 0:
        raise RuntimeError(f"oracle failed with status {rc}")
    return (results, ties) if want_ties else results


def run(X, labels, reference=None, **kw):
    """Convenience: ``(groups, p[G,N], U[G,N], fc[G,N])`` like the reference's three output planes."""
    P = Prepared(X, labels, reference)
    want_ties = kw.pop("want_ties", False)
    out = run_prepared(P, want_ties=want_ties, **kw)
    res, ties = out if want_ties else (out, None)
    ret = (P.uniq, res[:, :, 0], res[:, :, 1], res[:, :, 2])
    return ret + (ties,) if want_ties else ret


def max_threads() -> int:
    return int(lib().oracle_max_threads())
