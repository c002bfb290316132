"""Pins the CPU oracle (oracle/wilcoxon_oracle.c) to the reference.

Golden vectors come from the unmodified reference run in the build container
(tests/golden/make_golden.py); the scipy cross-check restates the reference's own
end-to-end test (reference tests/test_asymptotic_wilcoxon.py:111-194).
"""
import os

import numpy as np
import pytest

import oracle
from tests.golden import cases as C
from tests.parity import FC_RTOL, FC_RTOL_LOG1P_F32, assert_parity


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"{name}.npz"))


@pytest.mark.parametrize("name", list(C.CASES))
def test_oracle_matches_reference_golden(golden_dir, name):
    builder, grid, batch_size = C.CASES[name]
    X, labels, reference = builder()
    gold = _load(golden_dir, name)
    groups = gold["groups"]
    for fmt, test, cc, tc, alt, log1p in grid:
        Xf = C.to_format(X, fmt)
        ref = reference if test == "ovo" else None
        g, p, U, fc = oracle.run(Xf, labels, ref, is_log1p=log1p, use_continuity=cc, tie_correct=tc,
                                 alternative=alt, batch_size=batch_size, n_threads=2)
        assert list(g) == list(groups)
        want = gold[C.combo_key(fmt, test, cc, tc, alt, log1p)]
        ref_row = int(np.searchsorted(groups, reference)) if test == "ovo" else None
        fc_rtol = FC_RTOL_LOG1P_F32 if (log1p and X.dtype == np.float32) else FC_RTOL
        assert_parity((p, U, fc), (want[0], want[1], want[2]), ref_row=ref_row, fc_rtol=fc_rtol,
                      what=f"{name}:{fmt}:{test}:cc{cc}:tc{tc}:{alt}")


def test_primitives_known_answers(golden_dir):
    """reference tests/utils/test_ranking.py:13-56 restated + the reference's own outputs."""
    from scipy.stats import rankdata

    g = _load(golden_dir, "primitives")
    A, B = g["t1_A"], g["t1_B"]
    rs, ts = oracle.rank_sum_and_ties_from_sorted(A, B)
    assert (rs, ts) == tuple(g["t1_out"])
    comb = np.concatenate([A, B])
    assert rs == rankdata(comb)[len(A):].sum()
    _, cnt = np.unique(comb, return_counts=True)
    assert ts == float((cnt**3 - cnt).sum())

    arr, grp = g["t2_arr"], g["t2_grp"]
    ranks, ts = oracle.accumulate_group_ranksums_from_argsort(arr, np.argsort(arr), grp, 3)
    np.testing.assert_array_equal(ranks, g["t2_ranks"])
    assert ts == g["t2_ts"][0]
    manual = np.zeros(3)
    np.add.at(manual, grp, rankdata(arr))
    np.testing.assert_array_equal(ranks, manual)


def test_tie_sums_above_2_53_bit_exact(golden_dir):
    """f64 tie sums are order dependent above 2**53 (n = 300k); the oracle must reproduce the
    reference's sequential accumulation bit for bit (hard part 1)."""
    g = _load(golden_dir, "primitives")
    X, labels, _ = C.CASES["bign"][0]()
    lab = np.asarray(labels)
    codes = np.unique(lab, return_inverse=True)[1]
    exact_differs = False
    for j in range(X.shape[1]):
        col = X[:, j].astype(np.float64)
        ranks, ts = oracle.accumulate_group_ranksums_from_argsort(col, np.argsort(col, kind="stable"), codes, 3)
        assert ts == g["bign_ovr_tie"][j]
        np.testing.assert_array_equal(ranks, g["bign_ovr_ranksums"][j])
        _, cnt = np.unique(col, return_counts=True)
        exact = int((cnt.astype(object) ** 3 - cnt.astype(object)).sum())
        exact_differs |= float(exact) != ts
        rs, t2 = oracle.rank_sum_and_ties_from_sorted(np.sort(col[lab == "a"]), np.sort(col[lab == "b"]))
        assert rs == g["bign_ovo_ranksum_ab"][j] and t2 == g["bign_ovo_tie_ab"][j]
    assert exact_differs, "the case is meant to exercise order-dependent rounding"
    # full path: tie sums reported by the batch kernels, dense (in sorted position) vs sparse (zero block last)
    _, _, _, _, ties_dense = oracle.run(X, labels, None, want_ties=True)
    np.testing.assert_array_equal(ties_dense, g["bign_ovr_tie"])


def test_compute_pval_known_answers(golden_dir):
    rows = _load(golden_dir, "primitives")["pval_rows"]
    for n_ref, n_tgt, tie, U, cc, ai, want in rows:
        n_ref, n_tgt = int(n_ref), int(n_tgt)
        got = oracle.compute_pval(n_ref, n_tgt, n_ref + n_tgt, tie, U, n_ref * n_tgt / 2.0, cc, C.ALTERNATIVES[int(ai)])
        assert got == pytest.approx(want, rel=1e-13, abs=2.3e-308)


@pytest.mark.parametrize("alternative", C.ALTERNATIVES)
@pytest.mark.parametrize("use_continuity", [True, False])
@pytest.mark.parametrize("test", ["ovo", "ovr"])
@pytest.mark.parametrize("fmt", ["dense", "csc", "csr"])
def test_oracle_vs_scipy(fmt, test, use_continuity, alternative):
    """reference tests/test_asymptotic_wilcoxon.py:111-185 restated: U exact, p 1e-12, fc 1e-6 vs scipy.
    scipy >= 1.17 computes in float32 for float32 input, so the oracle's inputs are cast to f64 for scipy."""
    from scipy.stats import mannwhitneyu

    X, labels, reference = C.CASES["conftest"][0]()
    ref = reference if test == "ovo" else None
    groups, p, U, fc = oracle.run(C.to_format(X, fmt), labels, ref, use_continuity=use_continuity,
                                  alternative=alternative, batch_size=16)
    lab = np.asarray(labels)
    X64 = X.astype(np.float64)
    for gi, gname in enumerate(groups):
        if gname == ref:
            continue
        mask = lab == gname
        refX = X64[lab == ref] if ref is not None else X64[~mask]
        stats, pv = mannwhitneyu(refX, X64[mask], axis=0, method="asymptotic", use_continuity=use_continuity,
                                 alternative=alternative)
        np.testing.assert_array_equal(U[gi], stats)
        np.testing.assert_allclose(p[gi], pv, rtol=1e-12, atol=0)
        np.testing.assert_allclose(fc[gi], X64[mask].mean(0) / refX.mean(0), rtol=1e-6)


def test_empty_and_ragged_inputs():
    # zero genes in range, a group without any non-zero, CSR with empty rows
    from scipy import sparse

    rng = np.random.RandomState(2)
    X = (rng.rand(200, 5) < 0.1).astype(np.float32) * rng.randint(1, 4, size=(200, 5)).astype(np.float32)
    X[:50] = 0
    labels = ["a"] * 50 + ["b"] * 100 + ["c"] * 50
    outs = [oracle.run(f(X), labels, r) for f in (lambda a: a, sparse.csr_matrix, sparse.csc_matrix) for r in (None, "b")]
    for k in (0, 1):
        for o in outs[k::2][1:]:
            np.testing.assert_array_equal(o[2], outs[k][2])
            np.testing.assert_allclose(o[1], outs[k][1], rtol=1e-14)
    P = oracle.Prepared(X, labels, None)
    res = oracle.run_prepared(P, gene_lb=2, gene_ub=2)
    assert np.isnan(res).all()
