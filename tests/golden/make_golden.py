"""Generates the committed golden vectors by running the UNMODIFIED reference here.

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py conftest   # one case

Needs ``/root/reference`` (build container only).  Outputs ``tests/golden/<case>.npz``
(one array ``[3, G, N]`` = (p_value, statistic, fold_change) per flag combination)
and ``tests/golden/primitives.npz`` (known answers of the reference's ranking primitives,
incl. f64 tie sums above 2**53).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden import _ref_harness as H  # noqa: E402
from tests.golden import cases as C  # noqa: E402


def run_case(name: str) -> None:
    builder, grid, batch_size = C.CASES[name]
    X, labels, reference = builder()
    groups = np.unique(np.asarray(labels))
    out = {"groups": groups.astype(str), "reference": np.asarray(reference)}
    t0 = time.time()
    for fmt, test, cc, tc, alt, log1p in grid:
        Xf = C.to_format(X, fmt)
        ref = reference if test == "ovo" else None
        g, p, U, fc = H.ref_run(Xf, labels, ref, is_log1p=log1p, use_continuity=cc, tie_correct=tc,
                                alternative=alt, batch_size=batch_size)
        assert list(g) == list(groups)
        if test == "ovo":
            # dense OVO leaves the reference row uninitialised (reference ovo/dense_ovo.py:116-120);
            # store the sparse kernels' convention (ovo/sparse_ovo.py:140-143) instead of garbage.
            r = int(np.searchsorted(groups, reference))
            p[r, :] = 1.0
            U[r, :] = -1.0
        out[C.combo_key(fmt, test, cc, tc, alt, log1p)] = np.stack([p, U, fc])
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: {len(grid)} combos in {time.time() - t0:.1f}s", flush=True)


def run_primitives() -> None:
    """Known answers of the reference's own primitives (utils/ranking.py:7-158, utils/math.py:64-118)."""
    H.import_reference()
    from illico.utils.math import compute_pval
    from illico.utils.ranking import _accumulate_group_ranksums_from_argsort, rank_sum_and_ties_from_sorted

    out = {}
    # (1) the reference's own unit-test inputs (tests/utils/test_ranking.py:13-56)
    rng = np.random.RandomState(0)
    A = np.sort(rng.randint(0, 10, size=20))
    B = np.sort(rng.randint(0, 10, size=15))
    rs, ts = rank_sum_and_ties_from_sorted(A, B)
    out["t1_A"], out["t1_B"], out["t1_out"] = A, B, np.array([rs, ts])
    rng = np.random.RandomState(0)
    arr = rng.rand(30)
    grp = rng.randint(0, 3, size=30)
    ranks = np.zeros(3)
    ts = _accumulate_group_ranksums_from_argsort(arr, np.argsort(arr), grp, ranks)
    out["t2_arr"], out["t2_grp"], out["t2_ranks"], out["t2_ts"] = arr, grp, ranks, np.array([ts])

    # (2) order-dependent f64 tie sums on the big-n case (column-wise, all cells / pair a-b)
    X, labels, _ = C.CASES["bign"][0]()
    lab = np.asarray(labels)
    codes = np.unique(lab, return_inverse=True)[1].astype(np.int64)
    ovr_ts, ovr_rs, ovo_ts, ovo_rs = [], [], [], []
    for j in range(X.shape[1]):
        col = np.ascontiguousarray(X[:, j])
        ranks = np.zeros(3)
        ovr_ts.append(_accumulate_group_ranksums_from_argsort(col, np.argsort(col, kind="stable"), codes, ranks))
        ovr_rs.append(ranks)
        a = np.sort(col[lab == "a"])
        b = np.sort(col[lab == "b"])
        r, t = rank_sum_and_ties_from_sorted(a, b)
        ovo_rs.append(r)
        ovo_ts.append(t)
    out["bign_ovr_tie"] = np.array(ovr_ts)
    out["bign_ovr_ranksums"] = np.array(ovr_rs)
    out["bign_ovo_tie_ab"] = np.array(ovo_ts)
    out["bign_ovo_ranksum_ab"] = np.array(ovo_rs)

    # (3) compute_pval known answers over a grid, incl. huge z and the tie_corr cut-off
    rows = []
    rng = np.random.RandomState(1)
    for _ in range(400):
        n_ref = int(rng.randint(1, 300_000))
        n_tgt = int(rng.randint(1, 5_000))
        n = n_ref + n_tgt
        tie = float(rng.randint(0, max(1, (n**3 - n) // 2)))
        U = float(rng.randint(0, 2 * n_ref * n_tgt + 1)) / 2.0
        cc = float(rng.randint(0, 2)) * 0.5
        for ai, alt in enumerate(C.ALTERNATIVES):
            p = compute_pval(n_ref, n_tgt, n, tie, U, n_ref * n_tgt / 2.0, cc, alt)
            rows.append((n_ref, n_tgt, tie, U, cc, ai, p))
    for n_ref, n_tgt in ((5, 5), (10_000, 144), (299_000, 1_000)):
        n = n_ref + n_tgt
        for tie in (0.0, float(n**3 - n), float(n**3 - n) * (1 - 5e-10)):
            for U in (0.0, n_ref * n_tgt / 2.0, float(n_ref * n_tgt)):
                for ai, alt in enumerate(C.ALTERNATIVES):
                    p = compute_pval(n_ref, n_tgt, n, tie, U, n_ref * n_tgt / 2.0, 0.5, alt)
                    rows.append((n_ref, n_tgt, tie, U, 0.5, ai, p))
    out["pval_rows"] = np.array(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **out)
    print("primitives done", flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or (["primitives"] + list(C.CASES))
    for nm in names:
        if nm == "primitives":
            run_primitives()
        else:
            run_case(nm)
