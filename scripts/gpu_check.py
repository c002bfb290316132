"""Quick GPU diagnostic: runs every golden case through the CUDA path and prints the worst deviations.
Usage (on the GPU box): python scripts/gpu_check.py [case ...]"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from illico_b200 import asymptotic_wilcoxon  # noqa: E402
from tests.golden import cases as C  # noqa: E402
from tests.util import FakeAnnData, planes  # noqa: E402


def rel(a, b):
    with np.errstate(all="ignore"):
        d = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    d[(a == b)] = 0
    d[np.isnan(a) & np.isnan(b)] = 0
    return d


def main():
    names = sys.argv[1:] or list(C.CASES)
    bad = 0
    for name in names:
        builder, grid, batch_size = C.CASES[name]
        X, labels, reference = builder()
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
        groups = gold["groups"]
        for fmt, test, cc, tc, alt, log1p in grid:
            key = C.combo_key(fmt, test, cc, tc, alt, log1p)
            ref = reference if test == "ovo" else None
            t0 = time.time()
            try:
                df = asymptotic_wilcoxon(FakeAnnData(C.to_format(X, fmt), labels), is_log1p=log1p, group_keys="pert",
                                         reference=ref, batch_size=batch_size, alternative=alt, use_continuity=cc,
                                         tie_correct=tc)
            except Exception as e:
                print(f"[EXC ] {name} {key}: {type(e).__name__}: {e}")
                traceback.print_exc()
                bad += 1
                continue
            dt = time.time() - t0
            p, U, fc = planes(df, len(groups), X.shape[1])
            want = gold[key]
            rows = np.ones(len(groups), bool)
            if ref is not None:
                rows[int(np.searchsorted(groups, ref))] = False
            nU = int((U[rows] != want[1][rows]).sum())
            rp = rel(p[rows], want[0][rows])
            fin = np.isfinite(want[2])
            rf = rel(fc[fin], want[2][fin])
            infbad = int((np.isposinf(fc) != np.isposinf(want[2])).sum())
            ok = nU == 0 and rp.max(initial=0) <= 1e-12 and rf.max(initial=0) <= (1e-6 if log1p else 1e-10) and infbad == 0
            bad += not ok
            print(f"[{'ok  ' if ok else 'FAIL'}] {name} {key}: U mismatches {nU}/{U[rows].size}, p rel max {rp.max(initial=0):.2e}, "
                  f"fc rel max {rf.max(initial=0):.2e}, inf mismatch {infbad}  ({dt*1e3:.0f} ms)")
            if nU:
                idx = np.argwhere(U != want[1])
                for g, j in idx[:5]:
                    print(f"        U[{g},{j}] got {U[g, j]!r} want {want[1][g, j]!r}")
    print("TOTAL FAILURES", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
