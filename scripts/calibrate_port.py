"""Calibration of bench.py's CPU arm: the C port (oracle/wilcoxon_oracle.c, `kind: "port"`)
against the UNMODIFIED numba reference imported from /root/reference, on the same seeded K562-shape sample, same thread
count, same batch size (256, the reference's own benchmark setting, tests/test_asymptotic_wilcoxon.py:302), JIT
excluded by a prior tiny call.  Runs in the build container only (the reference does not travel to the GPU box);
writes profiles/r2_port_vs_numba.json.

    python scripts/calibrate_port.py [--genes 256] [--cells 300000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=300_000)
    ap.add_argument("--genes", type=int, default=256)
    ap.add_argument("--perts", type=int, default=2_000)
    a = ap.parse_args()
    import _ref_harness as H
    from scipy import sparse

    import oracle
    from illico_b200 import synth

    threads = os.cpu_count()
    X, labels = synth.k562_like(seed=5, n_cells=a.cells, n_genes=a.genes, n_perts=a.perts)
    G = len(set(labels))
    out = {"cells": a.cells, "genes": a.genes, "groups": G, "threads": threads, "batch_size": 256,
           "host": "build container", "cases": {}}
    tiny = X[:2000, :8].copy()
    tl = labels[:2000]
    if synth.CONTROL not in tl:
        tl = list(tl)
        tl[0] = synth.CONTROL
    for fmt in ("dense", "csr"):
        Xf = X if fmt == "dense" else sparse.csr_matrix(X)
        tf = tiny if fmt == "dense" else sparse.csr_matrix(tiny)
        for test, ref in (("ovo", synth.CONTROL), ("ovr", None)):
            H.ref_run(tf, tl, ref, batch_size=8, n_threads=threads)            # JIT warm-up, excluded
            t0 = time.perf_counter()
            g, p, U, fc = H.ref_run(Xf, labels, ref, batch_size=256, n_threads=threads)
            t_ref = time.perf_counter() - t0
            P = oracle.Prepared(Xf, labels, ref)
            oracle.run_prepared(P, batch_size=8, n_threads=threads, gene_lb=0, gene_ub=8)
            t0 = time.perf_counter()
            res = oracle.run_prepared(P, batch_size=256, n_threads=threads, gene_lb=0, gene_ub=a.genes)
            t_port = time.perf_counter() - t0
            out["cases"][f"{fmt}_{test}"] = {
                "numba_reference_s": round(t_ref, 3), "c_port_s": round(t_port, 3),
                "numba_reference_tests_per_s": round(G * a.genes / t_ref, 1), "c_port_tests_per_s": round(G * a.genes / t_port, 1),
                "port_over_reference": round(t_ref / t_port, 3)}
            print(fmt, test, out["cases"][f"{fmt}_{test}"], flush=True)
    out["reading"] = ("port_over_reference > 1 means the C port is FASTER than the numba reference on the same cores, i.e. a "
                      "GPU / CPU-arm ratio computed against the port understates the speed-up over the real reference. "
                      "Note that batch_size=256 with 256 genes gives ONE batch, i.e. one busy thread in both "
                      "implementations when genes <= 256; use --genes >= 256 * threads for the threaded figure.")
    with open(os.path.join(ROOT, "profiles", "r2_port_vs_numba.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
