"""Single-process multi-GPU check: asymptotic_wilcoxon(..., devices="all") against the single-GPU
call on the K562 shape, dense (pinned and pageable host matrix) and CSR, with timings.  Prints one JSON line per case.

    python scripts/exp/multi_gpu_check.py [--genes 8000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=300_000)
    ap.add_argument("--genes", type=int, default=8_000)
    ap.add_argument("--perts", type=int, default=2_000)
    a = ap.parse_args()
    import pandas as pd
    import torch
    from scipy import sparse

    from illico_b200 import asymptotic_wilcoxon, synth

    n_dev = torch.cuda.device_count()
    rng = np.random.RandomState(0)
    labels, _ = synth.perturbation_labels(rng, a.cells, a.perts)
    Xdev = synth.k562_like_torch(5, a.cells, a.genes, device="cuda:0")
    pin = torch.empty(Xdev.shape, dtype=torch.float32, pin_memory=True)
    pin.copy_(Xdev)
    Xpin = pin.numpy()
    Xpage = np.array(Xpin, copy=True)
    del Xdev
    torch.cuda.empty_cache()

    class Ad:
        pass

    def ad_of(X):
        ad = Ad()
        ad.X, ad.layers = X, {}
        ad.obs = pd.DataFrame({"pert": pd.Categorical(labels)})
        ad.var_names = pd.Index([f"g{i:05d}" for i in range(a.genes)])
        return ad

    def run(X, reps=3, **kw):
        ad = ad_of(X)
        ts, out = [], None
        for i in range(reps + 1):
            for d in range(n_dev):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            out = asymptotic_wilcoxon(ad, is_log1p=False, group_keys="pert", return_array=True, **kw)
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts[1:])), out[2]

    csr = sparse.csr_matrix(Xpage[:, : a.genes])
    for name, X in (("dense_pinned", Xpin), ("dense_pageable", Xpage), ("csr_pageable", csr)):
        for test, ref in (("ovo", synth.CONTROL), ("ovr", None)):
            t1, want = run(X, reference=ref, device="cuda:0")
            want = want.copy()
            rec = {"case": name, "test": test, "n_gpus_present": n_dev, "single_gpu_s": round(t1, 4)}
            for nd in sorted({2, 4, n_dev} & set(range(2, n_dev + 1))):
                tn, got = run(X, reference=ref, devices=nd)
                same = bool(np.array_equal(got[:, :, :2], want[:, :, :2], equal_nan=True)) and \
                    bool(np.allclose(got[:, :, 2], want[:, :, 2], rtol=1e-14, atol=0, equal_nan=True))
                rec[f"devices_{nd}_s"] = round(tn, 4)
                rec[f"devices_{nd}_identical"] = same
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
