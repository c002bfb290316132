"""Packed upload: first-call and steady-state time of the public call against the staging chunk size
(ILLICO_STAGE_CHUNK_MB), pinned and pageable source, K562 shape."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pandas as pd, torch
from illico_b200 import asymptotic_wilcoxon, synth, hostio

rng = np.random.RandomState(0)
labels, _ = synth.perturbation_labels(rng, 300_000, 2000)
Xdev = synth.k562_like_torch(5, 300_000, 8000, device="cuda:0")
pin = torch.empty(Xdev.shape, dtype=torch.float32, pin_memory=True); pin.copy_(Xdev); del Xdev
Xs = {"pinned": pin.numpy()}
if os.environ.get("PAGEABLE", "1") == "1":
    Xs["pageable"] = np.array(pin.numpy(), copy=True)
class Ad: pass
for name, X in Xs.items():
    ad = Ad(); ad.X, ad.layers = X, {}
    ad.obs = pd.DataFrame({"pert": pd.Categorical(labels)}); ad.var_names = pd.Index([f"g{i}" for i in range(8000)])
    ts = []
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        asymptotic_wilcoxon(ad, is_log1p=False, group_keys="pert", reference=synth.CONTROL, return_array=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"chunk {hostio.CHUNK_BYTES >> 20} MB, {name}: first {ts[0]:.3f} s, then {np.median(ts[2:]):.4f} s, upload {dict(hostio.LAST_UPLOAD)}", flush=True)
