"""Phase timings of the CSR rows -> genes repartition on the GPUs of this box (K562 shape)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pandas as pd, torch
from scipy import sparse
from illico_b200 import asymptotic_wilcoxon, synth

n_dev = torch.cuda.device_count()
rng = np.random.RandomState(0)
labels, _ = synth.perturbation_labels(rng, 300_000, 2000)
X = synth.k562_like_torch(5, 300_000, 8000, device="cuda:0").cpu().numpy()
csr = sparse.csr_matrix(X)
del X
class Ad: pass
ad = Ad(); ad.X, ad.layers = csr, {}
ad.obs = pd.DataFrame({"pert": pd.Categorical(labels)}); ad.var_names = pd.Index([f"g{i}" for i in range(8000)])
for nd in sorted({1, 2, n_dev}):
    for rep in range(4):
        os.environ["ILLICO_REPART_TIMING"] = "1" if rep == 3 else "0"
        for d in range(n_dev): torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        asymptotic_wilcoxon(ad, is_log1p=False, group_keys="pert", reference=synth.CONTROL, return_array=True, devices=nd)
        print(f"devices={nd} rep {rep}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
