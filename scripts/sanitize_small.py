"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): all formats, OVO + OVR, all tiers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from illico_b200 import asymptotic_wilcoxon  # noqa: E402
from tests.golden import cases as C  # noqa: E402
from tests.util import FakeAnnData  # noqa: E402

for name in ("edge", "negatives", "k562_mini_cont"):
    X, labels, reference = C.CASES[name][0]()
    X = X[:2500, :6].copy()
    labels = labels[:2500]
    for fmt in ("dense", "csr", "csc"):
        if name == "negatives" and fmt != "dense":
            continue
        for ref in (None, labels[0]):
            out = asymptotic_wilcoxon(FakeAnnData(C.to_format(X, fmt), labels), is_log1p=False, group_keys="pert",
                                      reference=ref, return_array=True)
            assert np.isfinite(out[2][:, :, 1]).all()
# count data: the fused single-pass kernels (TMA ring + mbarriers, shared-memory histograms), with genes handed back
from illico_b200 import synth  # noqa: E402

os.environ["ILLICO_FUSED_LIST_SHARE"] = "1.0"
X, labels = synth.k562_like(seed=3, n_cells=3000, n_genes=24, n_perts=9)
X[:, 5] = np.random.RandomState(0).poisson(30.0, X.shape[0])   # more than 12 distinct values: handed back
for fmt in ("dense", "csr"):
    for ref in (None, synth.CONTROL):
        for log1p in (False, True):
            Xi = np.log1p(X).astype(np.float32) if log1p else X
            out = asymptotic_wilcoxon(FakeAnnData(C.to_format(Xi, fmt), labels), is_log1p=log1p, group_keys="pert",
                                      reference=ref, return_array=True)
            assert np.isfinite(out[2][:, :, 1]).all()
# round 2: the wide-table pass (integer counts up to 64: dense high-count genes, a value past 64, a fractional one), the
# whole-batch hand-back decided on the device, and the CSR rows -> genes repartition kernels
os.environ["ILLICO_FUSED_LIST_SHARE"] = "0.125"
rng = np.random.RandomState(1)
X, labels = synth.k562_like(seed=4, n_cells=3000, n_genes=40, n_perts=9)
X[:, 3] = rng.poisson(12.0, X.shape[0])
X[:, 4] = rng.poisson(30.0, X.shape[0])
X[:, 9] = rng.poisson(50.0, X.shape[0])
X[7, 11] = 2.5
out = asymptotic_wilcoxon(FakeAnnData(X, labels), is_log1p=False, group_keys="pert", reference=synth.CONTROL, return_array=True)
assert np.isfinite(out[2][:, :, 1]).all()
Xc = np.log1p(X / (X.sum(1, keepdims=True) + 1.0) * 1e4).astype(np.float32)      # continuous: every gene handed back (HB_ALL)
for fmt in ("dense", "csr"):
    out = asymptotic_wilcoxon(FakeAnnData(C.to_format(Xc, fmt), labels), is_log1p=False, group_keys="pert",
                              reference=synth.CONTROL, return_array=True)
    assert np.isfinite(out[2][:, :, 1]).all()
import torch  # noqa: E402

from illico_b200 import repartition  # noqa: E402

csr = C.to_format(X, "csr")
shards = repartition.repartition_csr(csr, [torch.device("cuda", 0)] * 3, [0, 13, 27, 40])
torch.cuda.synchronize()
assert sum(int(m.data.numel()) for m in shards) == csr.nnz
print("sanitize run ok")
