"""float64 input through the public call (dense, CSR, CSC; OVO + OVR) -- for the kernel launch list that shows the whole
path runs on this repository's kernels (scripts/exp/float64_launches.sh)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
from illico_b200 import asymptotic_wilcoxon, synth
from tests.golden import cases as C
from tests.parity import assert_parity
from tests.util import FakeAnnData

X, labels = synth.k562_like(seed=9, n_cells=6000, n_genes=48, n_perts=12)
X = np.log1p(X.astype(np.float64) / 3.0)          # float64 values float32 cannot hold
for fmt in ("dense", "csr", "csc"):
    for ref in (synth.CONTROL, None):
        g, p, U, fc = oracle.run(X, labels, ref, is_log1p=True)
        groups, names, out = asymptotic_wilcoxon(FakeAnnData(C.to_format(X, fmt), labels), is_log1p=True, group_keys="pert",
                                                 reference=ref, return_array=True)
        ref_row = int(np.searchsorted(groups, ref)) if ref is not None else None
        assert_parity((out[:, :, 0], out[:, :, 1], out[:, :, 2]), (p, U, fc), ref_row=ref_row, what=f"float64 {fmt} {ref}")
print("float64 run ok")
