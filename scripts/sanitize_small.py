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
print("sanitize run ok")
