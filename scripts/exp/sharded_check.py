"""2-GPU check of asymptotic_wilcoxon_sharded (run under torchrun): same answer as the single-GPU call."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from illico_b200 import asymptotic_wilcoxon, synth
from illico_b200.parallel import asymptotic_wilcoxon_sharded
from tests.util import FakeAnnData
from scipy import sparse

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, labels = synth.k562_like(seed=5, n_cells=6000, n_genes=203, n_perts=20)
for fmt in ("dense", "csr"):
    Xf = X if fmt == "dense" else sparse.csr_matrix(X)
    ad = FakeAnnData(Xf, labels)
    for ref in (synth.CONTROL, None):
        want = asymptotic_wilcoxon(ad, is_log1p=False, group_keys="pert", reference=ref)
        got = asymptotic_wilcoxon_sharded(ad, is_log1p=False, group_keys="pert", reference=ref)
        np.testing.assert_array_equal(got.to_numpy(), want.to_numpy())
        assert got.index.equals(want.index)
dist.barrier()
if dist.get_rank() == 0:
    print("sharded ok", dist.get_world_size())
dist.destroy_process_group()
