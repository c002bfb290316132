"""Where the end-to-end time goes: phases of asymptotic_wilcoxon on pinned host input (K562 shape)."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from illico_b200 import synth, asymptotic_wilcoxon
from illico_b200.groups import encode_and_count_groups
from illico_b200.engine import Engine, make_flags, upload_dense, upload_sparse
from scipy import sparse
import pandas as pd

fmt = sys.argv[1] if len(sys.argv) > 1 else "csr"
n, N, P = 300_000, 8_000, 2_000
dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)
labels, _ = synth.perturbation_labels(rng, n, P)
Xdev = synth.k562_like_torch(1, n, N, device=dev)
def pin(t):
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True); h.copy_(t); return h
if fmt == "dense":
    Xh = pin(Xdev).numpy()
else:
    sp = Xdev.to_sparse_csr()
    d, i, p = pin(sp.values()), pin(sp.col_indices().to(torch.int32)), pin(sp.crow_indices().to(torch.int32))
    Xh = sparse.csr_matrix((d.numpy(), i.numpy(), p.numpy()), shape=(n, N))
del Xdev
torch.cuda.synchronize()
obs = pd.DataFrame({"pert": labels})
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T()
    M = upload_dense(Xh, dev) if fmt == "dense" else upload_sparse(Xh, "csr", dev)
    t1 = T()
    uniq, grpc = encode_and_count_groups(obs["pert"], synth.CONTROL)
    t2 = T()
    eng = Engine(grpc, dev)
    t3 = T()
    flags = make_flags(False, True, True, "two-sided", fmt)
    res = torch.empty((eng.n_groups, N, 3), dtype=torch.float64, device=dev)
    t4 = T()
    if fmt == "csr":
        eng.check_csr_sorted(M)
    t5 = T()
    eng.run_batch(M, 0, N, flags, res, 0)
    t6 = T()
    host = torch.empty(res.shape, dtype=torch.float64, pin_memory=True)
    t7 = T()
    host.copy_(res, non_blocking=True)
    t8 = T()
    print(f"[{fmt} it{it}] upload {1e3*(t1-t0):.1f} ms | encode {1e3*(t2-t1):.1f} | engine/plan {1e3*(t3-t2):.1f} | alloc res {1e3*(t4-t3):.1f} | "
          f"sorted check {1e3*(t5-t4):.1f} | kernels {1e3*(t6-t5):.1f} | pinned alloc {1e3*(t7-t6):.1f} | d2h {1e3*(t8-t7):.1f} | total {1e3*(t8-t0):.1f}")
    del M, res, host, eng
class Ad: pass
ad = Ad(); ad.X, ad.layers, ad.obs, ad.var_names = Xh, {}, obs, pd.Index([f"g{i}" for i in range(N)])
for it in range(3):
    t0 = T(); out = asymptotic_wilcoxon(ad, is_log1p=False, group_keys="pert", reference=synth.CONTROL, return_array=True); t1 = T()
    print(f"[{fmt}] public call {1e3*(t1-t0):.1f} ms")
