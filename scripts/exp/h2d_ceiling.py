"""Aggregate host-to-device bandwidth of this box: k GPUs copying from their own pinned buffers at the same time (the
ceiling the end-to-end figures of `bench.py --gpus N` are measured against).  One JSON line per k."""
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from illico_b200 import hostio

n_dev = torch.cuda.device_count()
GB = 2
def run(k, bind):
    bufs, devs, ready = [None] * k, [None] * k, threading.Barrier(k + 1)
    times = [0.0] * k
    def worker(i):
        if bind:
            hostio.bind_thread_to_device_node(i)
        torch.cuda.set_device(i)
        h = torch.empty(GB << 28, dtype=torch.float32, pin_memory=True)   # GB GiB
        h.zero_()
        d = torch.empty_like(h, device=f"cuda:{i}")
        d.copy_(h, non_blocking=True); torch.cuda.synchronize(i)
        ready.wait()            # all buffers exist
        ready.wait()            # go
        t0 = time.perf_counter()
        for _ in range(4):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize(i)
        times[i] = time.perf_counter() - t0
        ready.wait()
    ths = [threading.Thread(target=worker, args=(i,)) for i in range(k)]
    for t in ths: t.start()
    ready.wait(); ready.wait(); ready.wait()
    for t in ths: t.join()
    tot = 4 * GB * 1.073741824 * k
    return {"gpus": k, "numa_bound": bind, "aggregate_GBps": round(tot / max(times), 1), "per_gpu_GBps": round(tot / max(times) / k, 1)}
for k in sorted({1, 2, 4, n_dev} & set(range(1, n_dev + 1))):
    for bind in (True, False):
        print(json.dumps(run(k, bind)), flush=True)
