"""bench.py -- K562-shape asymptotic Wilcoxon rank-sum throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dense_ovo|dense_ovr|csr_ovo|csr_ovr]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU algorithm (oracle port) on the host cores

A "step" is one pass of the hot path over the whole workload (all genes x all groups).  `value` is
gene x group tests per second with the input resident in HBM; `e2e` is the same metric through the public
`asymptotic_wilcoxon` call with HOST (pinned) input, H2D and D2H inside the timed region.  Genes shard
across ranks with no data-path collective ("weak": every rank ranks its own K562-shape gene shard).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (format, test)
    "dense_ovo": ("dense", "ovo"),
    "dense_ovr": ("dense", "ovr"),
    "csr_ovo": ("csr", "ovo"),
    "csr_ovr": ("csr", "ovr"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dense_ovo", choices=list(WORKLOADS))
    ap.add_argument("--cells", type=int, default=300_000)
    ap.add_argument("--genes", type=int, default=8_000)
    ap.add_argument("--perts", type=int, default=2_000)
    ap.add_argument("--cpu-sample-genes", type=int, default=0, help="genes of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--sorted-cells", action="store_true", help="experiment: cells already in group order")
    ap.add_argument("--high-count-frac", type=float, default=0.0,
                    help="experiment: this fraction of the genes (scattered) gets Poisson(30) counts, i.e. more distinct "
                         "values than the fused path's 12-slot table holds (they are handed back to the general path)")
    ap.add_argument("--continuous", action="store_true",
                    help="stress variant (SURVEY 8d): log1p of library-size-normalised counts, almost no ties among the "
                         "non-zeros; takes the general stage + rank path")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows: list = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_labels(seed, n_cells, n_perts, test):
    from illico_b200 import synth

    rng = np.random.RandomState(seed)
    labels, _ = synth.perturbation_labels(rng, n_cells, n_perts)
    return labels, (synth.CONTROL if test == "ovo" else None)


class Ad:
    def __init__(self, X, labels, n_genes):
        import pandas as pd

        self.X, self.layers = X, {}
        self.obs = pd.DataFrame({"pert": labels})
        self.var_names = pd.Index([f"g{i:05d}" for i in range(n_genes)])


# ---------------------------------------------------------------------------------------------------------
def cpu_run(fmt, test, Xhost, labels, reference, n_genes_total, sample, threads):
    """Times the oracle port (the reference's algorithm, oracle/wilcoxon_oracle.c) on `sample` genes."""
    import oracle

    P = oracle.Prepared(Xhost, labels, reference)
    bs = max(1, -(-sample // max(threads, 1)))
    bs = min(bs, 256)  # the reference's own benchmark batch size is 256 (tests/test_asymptotic_wilcoxon.py:302)
    t0 = time.perf_counter()
    oracle.run_prepared(P, batch_size=bs, n_threads=threads, gene_lb=0, gene_ub=sample)
    dt = time.perf_counter() - t0
    G = P.counts.size
    return G * sample / dt, dt


def host_sample(Xdev, fmt, sample):
    """Host copy of the first `sample` genes of the device matrix (dense ndarray or scipy CSR)."""
    import torch
    from scipy import sparse

    sub = Xdev[:, :sample].contiguous().cpu().numpy()
    return sub if fmt == "dense" else sparse.csr_matrix(sub)


def main():
    a = parse()
    fmt, test = WORKLOADS[a.workload]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    n_tests_rank = None

    if a.impl == "reference":
        return reference_arm(a, fmt, test, rank, world)

    import torch
    import torch.distributed as dist

    from illico_b200 import _lib, asymptotic_wilcoxon, synth
    from illico_b200.engine import Engine, make_flags
    from illico_b200.groups import encode_and_count_groups

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic K562-shape shard of this rank (weak scaling: every rank owns a full-size gene shard)
    labels, reference = make_labels(a.seed, a.cells, a.perts, test)
    if a.sorted_cells:
        labels = sorted(labels)
    Xdev = synth.k562_like_torch(a.seed + 1000 * rank, a.cells, a.genes, device=dev)
    if a.high_count_frac > 0:
        gsel = torch.randperm(a.genes, device=dev, generator=torch.Generator(device=dev).manual_seed(7))[: max(1, int(a.genes * a.high_count_frac))]
        for r0 in range(0, a.cells, 16384):
            blk = Xdev[r0:r0 + 16384]
            blk[:, gsel] = torch.poisson(torch.full((blk.shape[0], gsel.numel()), 30.0, device=dev))
    if a.continuous:
        for r0 in range(0, a.cells, 16384):   # in place, chunked: log1p(x / library size * 1e4)
            blk = Xdev[r0:r0 + 16384]
            lib = blk.sum(dim=1, keepdim=True) + 1.0
            blk.copy_(torch.log1p(blk / lib * 1.0e4))
    uniq, grpc = encode_and_count_groups(labels, reference)
    G = grpc.counts.size
    n_tests_rank = G * a.genes

    eng = Engine(grpc, dev)
    flags = make_flags(False, True, True, "two-sided", fmt)
    if fmt == "dense":
        from illico_b200.engine import DeviceMatrix

        M = DeviceMatrix("dense", (a.cells, a.genes), Xdev)
        host_obj = None
    else:
        from scipy import sparse

        sp = Xdev.to_sparse_csr()
        from illico_b200.engine import DeviceMatrix

        M = DeviceMatrix("csr", (a.cells, a.genes), sp.values().contiguous(), sp.col_indices().to(torch.int32),
                         sp.crow_indices().to(torch.int64))
    results = torch.empty((G, a.genes, 3), dtype=torch.float64, device=dev)
    bmax = eng.max_batch_genes(a.genes)
    bounds = list(range(0, a.genes, bmax)) + [a.genes]
    batches = list(zip(bounds[:-1], bounds[1:]))

    def step():
        for lb, ub in batches:
            eng.run_batch(M, lb, ub, flags, results, lb)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # A step takes 2-4 ms, far below nvidia-smi's sampling period, and the SM clock needs tens of milliseconds of load
    # to reach its boost state.  So the sampler starts first, the warm-up runs the requested steps and then keeps the
    # same load up for at least 0.6 s (untimed), the K timed steps follow at once, and the load continues for 0.3 s
    # after them: every clock sample is taken under the step's load, with the timed region in the middle.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_load = time.perf_counter()
    n_warm = 0
    while n_warm < max(a.warmup, 3) or time.perf_counter() - t_load < 0.6:
        step()
        n_warm += 1
        if n_warm % 8 == 0:
            torch.cuda.synchronize(dev)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    launches_mark = _lib.launch_count()
    t_post = time.perf_counter()
    while time.perf_counter() - t_post < 0.3:
        step()
    barrier()
    launches = launches_mark - l0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = f"{n_warm} warm-up steps (>= 0.6 s of the same load) + the {a.steps} timed steps + 0.3 s of the same load"

    # ---- per-kernel timing of one step (stage vs rank), for the roofline of the dominant kernel
    lib = eng.lib
    import ctypes as C

    def time_kernels():
        ts, tr = 0.0, 0.0
        st = torch.cuda.current_stream(dev).cuda_stream
        for lb, ub in batches:
            b = ub - lb
            eng._ensure_buffers(b)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            if fmt == "dense":
                rc = lib.illico_stage_dense_f32(M.data.data_ptr(), M.ld, lb, b, C.byref(eng.plan), eng._ir_vals.data_ptr(),
                                                eng._ir_cnt.data_ptr(), st)
            else:
                rc = lib.illico_stage_csr_f32(M.data.data_ptr(), M.indices.data_ptr(), M.indptr.data_ptr(), lb, b,
                                              C.byref(eng.plan), eng._ir_vals.data_ptr(), eng._ir_cnt.data_ptr(),
                                              eng._ws.data_ptr(), eng._ws.numel(), st)
            _lib.check(rc, "stage")
            ev[1].record()
            fn = lib.illico_rank_ovo if test == "ovo" else lib.illico_rank_ovr
            rc = fn(eng._ir_vals.data_ptr(), eng._ir_cnt.data_ptr(), b, C.byref(eng.plan), C.byref(flags),
                    results.data_ptr() + lb * 24, a.genes * 3, eng._ws.data_ptr(), eng._ws.numel(), None, st)
            _lib.check(rc, "rank")
            ev[2].record()
            torch.cuda.synchronize(dev)
            ts += ev[0].elapsed_time(ev[1])
            tr += ev[1].elapsed_time(ev[2])
        return ts, tr

    time_kernels()
    ks = [time_kernels() for _ in range(3)]
    t_stage = float(np.median([k[0] for k in ks]))
    t_rank = float(np.median([k[1] for k in ks]))
    nnz = int((Xdev != 0).sum().item())
    peak, peak_src = load_peaks()
    # algorithmic bytes: stage = every input element read once; rank = staged non-zeros + counts
    # read once, results written once.
    in_bytes = a.cells * a.genes * 4 if fmt == "dense" else nnz * 8 + (a.cells + 1) * 8
    stage_bytes = in_bytes
    rank_bytes = nnz * 4 + eng.host_plan.n_segments * a.genes * 4 + 24 * G * a.genes
    path_bytes = in_bytes + 24 * G * a.genes + 4 * a.cells
    if t_stage >= t_rank:
        dom, dom_bytes, dom_ms = f"stage_{fmt}_kernel", stage_bytes, t_stage
    else:
        dom, dom_bytes, dom_ms = f"{test}_kernel", rank_bytes, t_rank
    # dense input with count-like values takes the fused single-pass path (fused.cu): the step is
    # table staging + fused_ctab + fused_pass_kernel + per-gene weights + epilogue; fused_pass_kernel dominates.
    # The library times that kernel itself (CUDA events on the launching stream) when ILLICO_PROFILE=1.
    fused_ms = None
    if fmt in ("dense", "csr"):
        os.environ["ILLICO_PROFILE"] = "1"
        fm = []
        for _ in range(4):
            tot = 0.0
            for lb, ub in batches:
                eng.run_batch(M, lb, ub, flags, results, lb)
                torch.cuda.synchronize(dev)
                v = float(lib.illico_last_fused_ms())
                tot = tot + v if v >= 0 else -1.0
                if tot < 0:
                    break
            fm.append(tot)
        os.environ.pop("ILLICO_PROFILE", None)
        if min(fm) >= 0:
            fused_ms = float(np.median(fm[1:]))
            # algorithmic bytes: every streamed row read once (OVO: all but the control's) + one 24-byte record per test
            n_ref = int(grpc.counts[grpc.encoded_ref_group]) if test == "ovo" else 0
            dom, dom_ms = ("fused_pass_kernel" if fmt == "dense" else "fused_csr_pass_kernel"), fused_ms
            if fmt == "dense":
                dom_bytes = (a.cells - n_ref) * a.genes * 4 + 24 * (G - (1 if test == "ovo" else 0)) * a.genes
            else:  # every stored value and index read once + the records
                dom_bytes = nnz * 8 + (a.cells + 1) * 8 + 24 * (G - (1 if test == "ovo" else 0)) * a.genes
    n_launch_dom = len(batches)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None  # dram bytes per launch from the committed ncu --set full capture (same shape only)
    try:
        if (a.cells, a.genes, a.perts) == (300_000, 8_000, 2_000) and n_launch_dom == 1:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                traffic = json.load(f)["bytes_per_launch"].get(dom)
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "launches_per_step": n_launch_dom, "ms_per_launch": round(dom_ms / n_launch_dom, 4),
                "algorithmic_bytes_per_launch": int(dom_bytes / n_launch_dom),
                "stage_ms": round(t_stage, 3), "rank_ms": round(t_rank, 3),
                "fused_ms": None if fused_ms is None else round(fused_ms, 3),
                "note": ("step = table staging (control / cell sample) + fused_ctab + fused_pass_kernel + epilogue; stage_ms / rank_ms "
                         "are the general two-kernel path (continuous data), timed separately for comparison")
                if fused_ms is not None else None,
                "path_achieved_GBps": round(path_bytes / (ms_step * 1e-3) / 1e9, 1),
                "path_frac": round(path_bytes / (ms_step * 1e-3) / 1e9 / peak, 4)}

    # ---- end to end through the public API with host (pinned) input
    e2e = None
    extra = {}
    if not a.no_e2e:
        def host_buffer(shape, dtype):
            try:
                return torch.empty(shape, dtype=dtype, pin_memory=True)
            except RuntimeError:  # not enough lockable memory on this host: pageable (slower H2D, still end to end)
                extra["host_memory"] = "pageable"
                return torch.empty(shape, dtype=dtype)

        if fmt == "dense":
            hostX = host_buffer((a.cells, a.genes), torch.float32)
            hostX.copy_(Xdev)
            Xh = hostX.numpy()
            h2d = Xh.nbytes
        else:
            from scipy import sparse

            pins = [host_buffer(t.shape, t.dtype) for t in (M.data, M.indices, sp.crow_indices().to(torch.int32))]
            for p_, t_ in zip(pins, (M.data, M.indices, sp.crow_indices().to(torch.int32))):
                p_.copy_(t_)
            Xh = sparse.csr_matrix((pins[0].numpy(), pins[1].numpy(), pins[2].numpy()), shape=(a.cells, a.genes))
            h2d = Xh.data.nbytes + Xh.indices.nbytes + Xh.indptr.nbytes * 2
        torch.cuda.synchronize(dev)
        ad = Ad(Xh, labels, a.genes)
        times = []
        for i in range(2 + 5):
            barrier()
            t0 = time.perf_counter()
            out = asymptotic_wilcoxon(ad, is_log1p=False, group_keys="pert", reference=reference, return_array=True,
                                      device=dev)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            if i >= 2:
                times.append(dt)
        tt = torch.tensor([float(np.median(times))], dtype=torch.float64, device=dev)   # median of 5 calls after 2 warm-ups
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
        e2e = {"value": round(n_tests_rank * world / e2e_s, 1), "unit": "tests/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(G * a.genes * 24), "s_per_step": round(e2e_s, 4)}
        if rank == 0:
            import pandas as pd

            from illico_b200.asymptotic_wilcoxon import _result_frame

            t0 = time.perf_counter()
            _result_frame(out[0], out[1], out[2])   # what the default (DataFrame) return adds to the call
            extra["dataframe_s"] = round(time.perf_counter() - t0, 4)
        del out

    # ---- CPU baseline beside it (rank 0, N == 1 only): the oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        import oracle

        threads = oracle.max_threads()
        # ~10 s of CPU work on the K562 shape (the sparse kernels of the reference are ~8 x faster per gene)
        sample = a.cpu_sample_genes or min(a.genes, (32 if fmt == "dense" else 256) * threads)
        Xs = host_sample(Xdev, fmt, sample)
        cpu_run(fmt, test, Xs[:, : min(sample, 4)] if fmt == "dense" else Xs[:, : min(sample, 4)], labels, reference, a.genes,
                min(sample, 4), threads)  # warm the page cache / thread pool
        v, dt = cpu_run(fmt, test, Xs, labels, reference, a.genes, sample, threads)
        cpu = {"value": round(v, 1), "unit": "tests/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} of {a.genes} genes, all {G} groups, {dt:.2f} s wall", "seconds": round(dt, 3)}

    if rank == 0:
        total_tests = n_tests_rank * world
        line = {
            "metric": "gene_x_group_tests_per_s", "value": round(total_tests / (ms_step * 1e-3), 1), "unit": "tests/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "warmup_steps_run": n_warm, "ms_per_step": round(ms_step, 4),
            "wall_s_per_step": round(ms_step * 1e-3, 6),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 keys / int64 ranks / f64 epilogue",
            "data": "synthetic", "impl": "b200",
            "config": {"workload": f"K562-shape {fmt} {test.upper()}: {a.cells} cells x {a.genes} genes x {G} groups per GPU"
                                   + (", reference=non-targeting" if test == "ovo" else "")
                                   + (", continuous values (log1p of normalised counts)" if a.continuous else ""),
                       "format": fmt, "test": test, "cells": a.cells, "genes_per_gpu": a.genes, "groups": G,
                       "nnz_fraction": round(nnz / (a.cells * a.genes), 4), "gene_batches": len(batches),
                       "l2": "inputs larger than L2 (9.6 GB streamed per step), no explicit flush",
                       "sharding": "genes sharded across ranks, no collective in the data path"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_arm(a, fmt, test, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port: the reference is Python + numba and does
    not travel to the GPU box) on all host cores, each step a bounded sample of the same workload."""
    if rank != 0:
        return 0
    import oracle
    from illico_b200 import synth

    threads = oracle.max_threads()
    sample = a.cpu_sample_genes or min(a.genes, (32 if fmt == "dense" else 256) * threads)
    labels, reference = make_labels(a.seed, a.cells, a.perts, test)
    try:
        import torch

        if torch.cuda.is_available():
            Xs = synth.k562_like_torch(a.seed, a.cells, sample, device="cuda").cpu().numpy()
        else:
            raise RuntimeError
    except Exception:
        Xs, _ = synth.k562_like(a.seed, a.cells, sample, a.perts)
    if fmt != "dense":
        from scipy import sparse

        Xs = sparse.csr_matrix(Xs)
    G = len(set(labels))
    vals, secs = [], []
    for i in range(a.warmup + a.steps):
        v, dt = cpu_run(fmt, test, Xs, labels, reference, a.genes, sample, threads)
        if i >= a.warmup:
            vals.append(v); secs.append(dt)
    v = float(np.mean(vals))
    # whole-workload equivalent: N ranks x genes x groups tests at this throughput
    line = {"metric": "gene_x_group_tests_per_s", "value": round(v, 1), "unit": "tests/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(float(np.mean(secs)) * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (CPU)", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": f"K562-shape {fmt} {test.upper()}: {a.cells} cells x {a.genes} genes x {G} groups per GPU"
                                   + (", reference=non-targeting" if test == "ovo" else ""),
                       "format": fmt, "test": test, "cells": a.cells, "genes_per_gpu": a.genes, "groups": G},
            "cpu_baseline": {"value": round(v, 1), "unit": "tests/s", "cores": threads, "kind": "port",
                             "sample": f"first {sample} of {a.genes} genes per step, all {G} groups"},
            "e2e": {"value": round(v, 1), "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
