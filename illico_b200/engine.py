"""Device engine: owns the plan tables, the batch buffers and the stream, and drives the C-ABI calls.

PyTorch is plumbing here (device memory, streams, H2D/D2H copies); every compute kernel is in
``libillico_b200.so``.  One engine per (device, group plan).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .groups import GroupContainer, HostPlan, build_plan

DENSE, CSC, CSR = "dense", "csc", "csr"


def _env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


@dataclass
class DeviceMatrix:
    """An expression matrix resident in HBM, in one of the three kernel formats."""

    fmt: str
    shape: tuple
    data: torch.Tensor            # dense: [n, ld] float32 ; sparse: [nnz] float32
    indices: torch.Tensor | None = None   # int32 [nnz]
    indptr: torch.Tensor | None = None    # int64
    gene_offset: int = 0          # genes [gene_offset, gene_offset + shape[1]) of the caller's matrix

    @property
    def ld(self) -> int:
        return int(self.data.stride(0)) if self.fmt == DENSE else 0


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.IllicoCudaError("illico_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise _lib.IllicoCudaError(f"illico_b200 runs on CUDA devices only, got {dev}")
    return dev


class Engine:
    def __init__(self, grpc: GroupContainer, device=None, seg_max: int | None = None):
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.grpc = grpc
        self.host_plan: HostPlan = build_plan(grpc, seg_max or _env_int("ILLICO_B200_SEG_MAX", 512))
        hp = self.host_plan
        with torch.cuda.device(self.device):
            self._tables = {k: torch.from_numpy(getattr(hp, k)).to(self.device) for k in HostPlan.TABLES}
        self.plan = _lib.Plan(
            n_cells=hp.n_cells, n_groups=hp.n_groups, n_segments=hp.n_segments, ref_group=hp.ref_group,
            max_group_size=hp.max_group_size, ref_group_size=hp.ref_group_size, ref_seg_begin=hp.ref_seg_begin,
            ref_seg_end=hp.ref_seg_end, slot_cap=hp.slot_cap,
            **{k: self._tables[k].data_ptr() for k in HostPlan.TABLES})
        self._buf_genes = 0
        self._ir_vals = self._ir_cnt = self._ws = None
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    # ---- sizes ---------------------------------------------------------------------------------------
    @property
    def n_groups(self) -> int:
        return self.host_plan.n_groups

    @property
    def is_ovo(self) -> bool:
        return self.host_plan.ref_group >= 0

    def max_batch_genes(self, n_genes: int) -> int:
        """Genes per device batch so that the staged lists stay within the memory budget."""
        free, _total = torch.cuda.mem_get_info(self.device)
        budget = min(_env_int("ILLICO_B200_IR_BYTES", 16 << 30), int(free * 0.35))
        per_gene = self.host_plan.slot_cap * 4 + self.host_plan.n_segments * 4
        b = max(1, budget // max(per_gene, 1))
        b = int(min(n_genes, b, _env_int("ILLICO_B200_BATCH_GENES", 1 << 30)))
        return b if b < 8 or b == n_genes else (b // 4) * 4  # multiples of 4 keep the 128-bit staging path

    def _ensure_buffers(self, b: int) -> None:
        if b <= self._buf_genes:
            return
        hp = self.host_plan
        self._ir_vals = self._ir_cnt = self._ws = None
        with torch.cuda.device(self.device):
            self._ir_vals = torch.empty(b * hp.slot_cap, dtype=torch.float32, device=self.device)
            self._ir_cnt = torch.empty(b * hp.n_segments, dtype=torch.int32, device=self.device)
            ws = int(self.lib.illico_rank_workspace_bytes(C.byref(self.plan), b))
            if ws == 0:
                raise _lib.IllicoCudaError("illico_rank_workspace_bytes failed: " + self.lib.illico_last_error().decode())
            self._ws = torch.empty(ws, dtype=torch.uint8, device=self.device)
        self._buf_genes = b

    # ---- uploads (plumbing) -----------------------------------------------------------------------------
    def upload_dense(self, X: np.ndarray, gene_lb: int = 0, gene_ub: int | None = None) -> DeviceMatrix:
        return upload_dense(X, self.device, gene_lb, gene_ub)

    def upload_sparse(self, X, fmt: str) -> DeviceMatrix:
        return upload_sparse(X, fmt, self.device)

    def check_csr_sorted(self, M: DeviceMatrix) -> bool:
        st = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.illico_check_csr_sorted(M.indices.data_ptr(), M.indptr.data_ptr(), M.shape[0],
                                              self._flag.data_ptr(), st)
        if rc < 0:
            _lib.check(1, "illico_check_csr_sorted")
        return rc == 1

    # ---- one gene batch ----------------------------------------------------------------------------------
    def run_batch(self, M: DeviceMatrix, lb: int, ub: int, flags: _lib.Flags, results: torch.Tensor,
                  result_gene0: int, debug: dict | None = None) -> None:
        """Enqueues stage + rank for genes ``[lb, ub)`` of ``M``; writes ``results[:, result_gene0 + k, :]``.

        ``results`` is a device tensor ``[G, N_total, 3]`` float64 (contiguous).
        """
        b = ub - lb
        if b <= 0:
            return
        if M.shape[0] != self.host_plan.n_cells:
            raise ValueError(f"matrix has {M.shape[0]} cells, groups describe {self.host_plan.n_cells}")
        if lb < 0 or ub > M.shape[1]:
            raise ValueError(f"Invalid chunk bounds: {(lb, ub)} for data with {M.shape[1]} columns.")
        self._ensure_buffers(b)
        G, Ntot = results.shape[0], results.shape[1]
        assert results.dtype == torch.float64 and results.is_contiguous() and G == self.n_groups
        st = torch.cuda.current_stream(self.device).cuda_stream
        buf = _lib.BatchBuffers(self._ir_vals.data_ptr(), self._ir_cnt.data_ptr(), self._ws.data_ptr(), self._ws.numel())
        dbg = None
        if debug is not None:
            shape_t = (G, b) if self.is_ovo else (b,)
            debug["u2"] = torch.zeros((G, b), dtype=torch.int64, device=self.device)
            debug["tie_sum"] = torch.zeros(shape_t, dtype=torch.float64, device=self.device)
            debug["tie_exact"] = torch.zeros(shape_t, dtype=torch.int64, device=self.device)
            dbg = _lib.Debug(debug["u2"].data_ptr(), debug["tie_sum"].data_ptr(), debug["tie_exact"].data_ptr())
        out_ptr = results.data_ptr() + result_gene0 * 3 * 8
        gstride = Ntot * 3
        test = "ovo" if self.is_ovo else "ovr"
        dbg_ref = C.byref(dbg) if dbg is not None else None
        with torch.cuda.device(self.device):
            if M.fmt == DENSE:
                fn = getattr(self.lib, f"illico_{test}_dense_f32")
                rc = fn(M.data.data_ptr(), M.ld, lb, b, C.byref(self.plan), C.byref(flags), C.byref(buf), out_ptr,
                        gstride, dbg_ref, st)
            else:
                fn = getattr(self.lib, f"illico_{test}_{M.fmt}_f32")
                rc = fn(M.data.data_ptr(), M.indices.data_ptr(), M.indptr.data_ptr(), lb, b, C.byref(self.plan),
                        C.byref(flags), C.byref(buf), out_ptr, gstride, dbg_ref, st)
        _lib.check(rc, f"illico_{test}_{M.fmt}_f32")


def upload_dense(X: np.ndarray, device, gene_lb: int = 0, gene_ub: int | None = None) -> DeviceMatrix:
    """Enqueues the copy of ``X[:, gene_lb:gene_ub]`` (any real dtype) into a float32 device matrix.
    Asynchronous when ``X`` lives in pinned memory: the caller can keep working on the host meanwhile."""
    device = require_cuda(device)
    n, N = X.shape
    gene_ub = N if gene_ub is None else gene_ub
    view = X[:, gene_lb:gene_ub]
    with torch.cuda.device(device):
        t = torch.from_numpy(view) if isinstance(view, np.ndarray) else view
        if t.dtype == torch.float32:
            d = torch.empty((n, gene_ub - gene_lb), dtype=torch.float32, device=device)
            d.copy_(t, non_blocking=True)
        else:
            d = _to_f32_exact(t.to(device))
    return DeviceMatrix(DENSE, (n, gene_ub - gene_lb), d, gene_offset=gene_lb)


def upload_sparse(X, fmt: str, device) -> DeviceMatrix:
    device = require_cuda(device)
    with torch.cuda.device(device):
        data = torch.from_numpy(np.ascontiguousarray(X.data))
        data = data.to(device, non_blocking=True)
        data = data if data.dtype == torch.float32 else _to_f32_exact(data)
        indices = torch.from_numpy(np.ascontiguousarray(X.indices, dtype=np.int32)).to(device, non_blocking=True)
        indptr = torch.from_numpy(np.ascontiguousarray(X.indptr, dtype=np.int64)).to(device, non_blocking=True)
    return DeviceMatrix(fmt, tuple(X.shape), data, indices, indptr)


def _to_f32_exact(t: torch.Tensor) -> torch.Tensor:
    """Converts a device tensor to float32, refusing to create ties that are not in the data."""
    f = t.to(torch.float32)
    if not bool((f.to(t.dtype) == t).all()):
        raise NotImplementedError(
            f"values of dtype {t.dtype} are not exactly representable in float32; the CUDA path ranks float32 keys "
            "(64-bit keys are not implemented yet)")
    return f


def make_flags(is_log1p: bool, use_continuity: bool, tie_correct: bool, alternative: str, fmt: str) -> _lib.Flags:
    if alternative not in _lib.ALTERNATIVES:
        raise ValueError(f"Unsupported alternative hypothesis: {alternative}")
    return _lib.Flags(int(bool(is_log1p)), int(bool(use_continuity)), int(bool(tie_correct)),
                      _lib.ALTERNATIVES[alternative], _lib.TIES_DENSE if fmt == DENSE else _lib.TIES_SPARSE)
