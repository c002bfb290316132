"""On-disk expression matrices behind the ``obj[:, lb:ub]`` protocol the backed data handlers consume.

The reference streams ``h5py.Dataset`` (dense) and anndata's ``_CSCDataset`` (``illico/utils/registry.py:162-188``);
those containers are registered in :mod:`illico_b200.registry` when the libraries are importable.  ``MemmapDense`` and
``MemmapCSC`` are the same thing without the dependency: plain ``np.memmap`` files (a dense C-order matrix, or a CSC
triplet) read one gene batch at a time.  BASELINE config 4 (K562 shape as on-disk CSC, streamed) runs on them wherever
h5py / anndata are absent.
"""
from __future__ import annotations

import json
import os

import numpy as np
from scipy import sparse


def save_dense(path: str, X: np.ndarray) -> None:
    """Writes ``X`` (C order) to ``path`` with a small JSON header next to it."""
    X = np.ascontiguousarray(X)
    X.tofile(path)
    with open(path + ".json", "w") as f:
        json.dump({"shape": list(X.shape), "dtype": str(X.dtype)}, f)


class MemmapDense:
    """``[n_cells, n_genes]`` C-order matrix in a flat file; ``obj[:, lb:ub]`` reads the batch's column block."""

    def __init__(self, path: str, shape=None, dtype=None):
        if shape is None:
            with open(path + ".json") as f:
                meta = json.load(f)
            shape, dtype = tuple(meta["shape"]), meta["dtype"]
        self._m = np.memmap(path, dtype=np.dtype(dtype or np.float32), mode="r", shape=tuple(shape))
        self.shape, self.dtype = tuple(shape), self._m.dtype

    def __getitem__(self, key):
        return np.asarray(self._m[key])


def save_csc(directory: str, data, indices, indptr, shape) -> None:
    """Writes a CSC triplet as three flat files (+ ``meta.json``) into ``directory``."""
    os.makedirs(directory, exist_ok=True)
    data, indices, indptr = np.ascontiguousarray(data), np.ascontiguousarray(indices), np.ascontiguousarray(indptr, dtype=np.int64)
    data.tofile(os.path.join(directory, "data.bin"))
    indices.tofile(os.path.join(directory, "indices.bin"))
    indptr.tofile(os.path.join(directory, "indptr.bin"))
    with open(os.path.join(directory, "meta.json"), "w") as f:
        json.dump({"shape": [int(shape[0]), int(shape[1])], "data_dtype": str(data.dtype), "index_dtype": str(indices.dtype),
                   "nnz": int(data.size)}, f)


class MemmapCSC:
    """CSC matrix on disk; ``obj[:, lb:ub]`` returns the batch's columns as an in-memory ``scipy.sparse.csc_matrix``
    (what anndata's backed ``_CSCDataset`` returns for the same slice)."""

    def __init__(self, directory: str):
        with open(os.path.join(directory, "meta.json")) as f:
            meta = json.load(f)
        self.shape = (int(meta["shape"][0]), int(meta["shape"][1]))
        nnz = int(meta["nnz"])
        self.dtype = np.dtype(meta["data_dtype"])
        self._data = np.memmap(os.path.join(directory, "data.bin"), dtype=self.dtype, mode="r", shape=(nnz,))
        self._indices = np.memmap(os.path.join(directory, "indices.bin"), dtype=np.dtype(meta["index_dtype"]), mode="r", shape=(nnz,))
        self.indptr = np.fromfile(os.path.join(directory, "indptr.bin"), dtype=np.int64)   # small: kept in RAM
        self.nbytes = self._data.nbytes + self._indices.nbytes + self.indptr.nbytes

    def __getitem__(self, key):
        rows, cols = key
        if rows != slice(None) or not isinstance(cols, slice) or cols.step not in (None, 1):
            raise IndexError("MemmapCSC supports obj[:, lb:ub] only")
        lb, ub, _ = cols.indices(self.shape[1])
        lo, hi = int(self.indptr[lb]), int(self.indptr[ub])
        return sparse.csc_matrix((np.asarray(self._data[lo:hi]), np.asarray(self._indices[lo:hi]), self.indptr[lb:ub + 1] - lo),
                                 shape=(self.shape[0], ub - lb))
