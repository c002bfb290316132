"""Group encoding and the device "plan" derived from it.

``encode_and_count_groups`` mirrors the reference (``illico/utils/groups.py:18-58``): same
``GroupContainer`` fields, same lexicographic group order (``np.unique``), same ``ValueError``
when the reference label is absent, ``encoded_ref_group == -1`` for one-versus-rest.  The
reference's per-label Python dict loop (0.33 s at 300k cells) is replaced by a vectorised
factorisation; the result is identical.

``build_plan`` turns a GroupContainer into the tables the CUDA kernels read
(``illico_plan_t`` in ``include/illico_b200.h``): cells permuted to group-contiguous order and every
group cut into segments of at most ``seg_max`` cells.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Any

import numpy as np

GroupContainer = namedtuple(
    "GroupContainer",
    ["encoded_groups", "counts", "indices", "indptr", "encoded_ref_group"],
)

SEG_MAX_DEFAULT = 512


def encode_and_count_groups(groups, ref_group: Any):
    """Returns ``(unique_groups, GroupContainer)`` exactly like the reference."""
    arr = np.asarray(groups)
    if arr.ndim != 1:
        raise ValueError("group labels must be one-dimensional")
    if arr.dtype == object:
        # mixed / str objects: np.unique on an object array compares Python objects; convert the common
        # all-str case to a unicode array first (what the reference gets from a list of str)
        try:
            arr = arr.astype(str) if all(isinstance(x, str) for x in arr[: min(arr.size, 1000)]) else arr
        except Exception:  # pragma: no cover
            pass
    unique_groups, inverse, counts = np.unique(arr, return_inverse=True, return_counts=True)
    if ref_group is not None:
        hit = np.nonzero(unique_groups == ref_group)[0]
        if hit.size == 0:
            raise ValueError(f"Reference group `{ref_group}` is not present in the group labels.")
        encoded_ref = int(hit[0])
    else:
        encoded_ref = -1
    encoded = np.ascontiguousarray(inverse.reshape(-1), dtype=np.int64)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    indices = np.argsort(encoded, kind="stable").astype(np.int64)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return unique_groups, GroupContainer(encoded, counts, indices, indptr, encoded_ref)


class HostPlan:
    """numpy form of ``illico_plan_t``."""

    def __init__(self, grpc: GroupContainer, seg_max: int = SEG_MAX_DEFAULT):
        enc = np.asarray(grpc.encoded_groups, dtype=np.int64)
        counts = np.asarray(grpc.counts, dtype=np.int64)
        n, G = enc.size, counts.size
        if n >= 2**31 - 1:
            raise ValueError("more than 2^31 cells are not supported")
        if G == 0 or n == 0:
            raise ValueError("empty input")
        if counts.sum() != n or counts.min() < 1:
            raise ValueError("inconsistent GroupContainer")
        self.n_cells, self.n_groups = int(n), int(G)
        self.ref_group = int(grpc.encoded_ref_group)
        self.seg_max = int(seg_max)
        perm = np.argsort(enc, kind="stable")  # stable: ascending cell index inside a group
        group_off = np.concatenate([[0], np.cumsum(counts)])
        nseg = np.maximum(1, -(-counts // seg_max))
        group_seg = np.concatenate([[0], np.cumsum(nseg)])
        S = int(group_seg[-1])
        seg_group = np.repeat(np.arange(G), nseg)
        k = np.arange(S) - group_seg[seg_group]
        seg_start = group_off[seg_group] + k * seg_max
        seg_len = np.minimum(seg_max, counts[seg_group] - k * seg_max)
        seg_pos = np.concatenate([seg_start, [n]])
        seg_base = np.concatenate([[0], np.cumsum((seg_len + 3) // 4 * 4)])
        if seg_base[-1] >= 2**31 - 1:
            raise ValueError("slot space exceeds 2^31")
        pos_seg = np.repeat(np.arange(S), seg_len)  # segment of each position
        cell_seg = np.empty(n, dtype=np.int64)
        cell_seg[perm] = pos_seg
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
        self.n_segments = S
        self.max_group_size = int(counts.max())
        self.ref_group_size = int(counts[self.ref_group]) if self.ref_group >= 0 else 0
        self.slot_cap = int(seg_base[-1])
        self.perm, self.cell_seg = i32(perm), i32(cell_seg)
        self.seg_pos, self.seg_base, self.seg_group = i32(seg_pos), i32(seg_base), i32(seg_group)
        self.group_seg, self.group_size = i32(group_seg), i32(counts)

    TABLES = ("perm", "cell_seg", "seg_pos", "seg_base", "seg_group", "group_seg", "group_size")


def build_plan(grpc: GroupContainer, seg_max: int = SEG_MAX_DEFAULT) -> HostPlan:
    return HostPlan(grpc, seg_max)
