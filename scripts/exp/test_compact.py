"""Held back until verified on the GPU: moves into tests/test_gpu_parity.py once green."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from tests.parity import assert_parity  # noqa: E402
from tests.test_gpu_parity import _fused_case, _run  # noqa: E402


@pytest.mark.parametrize("test", ["ovo", "ovr"])
def test_fused_dense_compacted_hand_back(monkeypatch, test):
    """Scattered handed-back genes (here 7 of 72): their columns are gathered into a compact matrix, ranked by the
    general path and scattered back; same answer as merged runs and as the general path alone."""
    X, labels, ref = _fused_case("middle")
    X = np.abs(X)
    X[X > 1e20] = 3.0
    reference = ref if test == "ovo" else None
    monkeypatch.setenv("ILLICO_FUSED_COMPACT", "1")
    groups, compact = _run(X, labels, reference, is_log1p=False)
    monkeypatch.setenv("ILLICO_FUSED_COMPACT", "0")
    monkeypatch.setenv("ILLICO_FUSED_MAX_HANDBACK", "1.0")
    _, merged = _run(X, labels, reference, is_log1p=False)
    monkeypatch.setenv("ILLICO_OVO_FUSED", "0")
    monkeypatch.setenv("ILLICO_OVR_FUSED", "0")
    _, general = _run(X, labels, reference, is_log1p=False)
    for a, b in zip(compact, merged):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(compact[1], general[1])
    rows = np.ones(len(groups), bool)
    if reference is not None:
        rows[int(np.searchsorted(groups, reference))] = False
    np.testing.assert_allclose(compact[0][rows], general[0][rows], rtol=1e-13, atol=2.3e-308)
    g, p, U, fc = oracle.run(X, labels, reference, is_log1p=False)
    ref_row = int(np.searchsorted(groups, reference)) if reference is not None else None
    assert_parity(compact, (p, U, fc), ref_row=ref_row, what=f"compacted hand-back {test}")
