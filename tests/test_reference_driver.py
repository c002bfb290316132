"""SURVEY.md section 8(f3) / VERDICT r1 next #9: the six GPU dispatchers RUN under the reference's own driver.

``illico_b200.register_into_reference()`` puts them into the reference's ``dispatcher_registry``
(``illico/utils/registry.py:61-64, 193-202``); ``illico.asymptotic_wilcoxon(..., precompile=False)`` -- the unmodified
reference: its batch iterator, joblib thread pool, result assembly and DataFrame -- then computes every batch on the B200.
The result is compared with the committed golden vectors (the same reference with its own numba kernels) and with the
oracle.  Needs the reference importable (``/root/reference`` or ``baseline/_ref``) AND a GPU; skipped otherwise.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import reference_import as R  # noqa: E402
from tests.golden import cases as C  # noqa: E402
from tests.parity import FC_RTOL, assert_parity  # noqa: E402


@pytest.fixture(scope="module")
def reference_with_gpu_dispatchers():
    if R.reference_root() is None:
        pytest.skip("the reference is not importable here (neither /root/reference nor baseline/_ref)")
    try:
        R.import_reference()
        from illico.utils import registry as ref
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference not importable: {e}")
    import illico_b200
    from illico_b200 import dispatch

    saved = dict(ref.dispatcher_registry)
    assert illico_b200.register_into_reference() is True
    yield ref
    ref.dispatcher_registry.clear()
    ref.dispatcher_registry.update(saved)
    dispatch.clear_caches()


@pytest.mark.parametrize("name,fmt,test,n_threads,batch_size", [
    ("batched", "dense", "ovo", 3, 64),      # 300 genes: several batches on joblib threads, concurrently
    ("batched", "csr", "ovr", 2, 128),
    ("batched", "csc", "ovo", 1, 100),
    ("k562_mini", "dense", "ovr", 1, 16),    # < 256 genes: the reference forces one batch, one thread
    ("k562_mini", "csr", "ovo", 1, 16),
    ("highcount", "dense", "ovo", 1, 16),
])
def test_reference_driver_runs_gpu_dispatchers(golden_dir, reference_with_gpu_dispatchers, name, fmt, test, n_threads, batch_size):
    from illico_b200 import _lib

    builder, grid, _bs = C.CASES[name]
    X, labels, reference = builder()
    ref = reference if test == "ovo" else None
    before = _lib.launch_count()
    groups, p, U, fc = R.ref_run(C.to_format(X, fmt), labels, ref, batch_size=batch_size, n_threads=n_threads,
                                 precompile=False)      # precompile would call numba's .compile() on the dispatcher
    assert _lib.launch_count() > before, "the reference's driver did not reach the GPU dispatchers"
    gold = np.load(os.path.join(golden_dir, f"{name}.npz"))
    want = gold[C.combo_key(fmt, test, True, True, "two-sided", False)]
    ref_row = int(np.searchsorted(groups, reference)) if test == "ovo" else None
    assert_parity((p, U, fc), (want[0], want[1], want[2]), ref_row=ref_row, fc_rtol=FC_RTOL,
                  what=f"reference driver + GPU dispatchers {name}:{fmt}:{test}")
    g, po, Uo, fco = oracle.run(C.to_format(X, fmt), labels, ref)
    assert_parity((p, U, fc), (po, Uo, fco), ref_row=ref_row, what="vs oracle")
