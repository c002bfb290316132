"""Shared parity assertions (tolerances from BASELINE.json north_star)."""
import numpy as np

P_RTOL = 1.0e-12      # p-values: relative 1e-12 (reference tests/test_asymptotic_wilcoxon.py:173-178)
P_ATOL = 2.3e-308     # sub-normal p-values cannot agree relatively (libm vs CUDA erfc underflow)
FC_RTOL = 1.0e-10     # fold change, is_log1p=False or f64 log1p data
FC_RTOL_LOG1P_F32 = 1.0e-6  # float32 + is_log1p: the reference evaluates expm1 in float32


def assert_parity(got, want, *, ref_row=None, fc_rtol=FC_RTOL, what=""):
    """got / want: (p, U, fc) triples of [G, N] arrays.  U bit-exact; p and fc to tolerance."""
    gp, gU, gfc = got
    wp, wU, wfc = want
    rows = np.ones(gU.shape[0], dtype=bool)
    if ref_row is not None:
        # OVO reference row: the reference's dense kernel leaves p/U uninitialised there
        rows[ref_row] = False
        assert np.all(gp[ref_row] == 1.0) and np.all(gU[ref_row] == -1.0), f"{what}: reference row must be (1, -1)"
    np.testing.assert_array_equal(gU[rows], wU[rows], err_msg=f"{what}: U statistic not bit-exact")
    np.testing.assert_allclose(gp[rows], wp[rows], rtol=P_RTOL, atol=P_ATOL, err_msg=f"{what}: p-value")
    fin = np.isfinite(wfc)
    np.testing.assert_array_equal(np.isposinf(gfc), np.isposinf(wfc), err_msg=f"{what}: fold change inf pattern")
    np.testing.assert_array_equal(np.isnan(gfc), np.isnan(wfc), err_msg=f"{what}: fold change nan pattern")
    np.testing.assert_allclose(gfc[fin], wfc[fin], rtol=fc_rtol, atol=0.0, err_msg=f"{what}: fold change")
