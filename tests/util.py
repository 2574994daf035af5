"""Shared helpers for the parity tests."""
import numpy as np


def assert_llr_close(got, want, rtol, exact=False):
    """Posterior LLR parity: +-inf and NaN positions must match; finite values bit-exact or within rtol."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    assert np.array_equal(np.isinf(got), np.isinf(want)), "inf pattern differs"
    inf = np.isinf(want)
    assert np.array_equal(np.sign(got[inf]), np.sign(want[inf])), "inf signs differ"
    fin = np.isfinite(want)
    if exact:
        assert np.array_equal(got[fin], want[fin]), \
            f"LLRs not bit-identical: max abs diff {np.max(np.abs(got[fin] - want[fin]))}"
    else:
        denom = np.maximum(np.abs(want[fin]), 1e-300)
        rel = np.abs(got[fin] - want[fin]) / denom
        assert rel.size == 0 or rel.max() <= rtol, f"LLR relative error {rel.max()} > {rtol}"


def assert_same_decode(got, want, llr_rtol=1e-5, llr_exact=False):
    """got/want = (decoding, converged, iters, llr)."""
    assert np.array_equal(got[0], want[0]), f"hard decisions differ in {(got[0] != want[0]).any(axis=1).sum()} rows"
    assert np.array_equal(np.asarray(got[1], bool), np.asarray(want[1], bool)), "converge flags differ"
    assert np.array_equal(got[2], want[2]), "iteration counts differ"
    if got[3] is not None and want[3] is not None:
        assert_llr_close(got[3], want[3], llr_rtol, exact=llr_exact)
