"""Bitwise parity at the sizes and configurations the bench numbers are quoted on (BASELINE.json configs 2-5).

The checker is the UNMODIFIED reference C++ (oracle/_ref/libref_bp.so, shipped prebuilt to the GPU box) on all host
cores; without it, the plain-C restatement (pinned to the reference bit-for-bit by tests/test_oracle_cpu.py).  Bar:
hard decisions, converge flags and iteration counts identical on every row; min-sum LLRs bit-identical, product-sum
LLRs within 1e-5 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from ldpc_b200 import BpDecoder, BpOsdDecoder, codes
from util import assert_llr_close, checker_decode

pytestmark = pytest.mark.gpu


def _same(got_dec, d, want, what):
    rows = (got_dec != want[0]).any(axis=1)
    assert not rows.any(), f"{what}: hard decisions differ in {int(rows.sum())} of {rows.size} rows (checker {want[4]})"
    assert np.array_equal(np.asarray(d.converge_batch, bool), np.asarray(want[1], bool)), f"{what}: converge differs"
    assert np.array_equal(d.iter_batch, want[2]), f"{what}: iteration counts differ"


@pytest.mark.parametrize("kernel", ["smem", "pair", "stream", "edge"])
def test_config2_n1000_minsum_2p17(kernel):
    """BASELINE config 2 on 2^17 syndromes (seed 7, the first 2^17 of the 2^20 workload), every kernel family."""
    H = codes.regular_ldpc(1000, 3, 6, seed=1)
    B = 1 << 17
    syn = codes.bsc_syndromes(H, 0.05, B, seed=7)
    kw = dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625)
    want = checker_decode(H, syn, 0.05, want_llr=False, **kw)
    d = BpDecoder(H, error_rate=0.05, input_vector_type="syndrome", kernel=kernel, **kw)
    got = d.decode_batch(syn)
    _same(got, d, want, f"config 2 / {kernel}")
    assert 0.985 < want[1].mean() < 0.999
    # LLRs bit-identical on a 2^13 slice that contains non-convergers
    sl = slice(0, 1 << 13)
    wl = checker_decode(H, syn[sl], 0.05, want_llr=True, **kw)
    d.decode_batch(syn[sl], return_llr=True)
    assert_llr_close(d.log_prob_ratios_batch, wl[3], 1e-5, exact=True)
    assert not wl[1].all()


@pytest.mark.parametrize("p", [0.02, 0.05, 0.08])
def test_config5_n10000_serial_maxiter100(p):
    """BASELINE config 5: n = 10^4 (3,6)-LDPC, serial min-sum, max_iter = 100, one point of the error-rate sweep."""
    H = codes.regular_ldpc(10000, 3, 6, seed=1)
    B = 1408
    syn = codes.bsc_syndromes(H, p, B, seed=int(1000 * p))
    kw = dict(max_iter=100, bp_method="ms", schedule="serial", ms_scaling_factor=0.625)
    want = checker_decode(H, syn, p, want_llr=True, **kw)
    d = BpDecoder(H, error_rate=p, input_vector_type="syndrome", **kw)
    got = d.decode_batch(syn, return_llr=True)
    _same(got, d, want, f"config 5 p={p}")
    assert_llr_close(d.log_prob_ratios_batch, want[3], 1e-5, exact=True)
    if p >= 0.08:
        assert (want[2] == 100).any()  # non-convergers run all 100 iterations


def test_config5_n10000_parallel_maxiter100():
    H = codes.regular_ldpc(10000, 3, 6, seed=1)
    syn = np.concatenate([codes.bsc_syndromes(H, 0.05, 1024, seed=7), codes.bsc_syndromes(H, 0.085, 256, seed=8)])
    kw = dict(max_iter=100, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625)
    want = checker_decode(H, syn, 0.05, want_llr=True, **kw)
    d = BpDecoder(H, error_rate=0.05, input_vector_type="syndrome", **kw)
    got = d.decode_batch(syn, return_llr=True)
    _same(got, d, want, "config 5 parallel")
    assert_llr_close(d.log_prob_ratios_batch, want[3], 1e-5, exact=True)
    assert not want[1].all()


def test_config3_surface13_bposd_1e5():
    """BASELINE config 3 with the OSD-0 post-processor: d = 13 rotated surface code, product-sum 30 iterations,
    10^5 syndromes (~85 % of them go through OSD-0)."""
    H = codes.rotated_surface_code_x(13)
    B = 100000
    syn = codes.bsc_syndromes(H, 0.05, B, seed=3)
    kw = dict(max_iter=30, bp_method="ps", schedule="parallel")
    want = checker_decode(H, syn, 0.05, osd=True, want_llr=False, **kw)
    d = BpOsdDecoder(H, error_rate=0.05, osd_method="osd0", **kw)
    got = d.decode_batch(syn)
    _same(got, d, want, "config 3 BP+OSD-0")
    assert (~want[1]).mean() > 0.5
    assert np.array_equal(codes.syndromes_of(H, got), syn)
    # plain BP LLRs within 1e-5 on a slice
    wl = checker_decode(H, syn[:8192], 0.05, want_llr=True, **kw)
    b = BpDecoder(H, error_rate=0.05, input_vector_type="syndrome", **kw)
    b.decode_batch(syn[:8192], return_llr=True)
    assert_llr_close(b.log_prob_ratios_batch, wl[3], 1e-5)


@pytest.mark.parametrize("p", [0.003, 0.02])
def test_config4_bb144_bposd_1e5(p):
    """BASELINE config 4: [[144,12,12]] bivariate bicycle code, min-sum + OSD-0, 10^5 syndromes."""
    H = codes.bivariate_bicycle_144()
    B = 100000
    syn = codes.bsc_syndromes(H, p, B, seed=5)
    kw = dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625)
    want = checker_decode(H, syn, p, osd=True, want_llr=False, **kw)
    d = BpOsdDecoder(H, error_rate=p, osd_method="osd0", **kw)
    got = d.decode_batch(syn)
    _same(got, d, want, f"config 4 p={p}")
    assert np.array_equal(codes.syndromes_of(H, got), syn)
