"""GPU parity: the CUDA path (through the C-ABI / the Python shim) against the CPU oracle on the same
seeded syndromes.  Bar (BASELINE.json north_star): hard decisions, converge flags and iteration counts
bit-exact; posterior LLRs within 1e-5 relative (min-sum: bit-identical), +-inf / NaN positions identical."""
import numpy as np
import pytest

from ldpc_b200 import BpDecoder, BpOsdDecoder, codes
from util import assert_same_decode

pytestmark = pytest.mark.gpu


FAMILIES = ["stream", "smem", "pair", "edge"]  # "pair" and "edge" serve the parallel schedule only


def _decode_gpu(H, syn, p, kernel="auto", **kw):
    dec = BpDecoder(H, kernel=kernel, error_channel=np.broadcast_to(np.asarray(p, float), (H.shape[1],)).copy(),
                    input_vector_type="syndrome", **kw)
    out = dec.decode_batch(syn, return_llr=True)
    return out, dec.converge_batch, dec.iter_batch, dec.log_prob_ratios_batch


@pytest.fixture(scope="module")
def H1000():
    return codes.regular_ldpc(1000, 3, 6, seed=1)


@pytest.mark.parametrize("method", ["ms", "ps"])
@pytest.mark.parametrize("schedule", ["parallel", "serial"])
@pytest.mark.parametrize("ms_scaling", [0.625, 0.0, 1.0])
@pytest.mark.parametrize("kernel", FAMILIES)
def test_regular_n1000(port_oracle, H1000, method, schedule, ms_scaling, kernel):
    if method == "ps" and ms_scaling != 0.625:
        pytest.skip("ms_scaling_factor is unused by product_sum")
    if kernel in ("edge", "pair") and schedule == "serial":
        pytest.skip("the edge-parallel and paired families serve the parallel schedule")
    B = 1536 if method == "ms" else 512
    syn = np.concatenate([codes.bsc_syndromes(H1000, 0.05, B, seed=7),
                          codes.bsc_syndromes(H1000, 0.09, B // 8, seed=8)])  # the second part mostly fails
    kw = dict(max_iter=50, bp_method=method, schedule=schedule, ms_scaling_factor=ms_scaling)
    want = port_oracle.decode_batch(H1000, syn, 0.05, **kw)
    got = _decode_gpu(H1000, syn, 0.05, kernel=kernel, **kw)
    assert_same_decode(got, want, llr_exact=(method == "ms"))
    assert 0 < want[1].mean() < 1  # both convergers and non-convergers are covered


@pytest.mark.parametrize("kernel", FAMILIES)
def test_surface_d13_product_sum(port_oracle, kernel):
    H = codes.rotated_surface_code_x(13)
    syn = codes.bsc_syndromes(H, 0.05, 4096, seed=3)
    kw = dict(max_iter=30, bp_method="ps", schedule="parallel")
    assert_same_decode(_decode_gpu(H, syn, 0.05, kernel=kernel, **kw), port_oracle.decode_batch(H, syn, 0.05, **kw))


@pytest.mark.parametrize("kernel", FAMILIES)
def test_hamming5_readme_config(port_oracle, kernel):
    if kernel == "pair":
        pytest.skip("row degree 16 with column degree 5 is beyond the paired family's buckets")
    H = codes.hamming_code(5)
    rng = np.random.default_rng(0)
    syn = rng.integers(0, 2, size=(100, 5)).astype(np.uint8)
    kw = dict(max_iter=2, bp_method="product_sum", schedule="parallel")
    assert_same_decode(_decode_gpu(H, syn, 0.1, kernel=kernel, **kw), port_oracle.decode_batch(H, syn, 0.1, **kw))


@pytest.mark.parametrize("method", ["ms", "ps"])
@pytest.mark.parametrize("kernel", FAMILIES)
def test_nonuniform_channel_with_certain_bits(port_oracle, method, kernel):
    """error_channel with p = 0 entries gives infinite priors (reference python_test/test_bp_decoder.py:188-192)."""
    H = codes.regular_ldpc(120, 3, 6, seed=5)
    rng = np.random.default_rng(11)
    p = rng.uniform(0.01, 0.2, size=120)
    p[::17] = 0.0
    err = (rng.random((600, 120)) < p).astype(np.uint8)
    syn = codes.syndromes_of(H, err)
    kw = dict(max_iter=20, bp_method=method, schedule="parallel", ms_scaling_factor=0.75)
    want = port_oracle.decode_batch(H, syn, p, **kw)
    assert_same_decode(_decode_gpu(H, syn, p, kernel=kernel, **kw), want, llr_exact=(method == "ms"))
    assert np.isinf(want[3]).any()


@pytest.mark.parametrize("kernel", FAMILIES)
def test_custom_serial_order(port_oracle, kernel):
    if kernel in ("edge", "pair"):
        pytest.skip("the edge-parallel and paired families serve the parallel schedule")
    H = codes.regular_ldpc(200, 3, 6, seed=2)
    order = np.random.default_rng(4).permutation(200)
    syn = codes.bsc_syndromes(H, 0.06, 700, seed=9)
    for method in ("ms", "ps"):
        kw = dict(max_iter=25, bp_method=method, schedule="serial", ms_scaling_factor=0.9)
        want = port_oracle.decode_batch(H, syn, 0.06, serial_schedule_order=order, **kw)
        got = _decode_gpu(H, syn, 0.06, kernel=kernel, serial_schedule_order=[int(x) for x in order], **kw)
        assert_same_decode(got, want, llr_exact=(method == "ms"))


@pytest.mark.parametrize("kernel", FAMILIES)
def test_irregular_degrees(port_oracle, kernel):
    """Row degrees 1..~12 and column degrees 1..~9: exercises the larger degree buckets and degree-1 rows
    (magnitude DBL_MAX * alpha)."""
    if kernel == "pair":
        pytest.skip("degrees (12, 9) are beyond the paired family's buckets")
    rng = np.random.default_rng(21)
    m, n = 60, 90
    dense = (rng.random((m, n)) < 0.07).astype(np.uint8)
    dense[0, :] = 0
    dense[0, 5] = 1  # a degree-1 check
    dense[:, dense.sum(0) == 0] |= (rng.random((m, 1)) < 0.05).astype(np.uint8)
    err = (rng.random((800, n)) < 0.04).astype(np.uint8)
    import scipy.sparse as sp
    H = sp.csr_matrix(dense)
    syn = codes.syndromes_of(H, err)
    for method, sched in (("ms", "parallel"), ("ps", "parallel"), ("ms", "serial"), ("ps", "serial")):
        if kernel in ("edge", "pair") and sched == "serial":
            continue
        kw = dict(max_iter=15, bp_method=method, schedule=sched, ms_scaling_factor=0.625)
        want = port_oracle.decode_batch(H, syn, 0.04, **kw)
        assert_same_decode(_decode_gpu(H, syn, 0.04, kernel=kernel, **kw), want, llr_exact=(method == "ms"))


def test_received_vector_input(port_oracle):
    """decode(v) == v ^ decode(Hv) (reference bp.hpp:162-180); n != m so AUTO resolves by length."""
    H = codes.regular_ldpc(120, 3, 6, seed=5)
    err = codes.bsc_errors(120, 0.05, 300, seed=1)
    dec = BpDecoder(H, error_rate=0.05, max_iter=20, bp_method="ps")
    got = dec.decode_batch(err)
    want = port_oracle.decode_batch(H, codes.syndromes_of(H, err), 0.05, max_iter=20, bp_method="ps")[0] ^ err
    assert np.array_equal(got, want)


def test_bposd_bb144(port_oracle):
    """Config 4: [[144,12,12]] BB code, min-sum + OSD-0 on the BP failures; compare against the oracle's
    BP followed by its OSD-0 restatement, and require H x == s for every row."""
    H = codes.bivariate_bicycle_144()
    syn = codes.bsc_syndromes(H, 0.02, 3000, seed=5)  # p high enough that BP fails sometimes
    d = BpOsdDecoder(H, error_rate=0.02, bp_method="ms", ms_scaling_factor=0.625, max_iter=50, osd_method="osd0")
    got = d.decode_batch(syn)
    bp = port_oracle.decode_batch(H, syn, 0.02, max_iter=50, bp_method="ms", ms_scaling_factor=0.625)
    want = bp[0].copy()
    bad = ~bp[1]
    assert bad.any()
    want[bad] = port_oracle.osd0_batch(H, syn[bad], bp[3][bad])
    assert np.array_equal(got, want)
    assert np.array_equal(codes.syndromes_of(H, got), syn)


@pytest.mark.parametrize("kernel", ["stream", "smem"])
def test_full_size_round_trip(H1000, kernel):
    """BASELINE config 2 at full size (2^20 syndromes, min-sum 50 iterations): size-independent properties.
    Every converged row reproduces its syndrome; iteration counts are in range; the statistics match the
    oracle's on this code (~99.3 % convergence, ~8.3 mean iterations)."""
    B = 1 << 20
    syn = codes.bsc_syndromes(H1000, 0.05, B, seed=7)
    d = BpDecoder(H1000, error_rate=0.05, max_iter=50, bp_method="ms", ms_scaling_factor=0.625,
                  input_vector_type="syndrome", kernel=kernel)
    dec = d.decode_batch(syn)
    conv, its = d.converge_batch, d.iter_batch
    assert dec.shape == (B, 1000) and dec.max() <= 1
    assert its.min() >= 1 and its.max() <= 50
    assert np.all(its[~conv] == 50)
    assert 0.985 < conv.mean() < 0.999
    assert 7.5 < its.mean() < 9.5
    idx = np.nonzero(conv)[0][:: 37]
    assert np.array_equal(codes.syndromes_of(H1000, dec[idx]), syn[idx])
    # batch-order independence: a permuted sub-batch decodes to the permuted results
    perm = np.random.default_rng(0).permutation(1 << 14)
    dec2 = d.decode_batch(syn[perm])
    assert np.array_equal(dec2, dec[perm])
    assert np.array_equal(d.iter_batch, its[perm])


@pytest.mark.parametrize("kernel", FAMILIES)
def test_golden_fixtures_from_reference(kernel):
    """tests/golden/*.npz were produced by the unmodified reference C++ (tests/golden/make_golden.py)."""
    import glob
    import os
    import scipy.sparse as sp
    paths = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
    assert paths
    for path in paths:
        z = np.load(path, allow_pickle=False)
        H = sp.csr_matrix((np.ones(z["rows"].size, np.uint8), (z["rows"], z["cols"])), shape=tuple(z["shape"]))
        if "kind" in z.files and str(z["kind"]) == "soft_info":
            if kernel != FAMILIES[0]:
                continue  # the soft-information kernel has no family choice: check it once
            from ldpc_b200 import SoftInfoBpDecoder
            d = SoftInfoBpDecoder(H, error_channel=z["channel"], max_iter=int(z["max_iter"]),
                                  ms_scaling_factor=float(z["ms_scaling_factor"]), cutoff=float(z["cutoff"]),
                                  sigma=float(z["sigma"]))
            got = d.decode_batch(z["soft"], return_llr=True)
            assert_same_decode((got, d.converge_batch, d.iter_batch, d.log_prob_ratios_batch),
                               (z["decoding"], z["converged"], z["iters"], z["llr"]), llr_exact=True)
            assert np.array_equal(d.soft_syndrome_batch.view(np.uint64), z["soft_out"].view(np.uint64))
            continue
        if kernel in ("edge", "pair") and str(z["schedule"]) == "serial":
            continue
        dc, dv = int(np.diff(H.indptr).max()), int(np.diff(H.tocsc().indptr).max())
        if kernel == "pair" and dc > 8 and dv > 4:
            continue  # beyond the paired family's degree buckets (hamming_code(5): 16 x 5)
        kw = dict(max_iter=int(z["max_iter"]), bp_method=str(z["bp_method"]), schedule=str(z["schedule"]),
                  ms_scaling_factor=float(z["ms_scaling_factor"]))
        got = _decode_gpu(H, z["syndromes"], z["channel"], kernel=kernel, **kw)
        assert_same_decode(got, (z["decoding"], z["converged"], z["iters"], z["llr"]),
                           llr_exact=(str(z["bp_method"]) == "ms"))


def test_large_degrees_generic_path(port_oracle):
    """Row degree 40 / column degree 20 exceed the register-array buckets: the streaming family falls back to the
    reference's own two-array, two-sweep scheme (any degree).  All four method x schedule combinations."""
    import scipy.sparse as sp
    rng = np.random.default_rng(33)
    m, n = 24, 90
    dense = (rng.random((m, n)) < 0.25).astype(np.uint8)
    dense[0, :40] = 1
    dense[:20, 7] = 1
    H = sp.csr_matrix(dense)
    assert H.sum(1).max() >= 40 and H.sum(0).max() >= 20
    err = (rng.random((400, n)) < 0.03).astype(np.uint8)
    syn = codes.syndromes_of(H, err)
    for method, sched in (("ms", "parallel"), ("ps", "parallel"), ("ms", "serial"), ("ps", "serial")):
        kw = dict(max_iter=8, bp_method=method, schedule=sched, ms_scaling_factor=0.8)
        want = port_oracle.decode_batch(H, syn, 0.03, **kw)
        got = _decode_gpu(H, syn, 0.03, **kw)
        assert_same_decode(got, want, llr_exact=(method == "ms"))


def test_bposd_surface_many_failures(port_oracle):
    """BP+OSD-0 where BP fails most of the time (rotated surface code, product-sum): every non-converged row goes
    through the device-side compaction and the host elimination; compare with the oracle's BP + OSD-0 restatement."""
    H = codes.rotated_surface_code_x(7)
    syn = codes.bsc_syndromes(H, 0.08, 5000, seed=17)
    kw = dict(max_iter=12, bp_method="ps", schedule="parallel")
    d = BpOsdDecoder(H, error_rate=0.08, osd_method="osd0", **kw)
    got = d.decode_batch(syn)
    bp = port_oracle.decode_batch(H, syn, 0.08, **kw)
    want = bp[0].copy()
    bad = ~bp[1]
    assert bad.mean() > 0.3
    want[bad] = port_oracle.osd0_batch(H, syn[bad], bp[3][bad])
    assert np.array_equal(d.converge_batch, bp[1]) and np.array_equal(d.iter_batch, bp[2])
    assert np.array_equal(got, want)
    assert np.array_equal(codes.syndromes_of(H, got), syn)


@pytest.mark.parametrize("schedule", ["parallel", "serial"])
def test_large_code_n10000(port_oracle, schedule):
    """BASELINE config 5's code (n = 10^4 (3,6)-LDPC): the messages of one syndrome (240 kB) do not fit in shared
    memory, so only the streaming family can run it, with the graph tables and syndrome words in global memory."""
    H = codes.regular_ldpc(10000, 3, 6, seed=1)
    syn = np.concatenate([codes.bsc_syndromes(H, 0.05, 40, seed=7), codes.bsc_syndromes(H, 0.08, 8, seed=8)])
    kw = dict(max_iter=40, bp_method="ms", schedule=schedule, ms_scaling_factor=0.625)
    want = port_oracle.decode_batch(H, syn, 0.05, **kw)
    d = BpDecoder(H, error_rate=0.05, input_vector_type="syndrome", **kw)
    got = d.decode_batch(syn, return_llr=True)
    assert_same_decode((got, d.converge_batch, d.iter_batch, d.log_prob_ratios_batch), want, llr_exact=True)
    assert d.info()["kernel_family"] in ((1, 3) if schedule == "parallel" else (1,)) and not want[1].all()
    with pytest.raises(Exception):
        BpDecoder(H, error_rate=0.05, kernel="smem", **kw).decode_batch(syn[:2])
    if schedule == "parallel":
        # the edge-parallel family takes the code with its messages in an L2-resident scratch (also the second stage of
        # the streaming kernel's ramp-down for codes of this size)
        e = BpDecoder(H, error_rate=0.05, input_vector_type="syndrome", kernel="edge", **kw)
        got_e = e.decode_batch(syn, return_llr=True)
        assert_same_decode((got_e, e.converge_batch, e.iter_batch, e.log_prob_ratios_batch), want, llr_exact=True)
        assert e.info()["kernel_family"] == 3
