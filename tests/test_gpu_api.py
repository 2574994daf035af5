"""GPU: the reference's known-answer tests, replayed through the drop-in Python API."""
import numpy as np
import pytest

from ldpc_b200 import BpDecoder, BpOsdDecoder, codes
from kat import KATS, kat_arrays

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("entry", KATS, ids=[k[0] for k in KATS])
def test_reference_kats_through_the_shim(entry):
    """The reference's known-answer tests (tests/kat.py) decoded one vector at a time with .decode()."""
    name, H, kw, inputs, expected, kind = kat_arrays(entry)
    channel = kw.pop("channel")
    chan = dict(error_rate=float(channel)) if np.isscalar(channel) else dict(error_channel=channel)
    d = BpDecoder(H, input_vector_type=kind, **chan, **kw)
    for vec, want in zip(inputs, expected):
        assert np.array_equal(d.decode(vec), want), (name, vec)
    # and as one batch
    assert np.array_equal(d.decode_batch(inputs), expected)


def test_zero_shortcut_and_dtype_echo():
    H = codes.rep_code(5)
    d = BpDecoder(H, error_rate=0.1, max_iter=5, bp_method="ms")
    out = d.decode(np.array([0, 1, 1, 0], dtype=np.int64))
    assert out.dtype == np.int64
    it = d.iter
    z = d.decode(np.zeros(4, dtype=np.int8))
    assert z.dtype == np.int8 and not z.any() and d.converge and d.iter == it  # iter untouched (:679-681)


def test_bposd_single_decode():
    H = codes.hamming_code(3)
    d = BpOsdDecoder(H, error_rate=0.1, max_iter=3, bp_method="ps", osd_method="osd0")
    for s in range(1, 8):
        syn = np.array([(s >> 2) & 1, (s >> 1) & 1, s & 1], dtype=np.uint8)
        out = d.decode(syn)
        assert np.array_equal(codes.syndromes_of(H, out[None, :])[0], syn)


def test_info_reports_native_kernel():
    H = codes.rep_code(5)
    d = BpDecoder(H, error_rate=0.1, max_iter=5, bp_method="ms")
    d.decode(np.array([1, 0, 0, 0]))
    inf = d.info()
    assert inf["launches"] >= 2 and inf["kernel_family"] in (1, 2, 3, 4) and inf["grid"] >= 1


@pytest.mark.parametrize("kernel", ["stream", "smem"])
def test_ragged_and_degenerate_batches(kernel):
    """Empty batch, one syndrome, a batch that is no multiple of the warp / group size, all-zero syndromes, and a
    batch with repeated rows: results must be per-row and independent of batch composition."""
    H = codes.regular_ldpc(120, 3, 6, seed=5)
    d = BpDecoder(H, error_rate=0.05, max_iter=20, bp_method="ms", ms_scaling_factor=0.625,
                  input_vector_type="syndrome", kernel=kernel)
    syn = codes.bsc_syndromes(H, 0.06, 77, seed=2)
    full = d.decode_batch(syn)
    its = d.iter_batch.copy()
    assert d.decode_batch(syn[:0]).shape == (0, 120)
    one = d.decode_batch(syn[5:6])
    assert np.array_equal(one[0], full[5]) and d.iter_batch[0] == its[5]
    for lo, hi in ((0, 33), (10, 77), (31, 32)):
        assert np.array_equal(d.decode_batch(syn[lo:hi]), full[lo:hi])
    z = d.decode_batch(np.zeros((3, 60), np.uint8))
    assert not z.any() and d.converge_batch.all() and (d.iter_batch == 1).all()  # C++ semantics for zero syndromes
    rep = np.repeat(syn[7:8], 40, axis=0)
    out = d.decode_batch(rep)
    assert (out == full[7]).all()
    # int64 in, int64 out (dtype echo of the batch API)
    assert d.decode_batch(syn.astype(np.int64)).dtype == np.int64


def test_multi_gpu_decoder_single_device():
    from ldpc_b200.parallel import MultiGpuBpDecoder
    H = codes.regular_ldpc(120, 3, 6, seed=5)
    syn = codes.bsc_syndromes(H, 0.06, 500, seed=2)
    kw = dict(error_rate=0.05, max_iter=20, bp_method="ms", ms_scaling_factor=0.625, input_vector_type="syndrome")
    want = BpDecoder(H, **kw).decode_batch(syn)
    multi = MultiGpuBpDecoder(H, devices=[0, 0], **kw)  # two handles on the same device: exercises the threaded split
    assert np.array_equal(multi.decode_batch(syn), want)


def test_v1_vs_v2_same_logical_error_rate():
    """reference python_test/test_bp_decoder.py:238-263: rep_code(100), min-sum, 10 iterations, 1000 seeded runs --
    the v1-syntax class and the v2 class give the same logical error rate (here both through decode_batch)."""
    import warnings
    from ldpc_b200 import MonteCarloBscSimulation, bp_decoder
    H = codes.rep_code(100)
    bpd = BpDecoder(H, error_rate=0.20, bp_method="ms", schedule="parallel", ms_scaling_factor=1.0, max_iter=10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bpd_v1 = bp_decoder(H, error_rate=0.20, bp_method="ms", ms_scaling_factor=1.0, max_iter=10)
    out1 = MonteCarloBscSimulation(H, error_rate=0.20, Decoder=bpd, target_run_count=1000, seed=42).run()
    out2 = MonteCarloBscSimulation(H, error_rate=0.20, Decoder=bpd_v1, target_run_count=1000, seed=42).run()
    assert out1["logical_error_rate"] == out2["logical_error_rate"]
    assert 0 < out1["fail_count"] < 1000
