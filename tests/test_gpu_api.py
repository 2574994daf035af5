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
    assert inf["launches"] >= 2 and inf["kernel_family"] in (1, 2) and inf["grid"] >= 1
