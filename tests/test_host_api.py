"""CPU: host-side logic of the drop-in API (no GPU needed) -- constructor defaults, aliases, exception
types, properties (reference python_test/test_bp_decoder.py:94-172, test_bp_decoder_input.py,
test_scipy_helpers.py), the C-ABI library's exports, and the host OSD-0 against the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

import ldpc_b200
from ldpc_b200 import BpDecoder, BpOsdDecoder, _capi, codes, io_test
from ldpc_b200.helpers import convert_to_binary_sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _capi.lib()
    header = open(os.path.join(ROOT, "include", "bp_b200.h")).read()
    declared = set(re.findall(r"\b(bpb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bp_b200.h but not exported"
    assert set(_capi.EXPORTS) <= declared
    assert b"sm_100a" in lib.bpb_version()


def test_cython_shim_is_the_binding():
    """The host side is a Cython shim over the C-ABI (like the reference's); ctypes only binds the same library when
    the extension has not been built."""
    from ldpc_b200 import _native
    if _native.BINDING != "cython":
        pytest.skip("ldpc_b200/_bp_shim not built (run __graft_entry__.build())")
    from ldpc_b200 import _bp_shim
    rows = np.array([0, 0, 1, 1], np.int32)
    cols = np.array([0, 1, 1, 2], np.int32)
    h = _bp_shim.NativeHandle(2, 3, rows, cols, -1)  # host-only handle
    h.set_channel(np.full(3, 0.1))
    inf = h.info()
    assert inf["m"] == 2 and inf["n"] == 3 and inf["nnz"] == 4 and inf["smem_family_available"] == 1
    with pytest.raises(_bp_shim.NativeError):
        h.decode_batch(0, np.ones((1, 2), np.uint8), np.zeros((1, 3), np.uint8), np.zeros(1, np.uint8),
                       np.zeros(1, np.int32), None)


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU, creating a decoder must fail loudly (never silently decode on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = BpDecoder(codes.rep_code(3), error_rate=0.1)
    with pytest.raises(_capi.BpbError):
        d.decode(np.array([1, 1]))
    # a host-only handle refuses to decode
    lib = _capi.lib()
    h = C.c_void_p()
    rows = np.array([0, 0, 1, 1], np.int32)
    cols = np.array([0, 1, 1, 2], np.int32)
    assert lib.bpb_create(2, 3, 4, rows.ctypes.data_as(_capi._i32p), cols.ctypes.data_as(_capi._i32p), -1,
                          C.byref(h)) == 0
    ch = np.full(3, 0.1)
    assert lib.bpb_set_channel(h, ch.ctypes.data_as(_capi._f64p), 3) == 0
    syn = np.array([[1, 1]], np.uint8)
    out = np.zeros((1, 3), np.uint8)
    rc = lib.bpb_decode_batch(h, 0, _capi.host_ptr(syn), 1, _capi.host_ptr(out), None, None, None)
    assert rc != 0 and b"no CPU fallback" in lib.bpb_last_error(h)
    lib.bpb_destroy(h)


def test_defaults_match_reference():
    # python_test/test_bp_decoder.py:94-136
    H = codes.rep_code(5)
    d = BpDecoder(H, error_rate=0.1)
    assert d.max_iter == 5  # 0 -> n
    assert d.bp_method == "product_sum"
    assert d.ms_scaling_factor == 1.0
    assert d.schedule == "parallel"
    assert d.omp_thread_count == 1
    assert d.random_schedule_seed == 0
    assert np.array_equal(d.serial_schedule_order, np.arange(5))
    assert d.input_vector_type == "auto"
    assert d.check_count == 4 and d.bit_count == 5
    assert np.allclose(d.error_rate, 0.1) and np.allclose(d.channel_probs, 0.1)
    assert not d.random_serial_schedule


def test_constructor_echo_and_aliases():
    H = codes.rep_code(5)
    d = BpDecoder(H, error_channel=[0.1, 0.2, 0.3, 0.2, 0.1], max_iter=10, bp_method="ms", ms_scaling_factor=0.5,
                  schedule="s", serial_schedule_order=[4, 3, 2, 1, 0])
    assert d.bp_method == "minimum_sum" and d.schedule == "serial" and d.max_iter == 10
    assert d.ms_scaling_factor == 0.5
    assert np.array_equal(d.serial_schedule_order, [4, 3, 2, 1, 0])
    assert np.allclose(d.error_channel, [0.1, 0.2, 0.3, 0.2, 0.1])
    for alias in ["prod_sum", "product_sum", "ps", "0", "prod sum", 0]:
        assert BpDecoder(H, error_rate=0.1, bp_method=alias).bp_method == "product_sum"
    for alias in ["min_sum", "minimum_sum", "ms", "1", "minimum sum", "min sum", 1]:
        assert BpDecoder(H, error_rate=0.1, bp_method=alias).bp_method == "minimum_sum"
    # schedule alias quirk: '0' -> parallel, '1' -> serial (_bp_decoder.pyx:426-431)
    assert BpDecoder(H, error_rate=0.1, schedule=0).schedule == "parallel"
    assert BpDecoder(H, error_rate=0.1, schedule=1).schedule == "serial"
    assert BpDecoder(H, error_rate=0.1, schedule="sr").schedule == "serial_relative"
    # legacy channel_probs kwarg
    assert np.allclose(BpDecoder(H, channel_probs=[0.3] * 5).error_channel, 0.3)
    d.update_channel_probs([0.05] * 5)
    assert np.allclose(d.channel_probs, 0.05)


def test_error_types_match_reference():
    # python_test/test_bp_decoder.py:138-172
    H = codes.rep_code(5)
    with pytest.raises(TypeError):
        BpDecoder("not a matrix", error_rate=0.1)
    with pytest.raises(ValueError):
        BpDecoder(H)  # no channel
    with pytest.raises(ValueError):
        BpDecoder(H, error_rate=1)  # must be float
    with pytest.raises(ValueError):
        BpDecoder(H, error_channel=[0.1, 0.2])
    with pytest.raises(ValueError):
        BpDecoder(H, error_rate=0.1, max_iter=-1)
    with pytest.raises(ValueError):
        BpDecoder(H, error_rate=0.1, max_iter=1.5)
    with pytest.raises(ValueError):
        BpDecoder(H, error_rate=0.1, bp_method="nope")
    with pytest.raises(ValueError):
        BpDecoder(H, error_rate=0.1, schedule="nope")
    with pytest.raises(TypeError):
        BpDecoder(H, error_rate=0.1, ms_scaling_factor="1.0")
    with pytest.raises(TypeError):
        BpDecoder(H, error_rate=0.1, omp_thread_count=0)
    with pytest.raises(Exception):
        BpDecoder(H, error_rate=0.1, serial_schedule_order=[0, 1])
    with pytest.raises(ValueError):
        BpDecoder(H, error_rate=0.1, serial_schedule_order=[0, 1, 2, 3, 7])
    with pytest.raises(ValueError):
        BpDecoder(H, error_rate=0.1, not_a_parameter=3)
    with pytest.raises(ValueError):
        BpDecoder(codes.ring_code(5), error_rate=0.1)  # square H needs an explicit input_vector_type
    assert BpDecoder(codes.ring_code(5), error_rate=0.1, input_vector_type="syndrome").input_vector_type == "syndrome"
    d = BpDecoder(H, error_rate=0.1, input_vector_type="syndrome")
    with pytest.raises(ValueError):
        d.decode(np.zeros(5, np.uint8))  # wrong length for syndrome input


def test_zero_syndrome_shortcut_needs_no_gpu():
    # _bp_decoder.pyx:676-681: all-zero input returns zeros of the input dtype and sets converge
    d = BpDecoder(codes.rep_code(5), error_rate=0.1)
    out = d.decode(np.zeros(4, dtype=np.int32))
    assert out.dtype == np.int32 and out.shape == (5,) and not out.any() and d.converge
    o = BpOsdDecoder(codes.rep_code(5), error_rate=0.1)
    out = o.decode(np.zeros(4, dtype=np.uint8))
    assert not out.any() and o.converge


def test_bposd_properties():
    H = codes.hamming_code(3)
    d = BpOsdDecoder(H, error_rate=0.1)
    assert d.osd_method == "OSD_0" and d.osd_order == 0 and d.input_vector_type == "syndrome"
    assert BpOsdDecoder(H, error_rate=0.1, osd_method="osd_cs", osd_order=3).osd_method == "OSD_CS"
    assert BpOsdDecoder(H, error_rate=0.1, osd_method="e", osd_order=2).osd_method == "OSD_E"
    with pytest.raises(ValueError):
        BpOsdDecoder(H, error_rate=0.1, osd_method="bogus")
    with pytest.raises(ValueError):
        BpOsdDecoder(H, error_rate=0.1, osd_method="osd0", osd_order=2)
    with pytest.raises(ValueError):
        BpOsdDecoder(H, error_rate=0.1, osd_method="osd_cs", osd_order=-1)
    with pytest.raises(ValueError):
        d.decode(np.zeros(5, np.uint8))


def test_input_matrix_formats_round_trip():
    # python_test/test_bp_decoder.py:10-29 and test_bp_decoder_input.py
    dense = codes.hamming_code(3).toarray()
    for mat in (dense, sp.csr_matrix(dense), sp.csc_matrix(dense), dense.astype(int), dense.astype(float)):
        out = io_test(mat)
        assert np.array_equal(out.toarray(), dense)
        BpDecoder(mat, error_rate=0.1)
        BpOsdDecoder(mat, error_rate=0.1)


def test_convert_to_binary_sparse_contract():
    # python_test/test_scipy_helpers.py
    with pytest.raises(TypeError):
        convert_to_binary_sparse([[1, 0], [0, 1]])
    with pytest.raises(TypeError):
        convert_to_binary_sparse(np.array([[1, 0]], dtype=np.complex64))
    with pytest.raises(ValueError):
        convert_to_binary_sparse(np.array([[2, 0], [0, 1]]))
    out = convert_to_binary_sparse(np.array([[1, 0], [0, 1]], dtype=np.uint8))
    assert sp.issparse(out) and out.nnz == 2
    s = sp.csr_matrix(np.array([[1.0, 0.0], [1.0, 1.0]]))
    assert convert_to_binary_sparse(s).dtype == np.uint8


def test_code_constructors():
    assert np.array_equal(codes.rep_code(3).toarray(), [[1, 1, 0], [0, 1, 1]])
    assert np.array_equal(codes.ring_code(3).toarray(), [[1, 1, 0], [0, 1, 1], [1, 0, 1]])
    assert np.array_equal(codes.hamming_code(3).toarray(), [[0, 0, 0, 1, 1, 1, 1], [0, 1, 1, 0, 0, 1, 1],
                                                             [1, 0, 1, 0, 1, 0, 1]])
    H = codes.regular_ldpc(1000, 3, 6, seed=1)
    assert H.shape == (500, 1000) and set(np.asarray(H.sum(0)).ravel()) == {3}
    assert set(np.asarray(H.sum(1)).ravel()) == {6}
    S = codes.rotated_surface_code_x(13)
    assert S.shape == (84, 169) and S.nnz == 312
    Bm = codes.bivariate_bicycle_144()
    assert Bm.shape == (72, 144) and Bm.nnz == 432
    # the X checks of a CSS code commute with the Z checks built the same way: B^T part; sanity: rank deficiency 6
    syn = codes.bsc_syndromes(H, 0.05, 64, seed=3)
    assert syn.shape == (64, 500) and syn.dtype == np.uint8


def _host_only_handle(H):
    lib = _capi.lib()
    coo = sp.coo_matrix(H)
    rows = np.ascontiguousarray(coo.row, np.int32)
    cols = np.ascontiguousarray(coo.col, np.int32)
    h = C.c_void_p()
    rc = lib.bpb_create(H.shape[0], H.shape[1], rows.size, rows.ctypes.data_as(_capi._i32p),
                        cols.ctypes.data_as(_capi._i32p), -1, C.byref(h))
    assert rc == 0
    return lib, h


@pytest.mark.parametrize("threads", [1, 4])
def test_host_osd0_matches_oracle_and_reference(port_oracle, ref_oracle, threads):
    """bpb_osd0_host (ldpc_b200/csrc/osd_host.cpp, bit-packed column elimination) against the reference's
    OsdDecoder on the LLRs the reference's BP produced for BP failures."""
    for mk, p, B, kw in ((codes.bivariate_bicycle_144, 0.03, 500,
                          dict(max_iter=20, bp_method="ms", ms_scaling_factor=0.625)),
                         (lambda: codes.rotated_surface_code_x(9), 0.08, 300, dict(max_iter=10, bp_method="ps")),
                         (lambda: codes.regular_ldpc(240, 3, 6, seed=3), 0.11, 100,
                          dict(max_iter=8, bp_method="ms", ms_scaling_factor=0.625))):
        H = mk()
        syn = codes.bsc_syndromes(H, p, B, seed=13)
        dec, conv, its, llr, bpdec = ref_oracle.decode_batch(H, syn, p, osd_method=1, osd_order=0, **kw)
        assert (~conv).any()
        lib, h = _host_only_handle(H)
        out = bpdec.copy()
        rc = lib.bpb_osd0_host(h, _capi.host_ptr(syn), _capi.host_ptr(llr),
                               _capi.host_ptr(conv.astype(np.uint8)), B, _capi.host_ptr(out), threads)
        assert rc == 0
        lib.bpb_destroy(h)
        assert np.array_equal(out, dec)
        assert np.array_equal(out[~conv], port_oracle.osd0_batch(H, syn[~conv], llr[~conv]))


def test_osd0_random_syndromes_bb144(ref_oracle):
    """Random syndromes on a rank-deficient H (bb144: rank 66 of 72 rows) with the reference's own LLRs.  Rows whose
    syndrome lies in the image of H must equal the reference's OsdDecoder output; for the others the reference's
    answer depends on its linked-list pivot-row order (documented divergence, osd_host.cpp header): here the result
    solves the equations of the pivot rows it picked and the library counts them (`osd_host_inconsistent`)."""
    H = codes.bivariate_bicycle_144()
    m, n = H.shape
    rng = np.random.default_rng(99)
    B = 400
    syn = rng.integers(0, 2, size=(B, m)).astype(np.uint8)
    syn[: B // 2] = codes.bsc_syndromes(H, 0.05, B // 2, seed=3)  # half of them are genuine s = H e
    kw = dict(max_iter=15, bp_method="ms", ms_scaling_factor=0.625)
    dec, conv, its, llr, bpdec = ref_oracle.decode_batch(H, syn, 0.05, osd_method=1, osd_order=0, **kw)
    lib, h = _host_only_handle(H)
    out = bpdec.copy()
    rc = lib.bpb_osd0_host(h, _capi.host_ptr(syn), _capi.host_ptr(llr), _capi.host_ptr(conv.astype(np.uint8)), B,
                           _capi.host_ptr(out), 2)
    assert rc == 0
    inf = _capi.BpbInfo()
    assert lib.bpb_get_info(h, C.byref(inf)) == 0
    lib.bpb_destroy(h)
    in_image = (codes.syndromes_of(H, dec) == syn).all(axis=1)  # the reference solved it <=> it is solvable
    assert in_image[: B // 2].all() and 0 < (~in_image).sum()
    assert np.array_equal(out[in_image], dec[in_image])
    assert inf.osd_host_inconsistent == int((~in_image & ~conv).sum())
    # outside the image: still a deterministic answer supported on at most rank(H) columns
    assert out[~in_image].sum(axis=1).max() <= 66


def test_osd_column_order_is_the_libc_qsort_merge_tree(tmp_path):
    """ldpc_b200/csrc/osd_order.h (what the device OSD-0 kernel walks) against the live libc qsort with the reference's
    record layout and comparator (sort.hpp:27-62): ties, +-0, +-inf and NaN keys."""
    import subprocess
    exe = str(tmp_path / "osd_order_check")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "osd_order_check.c")
    subprocess.run(["gcc", "-O2", "-o", exe, src], check=True)
    res = subprocess.run([exe, "3000"], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.startswith("0 orders differ"), res.stdout


def test_stl_sort_restatement_matches_libstdcxx(tmp_path):
    """ldpc_b200/csrc/stl_sort.h (the std::sort restatement behind the SERIAL_RELATIVE schedule, bp.hpp:469-482)
    against the real std::sort with the reference's comparator: random, tie-heavy, all-equal, NaN, sorted, reversed
    and median-of-3-killer inputs, also re-sorting a previous result with new keys."""
    exe = str(tmp_path / "stl_sort_check")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "stl_sort_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    res = subprocess.run([exe, "3000"], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.startswith("0 permutations differ"), res.stdout


@pytest.mark.parametrize("mk", [lambda: codes.regular_ldpc(1000, 3, 6, seed=1), codes.bivariate_bicycle_144,
                                lambda: codes.rotated_surface_code_x(13), lambda: codes.hamming_code(5),
                                lambda: codes.rep_code(40)], ids=["ldpc1000", "bb144", "surface13", "hamming5", "rep40"])
def test_shared_memory_placement_is_conflict_free(mk):
    """The on-chip family places messages by 16-colouring the (half-warp, slot) incidence graph (Koenig): no half-warp
    of the check pass or of the bit pass may hit a bank pair twice."""
    H = mk()
    lib, h = _host_only_handle(H)
    ch = np.full(H.shape[1], 0.05)
    assert lib.bpb_set_channel(h, ch.ctypes.data_as(_capi._f64p), H.shape[1]) == 0
    inf = _capi.BpbInfo()
    assert lib.bpb_get_info(h, C.byref(inf)) == 0
    lib.bpb_destroy(h)
    assert inf.smem_family_available == 1
    assert inf.smem_bank_multiplicity == 1
    assert inf.smem_bytes_per_syndrome >= 8 * H.nnz
    # the paired family places double2 slots by an 8-colouring over quarter-warps (16-byte accesses); hamming_code(5)
    # (row degree 16 with column degree 5) is beyond its degree buckets
    dc, dv = int(np.diff(H.tocsr().indptr).max()), int(np.diff(H.tocsc().indptr).max())
    if not (dc > 8 and dv > 4):
        assert inf.pair_family_available == 1
        assert inf.pair_bank_multiplicity == 1


def test_monte_carlo_driver_matches_reference_loop(port_oracle):
    """MonteCarloBscSimulation (batched) against the reference's one-run-at-a-time loop (mcs.py:124-149) with the
    same seed; the decoder here is a CPU stand-in with a decode_batch method so that the host logic is covered
    without a GPU."""
    from ldpc_b200 import MonteCarloBscSimulation
    H = codes.regular_ldpc(120, 3, 6, seed=5)
    kw = dict(max_iter=10, bp_method="ms", ms_scaling_factor=0.625)

    class OracleDecoder:
        def decode_batch(self, syn):
            return port_oracle.decode_batch(H, syn, 0.08, want_llr=False, **kw)[0]

    sim = MonteCarloBscSimulation(H, error_rate=0.08, Decoder=OracleDecoder(), target_run_count=300, seed=11,
                                  batch_size=128, tqdm_disable=True)
    out = sim.run()
    # reference loop
    np.random.seed(11)
    Hs = sp.csr_matrix(H, dtype=np.int32)
    fails = 0
    for _ in range(300):
        err = np.random.binomial(1, 0.08, 120).astype(np.uint8)
        syn = (Hs @ err % 2).astype(np.uint8)
        dec = port_oracle.decode_batch(H, syn[None, :], 0.08, want_llr=False, **kw)[0][0]
        fails += not np.array_equal(dec, err)
    assert out["run_count"] == 300 and out["fail_count"] == fails and 0 < fails < 300
    assert abs(out["logical_error_rate"] - fails / 300) < 1e-15
    with pytest.raises(ValueError):
        MonteCarloBscSimulation(H, error_rate=1, Decoder=OracleDecoder())
    with pytest.raises(ValueError):
        MonteCarloBscSimulation(H, error_rate=0.1, Decoder=None)


def test_legacy_v1_constructors():
    # python_test/test_bp_decoder_input.py: the v1 classes accept ndarray / csr / csc and v1 argument names
    from ldpc_b200 import bp_decoder, bposd_decoder
    dense = codes.hamming_code(3).toarray()
    for mat in (dense, sp.csr_matrix(dense), sp.csc_matrix(dense)):
        with pytest.warns(UserWarning):
            d = bp_decoder(mat, error_rate=0.1, bp_method="min_sum", ms_scaling_factor=1, max_iter=4)
        assert d.bp_method == "minimum_sum" and d.max_iter == 4 and d.ms_scaling_factor == 1.0
        with pytest.warns(UserWarning):
            o = bposd_decoder(mat, channel_probs=[0.1] * 7, bp_method="ps", osd_method="osd_0")
        assert o.osd_method == "OSD_0" and np.allclose(o.channel_probs, 0.1)
    with pytest.warns(UserWarning), pytest.raises(ValueError):
        bp_decoder(dense, channel_probs=[0.1, 0.2])
    with pytest.warns(UserWarning), pytest.raises(ValueError):
        bp_decoder(dense, error_rate=0)
    with pytest.warns(UserWarning), pytest.raises(ValueError):
        bposd_decoder(dense, error_rate=0.1, osd_method="bogus")
