"""CPU: pin the oracle.  (1) the reference's known-answer tests, (2) bit-for-bit agreement of the plain-C
restatement with the unmodified reference C++ (oracle/_ref) on the BASELINE codes, (3) the committed golden
fixtures generated from the reference (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from ldpc_b200 import codes
from kat import KATS, kat_arrays
from util import assert_same_decode

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(oracle_obj, H, kw, inputs, kind):
    kw = dict(kw)
    channel = kw.pop("channel")
    if kind == "received_vector":
        syn = codes.syndromes_of(H, inputs)
        out = oracle_obj.decode_batch(H, syn, channel, **kw)
        return out[0] ^ inputs
    return oracle_obj.decode_batch(H, inputs, channel, **kw)[0]


@pytest.mark.parametrize("entry", KATS, ids=[k[0] for k in KATS])
def test_port_matches_reference_kats(port_oracle, entry):
    name, H, kw, inputs, expected, kind = kat_arrays(entry)
    assert np.array_equal(_run(port_oracle, H, kw, inputs, kind), expected)


@pytest.mark.parametrize("entry", KATS, ids=[k[0] for k in KATS])
def test_ref_matches_its_own_kats(ref_oracle, entry):
    name, H, kw, inputs, expected, kind = kat_arrays(entry)
    assert np.array_equal(_run(ref_oracle, H, kw, inputs, kind), expected)


CASES = [
    ("ldpc1000_ms_par", lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 200,
     dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625)),
    ("ldpc1000_ms_par_adaptive_hard", lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.09, 60,
     dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.0)),
    ("ldpc1000_ps_par", lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 80,
     dict(max_iter=50, bp_method="ps", schedule="parallel")),
    ("ldpc1000_ps_par_hard", lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.09, 40,
     dict(max_iter=50, bp_method="ps", schedule="parallel")),
    ("ldpc1000_ms_ser", lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 100,
     dict(max_iter=50, bp_method="ms", schedule="serial", ms_scaling_factor=0.625)),
    ("ldpc1000_ps_ser", lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 40,
     dict(max_iter=50, bp_method="ps", schedule="serial")),
    ("surface13_ps", lambda: codes.rotated_surface_code_x(13), 0.05, 400,
     dict(max_iter=30, bp_method="ps", schedule="parallel")),
    ("bb144_ms", lambda: codes.bivariate_bicycle_144(), 0.02, 600,
     dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625)),
    ("hamming5_ps", lambda: codes.hamming_code(5), 0.1, 100, dict(max_iter=2, bp_method="ps", schedule="parallel")),
    # SERIAL_RELATIVE (bp.hpp:469-482): the port restates libstdc++'s std::sort, ties and all
    ("ldpc1000_ms_relative", lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 60,
     dict(max_iter=50, bp_method="ms", schedule="serial_relative", ms_scaling_factor=0.625)),
    ("ldpc1000_ms_relative_adaptive_hard", lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.085, 30,
     dict(max_iter=30, bp_method="ms", schedule="serial_relative", ms_scaling_factor=0.0)),
    ("ldpc240_ps_relative", lambda: codes.regular_ldpc(240, 3, 6, seed=3), 0.06, 100,
     dict(max_iter=30, bp_method="ps", schedule="serial_relative")),
    ("surface13_ms_relative", lambda: codes.rotated_surface_code_x(13), 0.05, 200,
     dict(max_iter=20, bp_method="ms", schedule="serial_relative", ms_scaling_factor=0.625)),
    ("hamming5_ps_relative", lambda: codes.hamming_code(5), 0.1, 100,
     dict(max_iter=5, bp_method="ps", schedule="serial_relative")),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_port_bit_identical_to_reference(port_oracle, ref_oracle, case):
    name, mk, p, B, kw = case
    H = mk()
    syn = codes.bsc_syndromes(H, p, B, seed=7)
    a = port_oracle.decode_batch(H, syn, p, **kw)
    b = ref_oracle.decode_batch(H, syn, p, **kw)
    assert_same_decode(a, b, llr_exact=True)


def test_port_serial_custom_order_matches_reference(port_oracle, ref_oracle):
    H = codes.regular_ldpc(200, 3, 6, seed=2)
    order = np.random.default_rng(4).permutation(200)
    syn = codes.bsc_syndromes(H, 0.06, 120, seed=9)
    for method in ("ms", "ps"):
        kw = dict(max_iter=25, bp_method=method, schedule="serial", ms_scaling_factor=0.9,
                  serial_schedule_order=order)
        assert_same_decode(port_oracle.decode_batch(H, syn, 0.06, **kw), ref_oracle.decode_batch(H, syn, 0.06, **kw),
                           llr_exact=True)


def test_port_osd0_matches_reference(port_oracle, ref_oracle):
    """OSD-0 restatement (oracle/osd_oracle.c) against ldpc::osd::OsdDecoder on BP failures."""
    for mk, p, B, kw in ((codes.bivariate_bicycle_144, 0.03, 400,
                          dict(max_iter=20, bp_method="ms", ms_scaling_factor=0.625)),
                         (lambda: codes.rotated_surface_code_x(7), 0.08, 300, dict(max_iter=10, bp_method="ps")),
                         (lambda: codes.hamming_code(4), 0.15, 200, dict(max_iter=3, bp_method="ps"))):
        H = mk()
        syn = codes.bsc_syndromes(H, p, B, seed=13)
        dec, conv, its, llr, bpdec = ref_oracle.decode_batch(H, syn, p, osd_method=1, osd_order=0, **kw)
        bad = ~conv
        assert bad.any()
        mine = port_oracle.osd0_batch(H, syn[bad], llr[bad])
        assert np.array_equal(mine, dec[bad])
        assert np.array_equal(codes.syndromes_of(H, mine), syn[bad])


def test_received_vector_matches_reference(port_oracle, ref_oracle):
    H = codes.regular_ldpc(120, 3, 6, seed=5)
    err = codes.bsc_errors(120, 0.05, 100, seed=1)
    want = ref_oracle.decode_received(H, err, 0.05, 20, bp_method="ps")
    got = port_oracle.decode_batch(H, codes.syndromes_of(H, err), 0.05, max_iter=20, bp_method="ps")[0] ^ err
    assert np.array_equal(got, want)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))),
                         ids=lambda p: os.path.basename(p))
def test_port_matches_golden_fixture(port_oracle, path):
    """Fixtures were produced by the unmodified reference (tests/golden/make_golden.py); they travel to
    machines where /root/reference does not exist."""
    z = np.load(path, allow_pickle=False)
    import scipy.sparse as sp
    H = sp.csr_matrix((np.ones(z["rows"].size, np.uint8), (z["rows"], z["cols"])), shape=tuple(z["shape"]))
    if "kind" in z.files and str(z["kind"]) == "soft_info":  # soft_info_decode_serial fixtures (bp.hpp:547-665)
        got = port_oracle.soft_info_decode_batch(H, z["soft"], z["channel"], int(z["max_iter"]),
                                                 float(z["ms_scaling_factor"]), float(z["cutoff"]), float(z["sigma"]))
        assert_same_decode(got[:4], (z["decoding"], z["converged"], z["iters"], z["llr"]), llr_exact=True)
        assert np.array_equal(got[4].view(np.uint64), z["soft_out"].view(np.uint64))
        return
    kw = dict(max_iter=int(z["max_iter"]), bp_method=str(z["bp_method"]), schedule=str(z["schedule"]),
              ms_scaling_factor=float(z["ms_scaling_factor"]))
    got = port_oracle.decode_batch(H, z["syndromes"], z["channel"], **kw)
    assert_same_decode(got, (z["decoding"], z["converged"], z["iters"], z["llr"]), llr_exact=True)


def _levelise(H, order):
    """Python mirror of build_serial_batches / build_smem_plan (ldpc_b200/csrc/bp_plan.cpp): level(q) = 1 + max level
    of the earlier schedule positions whose bit shares a check; stable sort by level."""
    import scipy.sparse as sp
    Hc = sp.csc_matrix(H)
    last = np.zeros(H.shape[0], dtype=int)
    level = []
    for j in order:
        rows = Hc.indices[Hc.indptr[j]:Hc.indptr[j + 1]]
        lv = (last[rows].max() if rows.size else 0) + 1
        last[rows] = lv
        level.append(lv)
    level = np.asarray(level)
    return [int(order[q]) for q in np.argsort(level, kind="stable")], level


@pytest.mark.parametrize("method", ["ms", "ps"])
def test_levelised_serial_schedule_is_equivalent(port_oracle, ref_oracle, method):
    """The GPU serial kernels process the schedule level by level (bits that share no check commute).  Pin that claim
    on the CPU with the reference itself: decoding in the levelised order gives bit-identical results (decisions,
    iterations, every LLR) to decoding in the original order, for the default and for a random custom order."""
    H = codes.regular_ldpc(240, 3, 6, seed=3)
    syn = np.concatenate([codes.bsc_syndromes(H, 0.05, 60, seed=1), codes.bsc_syndromes(H, 0.10, 20, seed=2)])
    kw = dict(max_iter=30, bp_method=method, schedule="serial", ms_scaling_factor=0.625)
    for order in (np.arange(240), np.random.default_rng(5).permutation(240)):
        lev_order, level = _levelise(H, order)
        assert sorted(lev_order) == sorted(int(x) for x in order) and level.max() < 240
        for orc in (port_oracle, ref_oracle):
            a = orc.decode_batch(H, syn, 0.05, serial_schedule_order=np.asarray(order), **kw)
            b = orc.decode_batch(H, syn, 0.05, serial_schedule_order=np.asarray(lev_order), **kw)
            assert_same_decode(a, b, llr_exact=True)
        assert not a[1].all() and a[1].any()


def _soft_inputs(H, p, B, sigma, seed):
    rng = np.random.default_rng(seed)
    err = (rng.random((B, H.shape[1])) < p).astype(np.uint8)
    syn = codes.syndromes_of(H, err)
    return (1 - 2.0 * syn) + rng.normal(0, sigma, size=syn.shape)


@pytest.mark.parametrize("mk,p", [(lambda: codes.rep_code(7), 0.1), (lambda: codes.regular_ldpc(240, 3, 6, seed=3), 0.05),
                                  (lambda: codes.rotated_surface_code_x(7), 0.05), (codes.bivariate_bicycle_144, 0.02)],
                         ids=["rep7", "ldpc240", "surface7", "bb144"])
def test_port_soft_info_bit_identical_to_reference(port_oracle, ref_oracle, mk, p):
    """soft_info_decode_serial (bp.hpp:547-665): restatement vs the reference, incl. the posterior LLRs and the soft
    syndrome the virtual check updates leave behind, for cutoffs that never / sometimes / mostly trigger them."""
    H = mk()
    for sigma, cutoff, ms in ((0.6, np.inf, 1.0), (0.8, 3.0, 0.625), (1.2, 10.0, 0.9), (0.3, 1.0, 1.0)):
        soft = _soft_inputs(H, p, 120, sigma, seed=int(sigma * 10))
        a = port_oracle.soft_info_decode_batch(H, soft, p, 20, ms, cutoff, sigma)
        b = ref_oracle.soft_info_decode_batch(H, soft, p, 20, ms, cutoff, sigma)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert np.array_equal(a[3].view(np.uint64), b[3].view(np.uint64))
        assert np.array_equal(a[4].view(np.uint64), b[4].view(np.uint64))
