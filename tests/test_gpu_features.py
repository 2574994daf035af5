"""GPU: the round-2 additions behind the C-ABI -- OSD-0 on the device, the in-library multi-device split, on-device
Monte-Carlo sampling/scoring, pageable-input staging, stream ordering."""
import ctypes as C

import numpy as np
import pytest

from ldpc_b200 import BpDecoder, BpOsdDecoder, MonteCarloBscSimulation, _capi, codes
from ldpc_b200.parallel import MultiGpuBpDecoder
from util import assert_same_decode, philox_bsc_errors

pytestmark = pytest.mark.gpu


def _bposd(H, p, where, **kw):
    chan = dict(error_rate=float(p)) if np.isscalar(p) else dict(error_channel=p)
    return BpOsdDecoder(H, osd_method="osd0", osd_location=where, **chan, **kw)


@pytest.mark.parametrize("case", ["surface13_ps", "surface7_ps_nan", "bb144_ms", "ldpc1000_ms", "hamming_ps", "irregular"])
def test_device_osd0_equals_host_osd0(case, port_oracle):
    """OSD-0 on the device (osd_device.cu) against the host elimination (osd_host.cpp, pinned to the reference's
    OsdDecoder on the CPU box) and against the oracle's BP + OSD-0, on codes where BP fails often; product-sum cases
    include rows whose posterior LLRs contain +-inf and NaN (column order = the libc merge tree)."""
    if case == "surface13_ps":
        H, p, B, kw = codes.rotated_surface_code_x(13), 0.05, 20000, dict(max_iter=30, bp_method="ps")
    elif case == "surface7_ps_nan":
        # a channel that declares some bits error-free (p = 0: prior +inf) while the sampled errors flip them anyway:
        # BP cannot converge and the posteriors of the failures contain +-inf and NaN (inf - inf)
        H, B, kw = codes.rotated_surface_code_x(7), 20000, dict(max_iter=40, bp_method="ps")
        p = np.full(H.shape[1], 0.1)
        p[::6] = 0.0
        p[1::6] = 1.0  # and some certainly flipped (prior -inf): neighbours receive +inf and -inf
    elif case == "bb144_ms":
        H, p, B, kw = codes.bivariate_bicycle_144(), 0.03, 20000, dict(max_iter=50, bp_method="ms", ms_scaling_factor=0.625)
    elif case == "ldpc1000_ms":
        H, p, B, kw = codes.regular_ldpc(1000, 3, 6, seed=1), 0.085, 600, dict(max_iter=20, bp_method="ms", ms_scaling_factor=0.625)
    elif case == "hamming_ps":
        H, p, B, kw = codes.hamming_code(5), 0.1, 4000, dict(max_iter=2, bp_method="ps")
    else:
        import scipy.sparse as sp
        rng = np.random.default_rng(5)
        dense = (rng.random((40, 75)) < 0.09).astype(np.uint8)
        dense[:, dense.sum(0) == 0] |= (rng.random((40, 1)) < 0.1).astype(np.uint8)
        dense[np.arange(40), np.arange(40)] = 1
        H, p, B, kw = sp.csr_matrix(dense), 0.08, 5000, dict(max_iter=6, bp_method="ms", ms_scaling_factor=0.9)
    if case == "hamming_ps":
        syn = codes.syndromes_of(H, codes.bsc_errors(H.shape[1], p, B, seed=3))
    elif case == "surface7_ps_nan":
        syn = codes.bsc_syndromes(H, 0.1, B, seed=21)
    else:
        syn = codes.bsc_syndromes(H, p, B, seed=21)
    dev = _bposd(H, p, "device", **kw)
    host = _bposd(H, p, "host", **kw)
    got = dev.decode_batch(syn, return_bp_decoding=True)
    want = host.decode_batch(syn)
    assert dev.info()["osd_device_available"] == 1
    assert dev.info()["osd_device_solved"] == int((~dev.converge_batch).sum()) > 0
    assert host.info()["osd_host_solved"] == int((~host.converge_batch).sum())
    assert np.array_equal(dev.converge_batch, host.converge_batch) and np.array_equal(dev.iter_batch, host.iter_batch)
    rows = (got != want).any(axis=1)
    assert not rows.any(), f"{int(rows.sum())} rows differ between device and host OSD-0"
    assert np.array_equal(codes.syndromes_of(H, got), syn)  # every syndrome is in the image here (s = H e)
    # raw BP output kept on request
    chan = dict(error_rate=float(p)) if np.isscalar(p) else dict(error_channel=p)
    bp = BpDecoder(H, input_vector_type="syndrome", **chan, **kw).decode_batch(syn)
    assert np.array_equal(dev.bp_decoding_batch, bp)
    # and the oracle's BP + OSD-0 restatement on a slice
    sl = slice(0, min(B, 3000))
    r = port_oracle.decode_batch(H, syn[sl], p, **{"ms_scaling_factor": 1.0, **kw})
    w = r[0].copy()
    if (~r[1]).any():
        w[~r[1]] = port_oracle.osd0_batch(H, syn[sl][~r[1]], r[3][~r[1]])
    assert np.array_equal(got[sl], w)
    if case == "surface7_ps_nan":
        bad_llr = r[3][~r[1]]
        assert np.isnan(bad_llr).any() and np.isinf(bad_llr).any(), "this case is meant to cover non-finite LLRs"


def test_device_osd_unavailable_for_large_code():
    H = codes.regular_ldpc(10000, 3, 6, seed=1)
    d = _bposd(H, 0.09, "device", max_iter=3, bp_method="ms")
    assert d.info()["osd_device_available"] == 0
    with pytest.raises(_capi.BpbError):
        d.decode_batch(codes.bsc_syndromes(H, 0.09, 4, seed=1))
    a = _bposd(H, 0.09, "auto", max_iter=3, bp_method="ms")  # falls back to the host elimination
    syn = codes.bsc_syndromes(H, 0.09, 6, seed=1)
    out = a.decode_batch(syn)
    assert np.array_equal(codes.syndromes_of(H, out), syn) and a.info()["osd_host_solved"] > 0


def test_bposd_device_pointer_api():
    """bpb_bposd_decode_batch_device: BP + OSD-0 enqueued on a caller stream, nothing crosses PCIe."""
    import torch
    H = codes.rotated_surface_code_x(9)
    B = 30000
    syn = codes.bsc_syndromes(H, 0.06, B, seed=2)
    d = _bposd(H, 0.06, "auto", max_iter=20, bp_method="ps")
    want = d.decode_batch(syn)
    h, L = d._ensure_handle(), _capi.lib()
    dev = torch.device("cuda", 0)
    t_syn = torch.from_numpy(syn).to(dev)
    t_dec = torch.zeros((B, H.shape[1]), dtype=torch.uint8, device=dev)
    t_bp = torch.zeros_like(t_dec)
    t_conv = torch.zeros(B, dtype=torch.uint8, device=dev)
    t_its = torch.zeros(B, dtype=torch.int32, device=dev)
    for st in (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)):  # alternate streams
        t_dec.zero_()
        torch.cuda.synchronize()
        rc = L.bpb_bposd_decode_batch_device(h, C.c_void_p(t_syn.data_ptr()), B, C.c_void_p(t_dec.data_ptr()),
                                             C.c_void_p(t_conv.data_ptr()), C.c_void_p(t_its.data_ptr()),
                                             C.c_void_p(t_bp.data_ptr()), C.c_void_p(st.cuda_stream))
        _capi.check(h, rc)
        st.synchronize()
        assert np.array_equal(t_dec.cpu().numpy(), want)
        assert np.array_equal(t_conv.cpu().numpy().astype(bool), d.converge_batch)
    assert (t_bp.cpu().numpy() != want).any()


def test_set_devices_split_equals_single_device():
    """bpb_set_devices with the same ordinal twice (one GPU on this box): two full decoders behind one handle, the
    batch split in two contiguous slices; results equal the single-device decode, odd batch sizes included."""
    H = codes.regular_ldpc(1000, 3, 6, seed=1)
    kw = dict(error_rate=0.05, max_iter=50, bp_method="ms", ms_scaling_factor=0.625, input_vector_type="syndrome")
    syn = codes.bsc_syndromes(H, 0.06, 40001, seed=4)
    one = BpDecoder(H, **kw)
    want = one.decode_batch(syn, return_llr=True)
    two = BpDecoder(H, devices=[0, 0], **kw)
    got = two.decode_batch(syn, return_llr=True)
    assert np.array_equal(got, want) and np.array_equal(two.iter_batch, one.iter_batch)
    assert np.array_equal(two.converge_batch, one.converge_batch)
    assert np.array_equal(two.log_prob_ratios_batch, one.log_prob_ratios_batch)
    m = MultiGpuBpDecoder(H, devices=[0, 0, 0], **kw)
    assert np.array_equal(m.decode_batch(syn[:1001]), want[:1001])
    # parameter changes after the split reach every device
    two.max_iter = 3
    one.max_iter = 3
    assert np.array_equal(two.decode_batch(syn[:5000]), one.decode_batch(syn[:5000]))
    assert two.iter_batch.max() == 3
    # BP+OSD through the split
    Hs = codes.rotated_surface_code_x(7)
    ss = codes.bsc_syndromes(Hs, 0.08, 9001, seed=17)
    a = BpOsdDecoder(Hs, error_rate=0.08, osd_method="osd0", max_iter=12, bp_method="ps")
    b = BpOsdDecoder(Hs, error_rate=0.08, osd_method="osd0", max_iter=12, bp_method="ps", devices=[0, 0])
    assert np.array_equal(a.decode_batch(ss), b.decode_batch(ss))


@pytest.mark.parametrize("case", ["ldpc1000", "bb144_osd", "nonuniform"])
def test_monte_carlo_on_device_counts_exactly(case):
    """bpb_mc_bsc draws, decodes and scores on the device; the same Philox errors pushed through decode_batch on the
    host path give the same counters, run for run (exact, not statistical).  Chunking, the run offset and the
    multi-device split do not change the drawn errors."""
    if case == "ldpc1000":
        H, p, runs, osd = codes.regular_ldpc(1000, 3, 6, seed=1), 0.07, 30000, False
        d = BpDecoder(H, error_rate=p, max_iter=50, bp_method="ms", ms_scaling_factor=0.625, input_vector_type="syndrome")
    elif case == "bb144_osd":
        H, p, runs, osd = codes.bivariate_bicycle_144(), 0.03, 50000, True
        d = BpOsdDecoder(H, error_rate=p, max_iter=30, bp_method="ms", ms_scaling_factor=0.625, osd_method="osd0")
    else:
        H, runs, osd = codes.regular_ldpc(240, 3, 6, seed=3), 20000, False
        p = np.linspace(0.0, 0.12, 240)
        p[7] = 1.0
        d = BpDecoder(H, error_channel=p, max_iter=25, bp_method="ps", input_vector_type="syndrome")
    seed, first = 0x1234567890ab, 1000
    got = d.monte_carlo_bsc(runs, seed=seed, first_run=first, with_osd=osd)
    err = philox_bsc_errors(H.shape[1], p, runs, seed=seed, first_run=first)
    syn = codes.syndromes_of(H, err)
    dec = d.decode_batch(syn)
    fails = (dec != err).any(axis=1)
    want = {"run_count": runs, "fail_count": int(fails.sum()), "bp_converged": int(d.converge_batch.sum()),
            "iter_sum": int(d.iter_batch.sum()), "converged_wrong": int((fails & d.converge_batch).sum())}
    assert got == want
    assert 0 < want["fail_count"] < runs
    # split into two calls / two "devices": same totals
    a = d.monte_carlo_bsc(runs // 3, seed=seed, first_run=first, with_osd=osd)
    b = d.monte_carlo_bsc(runs - runs // 3, seed=seed, first_run=first + runs // 3, with_osd=osd)
    assert {k: a[k] + b[k] for k in a} == want
    if case == "ldpc1000":
        m = BpDecoder(H, error_rate=p, max_iter=50, bp_method="ms", ms_scaling_factor=0.625,
                      input_vector_type="syndrome", devices=[0, 0])
        assert m.monte_carlo_bsc(runs, seed=seed, first_run=first) == want
        sim = MonteCarloBscSimulation(H, error_rate=p, Decoder=d, target_run_count=runs, seed=12345, device_side=True,
                                      tqdm_disable=True)
        res = sim.run()
        assert res["run_count"] == runs and 0 < res["fail_count"] < runs
        host = MonteCarloBscSimulation(H, error_rate=p, Decoder=d, target_run_count=runs, seed=1, tqdm_disable=True).run()
        # statistical agreement with the numpy-driven host path (different random stream): 5 sigma
        sig = np.sqrt(host["logical_error_rate"] * (1 - host["logical_error_rate"]) * 2 / runs)
        assert abs(host["logical_error_rate"] - res["logical_error_rate"]) < 5 * sig + 1e-9


def test_pageable_and_pinned_inputs_agree_and_pool_recycles():
    H = codes.regular_ldpc(1000, 3, 6, seed=1)
    d = BpDecoder(H, error_rate=0.05, max_iter=50, bp_method="ms", ms_scaling_factor=0.625, input_vector_type="syndrome")
    B = (1 << 18) + 77
    syn = codes.bsc_syndromes(H, 0.05, B, seed=9)
    a = d.decode_batch(syn).copy()
    pin = _capi.PinnedArray(syn.shape, np.uint8)
    pin.array[...] = syn
    b = d.decode_batch(pin.array).copy()
    assert np.array_equal(a, b)
    before = _capi._POOL_BYTES[0]
    for nb in (1000, 3000, 70000, B // 2, 5, B):
        d.decode_batch(syn[:nb])
    assert _capi._POOL_BYTES[0] <= max(before, 1) * 2 + (64 << 20)  # power-of-two buckets are reused, not leaked


@pytest.mark.parametrize("method", ["ms", "ps"])
def test_serial_relative_schedule_matches_oracle(port_oracle, method):
    """schedule='serial_relative' (bp.hpp:469-482: the schedule is re-sorted by posterior LLR before
    every sweep, with libstdc++'s std::sort -- ties and all).  Every row of a batch starts from the configured order.
    Bar: decisions, converge, iterations exact; LLRs bit-identical for min-sum, 1e-5 for product-sum."""
    for H, p, B, iters in ((codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 96, 40),
                           (codes.regular_ldpc(240, 3, 6, seed=3), 0.075, 160, 25),
                           (codes.rotated_surface_code_x(13), 0.05, 200, 20),
                           (codes.hamming_code(5), 0.1, 64, 5)):
        syn = codes.bsc_syndromes(H, p, B, seed=17)
        kw = dict(max_iter=iters, bp_method=method, schedule="serial_relative", ms_scaling_factor=0.625)
        want = port_oracle.decode_batch(H, syn, p, **kw)
        d = BpDecoder(H, error_rate=p, input_vector_type="syndrome", **kw)
        got = d.decode_batch(syn, return_llr=True)
        assert_same_decode((got, d.converge_batch, d.iter_batch, d.log_prob_ratios_batch), want,
                           llr_exact=(method == "ms"))


def test_serial_relative_custom_order_and_decode_state(port_oracle):
    """A custom initial order, and the state the reference object carries from one decode() to the next: its
    serial_schedule_order member is sorted in place, so decode k+1 starts from the order decode k ended with."""
    H = codes.regular_ldpc(240, 3, 6, seed=3)
    order = np.random.default_rng(8).permutation(240)
    syn = codes.bsc_syndromes(H, 0.07, 40, seed=23)
    kw = dict(max_iter=25, bp_method="ms", schedule="serial_relative", ms_scaling_factor=0.625)
    want = port_oracle.decode_batch(H, syn, 0.07, serial_schedule_order=order, **kw)
    d = BpDecoder(H, error_rate=0.07, input_vector_type="syndrome", serial_schedule_order=[int(x) for x in order], **kw)
    got = d.decode_batch(syn, return_llr=True)
    assert_same_decode((got, d.converge_batch, d.iter_batch, d.log_prob_ratios_batch), want, llr_exact=True)
    # sequential decode(): feed the oracle the order the previous decode ended with
    cur = order.copy()
    for b in range(8):
        if not syn[b].any():
            continue
        w = port_oracle.decode_batch(H, syn[b:b + 1], 0.07, serial_schedule_order=cur, **kw)
        out = d.decode(syn[b])
        assert np.array_equal(out, w[0][0]) and d.iter == int(w[2][0]) and d.converge == bool(w[1][0])
        new = np.asarray(d.serial_schedule_order)
        assert sorted(new.tolist()) == list(range(240))
        cur = new


@pytest.mark.parametrize("mk,p,osd", [(lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, False),
                                      (codes.bivariate_bicycle_144, 0.02, True),
                                      (lambda: codes.rotated_surface_code_x(13), 0.05, True)],
                         ids=["ldpc1000_bp", "bb144_bposd", "surface13_bposd"])
def test_bit_packed_io_equals_unpacked(mk, p, osd):
    """bpb_decode_batch_b8 (stim's b8 rows in, b8 decisions and observable parities out, everything in between on the
    device) against decode_batch on the unpacked syndromes; m and n that are no multiples of 8 / 32 (surface code:
    84 x 169) and garbage in the pad bits of the last input byte."""
    H = mk()
    m, n = H.shape
    B = 5000
    syn = codes.bsc_syndromes(H, p, B, seed=31)
    kw = dict(max_iter=20, bp_method="ms", ms_scaling_factor=0.625)
    cls = BpOsdDecoder if osd else BpDecoder
    extra = dict(osd_method="osd0") if osd else dict(input_vector_type="syndrome")
    d = cls(H, error_rate=p, **kw, **extra)
    want = d.decode_batch(syn)
    want_conv, want_its = d.converge_batch.copy(), d.iter_batch.copy()
    packed = np.packbits(syn, axis=1, bitorder="little")
    if m % 8:
        packed[:, -1] |= np.uint8((0xff << (m % 8)) & 0xff)  # pad bits must be ignored
    rng = np.random.default_rng(5)
    O = (rng.random((11, n)) < 0.1).astype(np.uint8)
    d.set_observables(O)
    dec8, obs8 = d.decode_batch_b8(packed, decoding=True, observables=True)
    got = np.unpackbits(dec8, axis=1, bitorder="little")[:, :n]
    assert np.array_equal(got, want)
    assert np.array_equal(d.converge_batch, want_conv) and np.array_equal(d.iter_batch, want_its)
    want_obs = (want.astype(np.int32) @ O.T.astype(np.int32)) % 2
    assert np.array_equal(np.unpackbits(obs8, axis=1, bitorder="little")[:, :11], want_obs.astype(np.uint8))
    only_obs = d.decode_batch_b8(packed, decoding=False, observables=True)
    assert np.array_equal(only_obs, obs8)


@pytest.mark.parametrize("mk,p", [(lambda: codes.rep_code(7), 0.1), (lambda: codes.regular_ldpc(240, 3, 6, seed=3), 0.05),
                                  (lambda: codes.rotated_surface_code_x(7), 0.05), (codes.bivariate_bicycle_144, 0.02),
                                  (lambda: codes.regular_ldpc(1000, 3, 6, seed=1), 0.04)],
                         ids=["rep7", "ldpc240", "surface7", "bb144", "ldpc1000"])
def test_soft_info_decoder_matches_oracle(port_oracle, mk, p):
    """SoftInfoBpDecoder (bp.hpp:547-665) on the device against the oracle: decisions, converge flags,
    iteration counts, posterior LLRs and the final soft syndromes, all bit-identical (min-sum arithmetic only)."""
    from ldpc_b200 import SoftInfoBpDecoder
    H = mk()
    rng = np.random.default_rng(12)
    B = 150
    err = (rng.random((B, H.shape[1])) < p).astype(np.uint8)
    syn = codes.syndromes_of(H, err)
    order = rng.permutation(H.shape[1])
    for sigma, cutoff, ms, use_order in ((0.6, np.inf, 1.0, False), (0.8, 3.0, 0.625, True), (0.3, 1.0, 1.0, False)):
        soft = (1 - 2.0 * syn) + rng.normal(0, sigma, size=syn.shape)
        kw = dict(serial_schedule_order=[int(x) for x in order]) if use_order else {}
        d = SoftInfoBpDecoder(H, error_rate=p, max_iter=20, ms_scaling_factor=ms, cutoff=cutoff, sigma=sigma, **kw)
        got = d.decode_batch(soft, return_llr=True)
        want = port_oracle.soft_info_decode_batch(H, soft, p, 20, ms, cutoff, sigma,
                                                  serial_schedule_order=order if use_order else None)
        assert np.array_equal(got, want[0])
        assert np.array_equal(d.converge_batch, want[1]) and np.array_equal(d.iter_batch, want[2])
        assert np.array_equal(d.log_prob_ratios_batch.view(np.uint64), want[3].view(np.uint64))
        assert np.array_equal(d.soft_syndrome_batch.view(np.uint64), want[4].view(np.uint64))
        one = d.decode(soft[3])
        assert np.array_equal(one, want[0][3]) and d.iter == int(want[2][3]) and d.converge == bool(want[1][3])
        assert np.array_equal(d.soft_syndrome.view(np.uint64), want[4][3].view(np.uint64))
