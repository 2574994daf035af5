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


def usable_cores():
    """CPU threads this process may use: min(affinity, cgroup quota)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(round(int(quota) / int(period)))))
    except Exception:
        pass
    return max(1, n)


def checker_decode(H, syn, channel, osd=False, want_llr=True, **kw):
    """Reference results for a (large) batch on all host cores.

    Uses the UNMODIFIED reference C++ (oracle/_ref, multi-threaded, one decoder object per thread) when the prebuilt
    library is present, else the plain-C restatement (oracle/_build) fanned out over a thread pool.  Returns
    (decoding, converged, iters, llr|None, kind); with osd=True `decoding` is the BP+OSD-0 output.
    """
    import oracle
    cores = usable_cores()
    if oracle.have_ref():
        ref = oracle.RefOracle()
        out = ref.decode_batch(H, syn, channel, want_llr=want_llr, osd_method=1 if osd else 0, threads=cores, **kw)
        return out[0], out[1], out[2], out[3], "reference"
    if not oracle.have_port():
        oracle.build()
    port = oracle.PortOracle()
    from concurrent.futures import ThreadPoolExecutor
    B = syn.shape[0]
    bounds = [(B * t // cores, B * (t + 1) // cores) for t in range(cores)]

    def work(lohi):
        lo, hi = lohi
        if hi <= lo:
            return None
        r = port.decode_batch(H, syn[lo:hi], channel, want_llr=(want_llr or osd), **kw)
        dec = r[0]
        if osd:
            bad = ~r[1]
            if bad.any():
                dec = dec.copy()
                dec[bad] = port.osd0_batch(H, syn[lo:hi][bad], r[3][bad])
        return dec, r[1], r[2], (r[3] if want_llr else None)

    with ThreadPoolExecutor(cores) as ex:
        parts = [p for p in ex.map(work, bounds) if p is not None]
    cat = lambda k: np.concatenate([p[k] for p in parts])
    return cat(0), cat(1), cat(2), (cat(3) if want_llr else None), "port"


def philox_bsc_errors(n, flip_prob, runs, seed=0, first_run=0):
    """The errors bpb_mc_bsc draws (include/bp_b200.h): run r flips bit j iff word j % 4 of Philox4x32-10(counter =
    (r_lo, r_hi, j // 4, 0), key = (seed_lo, seed_hi)) is below floor(p_j * 2^32).  Vectorised numpy restatement of the
    published Philox4x32-10 (Salmon et al., SC'11) for the tests."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    r = (np.arange(runs, dtype=np.uint64) + np.uint64(first_run))[:, None]
    groups = (n + 3) // 4
    g = np.arange(groups, dtype=np.uint64)[None, :]
    c0 = np.broadcast_to(r & np.uint64(0xffffffff), (runs, groups)).copy()
    c1 = np.broadcast_to(r >> np.uint64(32), (runs, groups)).copy()
    c2 = np.broadcast_to(g, (runs, groups)).copy()
    c3 = np.zeros((runs, groups), np.uint64)
    k0, k1 = np.uint64(seed & 0xffffffff), np.uint64((seed >> 32) & 0xffffffff)
    mask = np.uint64(0xffffffff)
    for _ in range(10):
        p0 = c0 * np.uint64(M0)
        p1 = c2 * np.uint64(M1)
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    words = np.stack([c0, c1, c2, c3], axis=2).reshape(runs, groups * 4)[:, :n]
    p = np.broadcast_to(np.asarray(flip_prob, dtype=np.float64), (n,))
    thresh = np.floor(np.clip(p, 0.0, 1.0) * 4294967296.0).astype(np.uint64)
    return (words < thresh[None, :]).astype(np.uint8)
