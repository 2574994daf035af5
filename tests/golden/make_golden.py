"""Generate tests/golden/*.npz from the UNMODIFIED reference C++ (oracle/_ref/libref_bp.so, built from
/root/reference/src_cpp by oracle/Makefile).  Run in the build container: `python tests/golden/make_golden.py`.
The fixtures are small (tens of syndromes per BASELINE config) and are committed; they let machines without
/root/reference check both the oracle restatement and the CUDA path against reference outputs."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from ldpc_b200 import codes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CONFIGS = [
    # BASELINE.json configs[0]: hamming_code(5), product_sum, max_iter=2 (README example)
    ("cfg1_hamming5_ps", codes.hamming_code(5), 0.1, 64, 0, dict(max_iter=2, bp_method="ps", schedule="parallel", ms_scaling_factor=1.0)),
    # configs[1]: (3,6)-regular n=1000, min_sum 50 iterations, BSC p=0.05 (+ a harder tail so failures are covered)
    ("cfg2_ldpc1000_ms", codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 48, 16, dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625)),
    ("cfg2_ldpc1000_ms_adaptive", codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 24, 8, dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.0)),
    ("cfg2_ldpc1000_ps", codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 24, 8, dict(max_iter=50, bp_method="ps", schedule="parallel", ms_scaling_factor=1.0)),
    # configs[2]: d=13 rotated surface code X checks, product_sum 30 iterations
    ("cfg3_surface13_ps", codes.rotated_surface_code_x(13), 0.05, 96, 0, dict(max_iter=30, bp_method="ps", schedule="parallel", ms_scaling_factor=1.0)),
    # configs[3]: [[144,12,12]] BB code, min-sum (BP part; OSD-0 is covered by test_port_osd0_matches_reference)
    ("cfg4_bb144_ms", codes.bivariate_bicycle_144(), 0.02, 128, 0, dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625)),
    # configs[4]: serial-schedule min-sum (n=1000 here to keep the fixture small; n=10^4 is run live on the GPU box)
    ("cfg5_ldpc1000_ms_serial", codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 32, 8, dict(max_iter=100, bp_method="ms", schedule="serial", ms_scaling_factor=0.625)),
    ("cfg5_ldpc1000_ps_serial", codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 16, 4, dict(max_iter=50, bp_method="ps", schedule="serial", ms_scaling_factor=1.0)),
    # SERIAL_RELATIVE (the schedule re-sorted by LLR every iteration, bp.hpp:469-482); every syndrome is
    # decoded from the initial order 0..n-1 (oracle/ref_wrap.cpp resets the member before each decode)
    ("f4_ldpc1000_ms_serial_relative", codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, 24, 8, dict(max_iter=50, bp_method="ms", schedule="serial_relative", ms_scaling_factor=0.625)),
    ("f4_bb144_ps_serial_relative", codes.bivariate_bicycle_144(), 0.02, 64, 0, dict(max_iter=30, bp_method="ps", schedule="serial_relative", ms_scaling_factor=1.0)),
]


# soft-information decoding (bp.hpp:547-665): name, H, p, shots, sigma, cutoff, ms_scaling_factor, max_iter
SOFT_CONFIGS = [
    ("f4_softinfo_ldpc240", codes.regular_ldpc(240, 3, 6, seed=3), 0.05, 48, 0.8, 3.0, 0.625, 20),
    ("f4_softinfo_surface7", codes.rotated_surface_code_x(7), 0.05, 64, 0.5, 2.0, 1.0, 15),
]


def main_soft(ref, only):
    for name, H, p, B, sigma, cutoff, ms, max_iter in SOFT_CONFIGS:
        if only and name not in only:
            continue
        rng = np.random.default_rng(21)
        err = (rng.random((B, H.shape[1])) < p).astype(np.uint8)
        soft = (1 - 2.0 * codes.syndromes_of(H, err)) + rng.normal(0, sigma, size=(B, H.shape[0]))
        dec, conv, its, llr, soft_out = ref.soft_info_decode_batch(H, soft, p, max_iter, ms, cutoff, sigma)
        coo = sp.coo_matrix(H)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), rows=coo.row.astype(np.int32),
                            cols=coo.col.astype(np.int32), shape=np.asarray(H.shape), channel=np.full(H.shape[1], p),
                            soft=soft, decoding=dec, converged=conv, iters=its, llr=llr, soft_out=soft_out,
                            max_iter=max_iter, ms_scaling_factor=ms, cutoff=cutoff, sigma=sigma, kind="soft_info")
        print(name, soft.shape, "converged", conv.mean(), "mean iters", its.mean())


def main():
    ref = oracle.RefOracle()
    main_soft(ref, set(sys.argv[1:]))
    only = set(sys.argv[1:])  # optional: names of the fixtures to (re)generate
    for name, H, p, B, Bhard, kw in CONFIGS:
        if only and name not in only:
            continue
        syn = codes.bsc_syndromes(H, p, B, seed=7)
        if Bhard:
            syn = np.concatenate([syn, codes.bsc_syndromes(H, 1.8 * p, Bhard, seed=8)])
        dec, conv, its, llr = ref.decode_batch(H, syn, p, **kw)
        coo = sp.coo_matrix(H)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), rows=coo.row.astype(np.int32),
                            cols=coo.col.astype(np.int32), shape=np.asarray(H.shape), channel=np.full(H.shape[1], p),
                            syndromes=syn, decoding=dec, converged=conv, iters=its, llr=llr,
                            max_iter=kw["max_iter"], bp_method=kw["bp_method"], schedule=kw["schedule"],
                            ms_scaling_factor=kw["ms_scaling_factor"])
        print(name, syn.shape, "converged", conv.mean(), "mean iters", its.mean())


if __name__ == "__main__":
    main()
