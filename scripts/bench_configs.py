#!/usr/bin/env python
"""Throughput + parity spot-check on the other BASELINE.json configurations (3: surface d=13 product-sum,
4: [[144,12,12]] BB min-sum + OSD-0, 5: n=10^4 (3,6)-LDPC serial min-sum 100 iterations).  One JSON line per
configuration on stdout.  These are not bench.py lines (the metric is quoted on config 2); they document that the
same kernels serve the other codes and how fast.  Usage: python scripts/bench_configs.py [--small]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (checker only: parity spot-check on a subsample)
from ldpc_b200 import BpDecoder, BpOsdDecoder, codes  # noqa: E402


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    args = ap.parse_args()
    port = oracle.PortOracle()
    scale = 16 if args.small else 1

    def report(name, H, B, p, dec, check_kw, osd=False, sub=256):
        syn = codes.bsc_syndromes(H, p, B, seed=7)
        dt, out = timed(lambda: dec.decode_batch(syn))
        want = port.decode_batch(H, syn[:sub], p, **check_kw)
        ok = bool(np.array_equal(dec.iter_batch[:sub], want[2]) and np.array_equal(dec.converge_batch[:sub], want[1]))
        if osd:
            w = want[0].copy()
            bad = ~want[1]
            if bad.any():
                w[bad] = port.osd0_batch(H, syn[:sub][bad], want[3][bad])
            ok = ok and bool(np.array_equal(out[:sub], w))
            ok = ok and bool(np.array_equal(codes.syndromes_of(H, out[: 4096]), syn[: 4096]))
        else:
            ok = ok and bool(np.array_equal(out[:sub], want[0]))
        inf = dec.info()
        print(json.dumps({"config": name, "batch": B, "decodes_per_s_e2e_pageable": B / dt, "seconds": dt,
                          "mean_iterations": float(dec.iter_batch.mean()), "converged": float(dec.converge_batch.mean()),
                          "kernel_family": {1: "stream", 2: "smem"}.get(inf["kernel_family"]),
                          "kernel_ms_last_chunk": inf["last_kernel_ms"], "parity_subsample_ok": ok}), flush=True)

    H = codes.rotated_surface_code_x(13)
    kw = dict(max_iter=30, bp_method="ps", schedule="parallel")
    report("3: surface d=13 X checks, product_sum 30 it, p=0.05", H, (1 << 20) // scale, 0.05,
           BpDecoder(H, error_rate=0.05, input_vector_type="syndrome", **kw), kw)

    H = codes.bivariate_bicycle_144()
    kw = dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625)
    report("4: [[144,12,12]] BB, min_sum 50 it + OSD-0, p=0.003", H, (1 << 20) // scale, 0.003,
           BpOsdDecoder(H, error_rate=0.003, osd_method="osd0", **kw), kw, osd=True, sub=4096)
    report("4b: [[144,12,12]] BB, min_sum 50 it + OSD-0, p=0.02 (BP fails more often)", H, (1 << 18) // scale, 0.02,
           BpOsdDecoder(H, error_rate=0.02, osd_method="osd0", **kw), kw, osd=True, sub=4096)

    H = codes.regular_ldpc(10000, 3, 6, seed=1)
    kw = dict(max_iter=100, bp_method="ms", schedule="serial", ms_scaling_factor=0.625)
    for p in (0.02, 0.05, 0.08):
        report(f"5: (3,6) n=10^4, serial min_sum 100 it, p={p}", H, (1 << 16) // scale, p,
               BpDecoder(H, error_rate=p, input_vector_type="syndrome", **kw), kw, sub=24)
    kw = dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625)
    report("5b: (3,6) n=10^4, parallel min_sum 50 it, p=0.05 (streaming family, too large for shared memory)", H,
           (1 << 17) // scale, 0.05, BpDecoder(H, error_rate=0.05, input_vector_type="syndrome", **kw), kw, sub=24)


if __name__ == "__main__":
    main()
