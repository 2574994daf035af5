#!/usr/bin/env python
"""Strong scaling of ONE host batch over the GPUs of one box, single process (BASELINE.json north_star: "host-side
split and final gather"): one C-ABI handle, bpb_set_devices(k), one pinned host input array split contiguously, the
D2H copies land in disjoint ranges of one pinned host output array.

Usage: python scripts/strong_scaling.py [--config 2] [--total 8388608] [--gpus 1,2,4,8] [--reps 3]
Prints one JSON line per device count: decodes/s end to end (host buffers in, host buffers out), speed-up over one
device, and the same through the Python class (BpDecoder(devices=[...]).decode_batch(pinned in, out=pinned out))."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (config table, host syndromes)
from ldpc_b200 import BpDecoder, BpOsdDecoder, _capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--total", type=int, default=8 << 20)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--b8", action="store_true", help="bit-packed I/O (bpb_decode_batch_b8): 1/8 of the PCIe bytes")
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    cfg = bench.config_table(args.config)
    H = cfg["H"]().tocsr()
    m, n = H.shape
    total = args.total
    # distinct syndromes for the first 2^20 rows, tiled to the full batch (the decode of a row does not depend on
    # the others; generating 8M x n errors on the host would take longer than the measurement)
    base = bench.host_syndromes(cfg, H, min(total, 1 << 20))
    pin_in = _capi.PinnedArray((total, m), np.uint8)
    for lo in range(0, total, base.shape[0]):
        hi = min(total, lo + base.shape[0])
        pin_in.array[lo:hi] = base[: hi - lo]
    if args.b8:
        mb8, nb8 = (m + 7) // 8, (n + 7) // 8
        pk_in = _capi.PinnedArray((total, mb8), np.uint8)
        step_rows = 1 << 18
        for lo in range(0, total, step_rows):
            pk_in.array[lo:lo + step_rows] = np.packbits(pin_in.array[lo:lo + step_rows], axis=1, bitorder="little")
        pk_out = _capi.PinnedArray((total, nb8), np.uint8)
    pin_dec = _capi.PinnedArray((total, n), np.uint8)
    pin_conv = _capi.PinnedArray((total,), np.uint8)
    pin_its = _capi.PinnedArray((total,), np.int32)
    L = _capi.lib()
    one = None
    ref_dec = None
    for k in [int(x) for x in args.gpus.split(",")]:
        if k > have:
            print(json.dumps({"gpus": k, "skipped": f"only {have} devices visible"}))
            continue
        cls = BpOsdDecoder if cfg["osd"] else BpDecoder
        kw = dict(cfg["kw"])
        if cfg["osd"]:
            kw["osd_method"] = "osd0"
        d = cls(H, error_rate=cfg["p"], input_vector_type="syndrome", devices=list(range(k)), **kw)
        h = d._ensure_handle()

        def step():
            if args.b8:
                rc = L.bpb_decode_batch_b8(h, 1 if cfg["osd"] else 0, _capi.host_ptr(pk_in.array), total,
                                           _capi.host_ptr(pk_out.array), None, _capi.host_ptr(pin_conv.array),
                                           _capi.host_ptr(pin_its.array))
            elif cfg["osd"]:
                rc = L.bpb_bposd_decode_batch(h, _capi.host_ptr(pin_in.array), total, _capi.host_ptr(pin_dec.array),
                                              _capi.host_ptr(pin_conv.array), _capi.host_ptr(pin_its.array), None, 0)
            else:
                rc = L.bpb_decode_batch(h, 0, _capi.host_ptr(pin_in.array), total, _capi.host_ptr(pin_dec.array),
                                        _capi.host_ptr(pin_conv.array), _capi.host_ptr(pin_its.array), None)
            _capi.check(h, rc)

        step()  # warm-up: allocations, first-touch of the staging buffers
        step()
        times = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            step()
            times.append(time.perf_counter() - t0)
        best = min(times)
        if args.b8:
            pin_dec.array[: 1 << 16] = np.unpackbits(pk_out.array[: 1 << 16], axis=1, bitorder="little")[:, :n]
        if ref_dec is None:
            ref_dec = pin_dec.array[: 1 << 16].copy()
            same = True
        else:
            same = bool(np.array_equal(ref_dec, pin_dec.array[: 1 << 16]))
        # the Python class on the same pinned arrays
        t_py = 1e30
        for _ in range(0 if args.b8 else 2):
            t0 = time.perf_counter()
            out = d.decode_batch(pin_in.array) if cfg["osd"] else d.decode_batch(pin_in.array, out=pin_dec.array)
            t_py = min(t_py, time.perf_counter() - t0)
        if not args.b8:
            same = same and bool(np.array_equal(out[: 1 << 16], ref_dec))
            del out
        rate = total / best
        if one is None:
            one = rate
        print(json.dumps({"config": args.config, "workload": bench.workload_name(cfg, total), "gpus": k,
                          "scaling": "strong", "total_syndromes": total, "seconds": best, "decodes_per_s": rate,
                          "speedup_vs_first": rate / one,
                          "python_api_decodes_per_s": None if args.b8 else total / t_py,
                          "io": "b8 (bit-packed)" if args.b8 else "uint8 per bit",
                          "h2d_bytes": total * ((m + 7) // 8 if args.b8 else m),
                          "d2h_bytes": total * (((n + 7) // 8 if args.b8 else n) + 5),
                          "host_gb_per_s": total * (((m + 7) // 8 + (n + 7) // 8) if args.b8 else (m + n)) / best / 1e9,
                          "matches_first_run": same,
                          "times": times}), flush=True)
        del d


if __name__ == "__main__":
    main()
