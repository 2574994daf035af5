#!/usr/bin/env python
"""Latency / throughput of the three kernel families at small batch sizes (device-resident, CUDA events, median of 20)
and of a single BpDecoder.decode() call.  Usage: python scripts/latency_table.py [config: 2|3]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldpc_b200 import BpDecoder, _capi, codes  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if cfg == 2:
    H, p, kw = codes.regular_ldpc(1000, 3, 6, seed=1), 0.05, dict(max_iter=50, bp_method="ms", ms_scaling_factor=0.625)
else:
    H, p, kw = codes.rotated_surface_code_x(13), 0.05, dict(max_iter=30, bp_method="ps")
m, n = H.shape
dev = torch.device("cuda", 0)
L = _capi.lib()
st = torch.cuda.current_stream(dev)
rows = []
for B in (1, 8, 32, 148, 296, 1024, 4096, 16384):
    syn = codes.bsc_syndromes(H, p, B, seed=5)
    d_syn = torch.from_numpy(syn).to(dev)
    d_dec = torch.empty((B, n), dtype=torch.uint8, device=dev)
    d_conv = torch.empty(B, dtype=torch.uint8, device=dev)
    d_its = torch.empty(B, dtype=torch.int32, device=dev)
    row = {"config": cfg, "batch": B}
    for fam in ("pair", "smem", "stream", "edge"):
        d = BpDecoder(H, error_rate=p, input_vector_type="syndrome", kernel=fam, **kw)
        h = d._ensure_handle()

        def step():
            rc = L.bpb_decode_batch_device(h, 0, C.c_void_p(d_syn.data_ptr()), B, C.c_void_p(d_dec.data_ptr()),
                                           C.c_void_p(d_conv.data_ptr()), C.c_void_p(d_its.data_ptr()), None,
                                           C.c_void_p(st.cuda_stream))
            _capi.check(h, rc)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            step()
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        row[fam + "_us"] = round(float(np.median(ts)), 1)
        del d
    row["mean_it"] = float(d_its.float().mean().item())
    rows.append(row)
    print(json.dumps(row), flush=True)
# single .decode() through the Python class (host call, includes H2D/D2H and the launch overheads)
for fam in ("auto", "pair", "edge"):
    d = BpDecoder(H, error_rate=p, input_vector_type="syndrome", kernel=fam, **kw)
    syn = codes.bsc_syndromes(H, p, 64, seed=6)
    syn = syn[syn.any(axis=1)]
    d.decode(syn[0])
    t0 = time.perf_counter()
    for s in syn:
        d.decode(s)
    print(json.dumps({"config": cfg, "single_decode_us": round((time.perf_counter() - t0) / len(syn) * 1e6, 1), "kernel": fam}))
